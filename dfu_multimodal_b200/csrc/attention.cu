// Fused softmax attention for the ViT branch (timm Attention: softmax(q k^T / 8) v, 12 heads of 64, 197 tokens),
// forward and backward, on tcgen05 tensor cores with TMEM accumulators and TMA-fed operands.
//
// One persistent CTA per SM walks the (batch, head) pairs.  The whole sequence of one head fits on chip, so there
// is no online-softmax rescaling: S = Q K^T lands in TMEM, the softmax warps read it back row by row (one thread per
// row), write the unnormalised probabilities as a bf16 K-major UMMA operand into shared memory and the second MMA
// produces O, which is scaled by 1/rowsum on its way out.  Only O and the row log-sum-exp leave the chip; the
// [B,H,T,T] probability tensor of the unfused path is never materialised.
//
// Backward recomputes P from Q, K and the saved log-sum-exp in the transposed orientation (key rows on TMEM lanes):
//   S^T = K Q^T, dP^T = V dO^T  ->  P^T = exp(scale S^T - lse), dS^T = scale P^T (dP^T - D)   (threads, -> smem)
//   dV += P^T dO, dK += dS^T Q   (A = the smem tiles, K-major)      dQ += dS K   (A = the same dS^T tile, MN-major)
// in 2 x 2 (key tile, query chunk) steps of 128 x 128, all five accumulators in the 512 TMEM columns.
//
// Warp roles (640 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 softmax /
// epilogue (warp w owns TMEM lanes 32*(w%4)..+31; the four warps of a lane quarter split the columns).
#include <stdlib.h>
#include "tc_common.cuh"
#include "../../include/dfu_b200.h"

namespace {
using namespace tc;

constexpr int HD = 64;                        // head dim
constexpr int THREADS = 640;                  // 4 control warps + 16 softmax / epilogue warps
constexpr int KV_ROWS = 208;                  // 197 keys padded to a multiple of 16
constexpr int TILE128_BYTES = 128 * 128;      // 128 rows x 64 bf16
constexpr int KV_BYTES = KV_ROWS * 128;       // 26624 (26 x 1024: keeps every buffer 1024-byte aligned)
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void st_global_v4(void* p, const bf16x8& v) {
    const uint4 u = *reinterpret_cast<const uint4*>(&v);
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, const bf16x8& v) {
    const uint4 u = *reinterpret_cast<const uint4*>(&v);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ bf16x8 ld_global_v4(const void* p) {
    uint4 u;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    return *reinterpret_cast<bf16x8*>(&u);
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int F_Q0 = 0, F_Q1 = TILE128_BYTES, F_K = 2 * TILE128_BYTES, F_V = F_K + KV_BYTES, F_P0 = F_V + KV_BYTES,
              F_P1 = F_P0 + 4 * TILE128_BYTES, F_END = F_P1 + 4 * TILE128_BYTES;
constexpr int FWD_SMEM = F_END + 1024;

__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                bf16* __restrict__ out, float* __restrict__ lse, int B, int T, int H, float scale) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t qk_full, qk_empty, v_full, v_empty, s_full[2], p_full[2], o_full[2], o_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float pmax_s[2][2][128], psum_s[2][2][128];   // partial row max / sum exchanged by the two column halves
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NK = (T + 15) & ~15;            // key columns of S (multiple of 16)
    const int items = B * H;

    if (warp_idx == 0 && lane == 0) { prefetch_tmap(&map_q); prefetch_tmap(&map_kv); }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(&qk_full, 1); mbar_init(&qk_empty, 1); mbar_init(&v_full, 1); mbar_init(&v_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 8);
        }
        fence_mbar_init();
    }
    if (warp_idx == 2) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp_idx == 0) {
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                const int b = w / H, h = w - b * H;
                const int row0 = b * T;
                mbar_wait(&qk_empty, ph ^ 1);
                mbar_expect_tx(&qk_full, 2 * TILE128_BYTES + KV_BYTES);
                tma_load_2d(&map_q, &qk_full, smem + F_Q0, h * HD, row0);
                tma_load_2d(&map_q, &qk_full, smem + F_Q1, h * HD, row0 + 128);
                tma_load_2d(&map_kv, &qk_full, smem + F_K, (H + h) * HD, row0);
                mbar_wait(&v_empty, ph ^ 1);
                mbar_expect_tx(&v_full, KV_BYTES);
                tma_load_2d(&map_kv, &v_full, smem + F_V, (2 * H + h) * HD, row0);
                if (w + (int)gridDim.x < items) {        // next head's operands -> L2 while this one is computed
                    const int wn = w + gridDim.x, bn = wn / H, hn = wn - bn * H;
                    tma_prefetch_2d(&map_q, hn * HD, bn * T);
                    tma_prefetch_2d(&map_q, hn * HD, bn * T + 128);
                    tma_prefetch_2d(&map_kv, (H + hn) * HD, bn * T);
                    tma_prefetch_2d(&map_kv, (2 * H + hn) * HD, bn * T);
                }
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc(128, NK, 0, 0);
            const uint32_t idesc_o = make_idesc(128, HD, 0, 1);
            const uint32_t sQ = smem_u32(smem + F_Q0), sK = smem_u32(smem + F_K), sV = smem_u32(smem + F_V),
                           sP = smem_u32(smem + F_P0);
            const int nks = NK / 16;
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                mbar_wait(&qk_full, ph);
                tc_fence_after();
                for (int i = 0; i < 2; ++i) {
                    mbar_wait(&o_empty[i], ph ^ 1);      // the previous head's O_i has been read out of TMEM
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        tc_mma_f16(tmem + i * 256, make_desc(sQ + i * TILE128_BYTES + k * 32, 16, 1024),
                                   make_desc(sK + k * 32, 16, 1024), idesc_s, k > 0);
                    tc_commit(&s_full[i]);
                }
                tc_commit(&qk_empty);                    // Q and K may be overwritten once both S tiles are done
                mbar_wait(&v_full, ph);
                tc_fence_after();
                for (int i = 0; i < 2; ++i) {
                    mbar_wait(&p_full[i], ph);           // softmax warps wrote P_i (and finished reading S_i)
                    tc_fence_after();
                    for (int k = 0; k < nks; ++k)
                        tc_mma_f16(tmem + i * 256,
                                   make_desc(sP + i * 4 * TILE128_BYTES + (k >> 2) * TILE128_BYTES + (k & 3) * 32, 16, 1024),
                                   make_desc(sV + k * 2048, 8192, 1024), idesc_o, k > 0);
                    tc_commit(&o_full[i]);
                }
                tc_commit(&v_empty);
            }
        }
    } else if (warp_idx >= 4) {
        // 8 warps per query tile: two per TMEM lane quarter, each taking every other 32-column chunk of the row.
        const int ew = warp_idx - 4;
        const int i = ew >> 3;                           // query tile
        const int half = (ew >> 2) & 1;
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;             // row inside the 128-query tile == TMEM lane
        const int q = i * 128 + row;
        const float sl2 = scale * LOG2E;
        const uint32_t sP = smem_u32(smem + (i ? F_P1 : F_P0)) + row * 128;
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + i * 256;
        const int nchunks = (NK + 31) / 32;
        const int my_chunks = (nchunks - half + 1) / 2;  // chunks half, half+2, ...
        int it = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            const int b = w / H, h = w - b * H;
            mbar_wait(&s_full[i], ph);
            tc_fence_after();
            uint32_t ra[32], rb[32];
            // pass 1: row maximum of the raw scores (the next chunk's TMEM load is in flight while this one is reduced)
            float m = -INFINITY;
            tmem_ld32(taddr + half * 32, ra);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                if (cc < my_chunks) {
                    const int c = half + 2 * cc;
                    uint32_t* cur = (cc & 1) ? rb : ra;
                    uint32_t* nxt = (cc & 1) ? ra : rb;
                    tmem_ld_wait();
                    if (cc + 1 < my_chunks) tmem_ld32(taddr + (c + 2) * 32, nxt);
                    if (c * 32 + 32 <= T) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(cur[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j < T) m = fmaxf(m, __uint_as_float(cur[j]));
                    }
                }
            }
            pmax_s[i][half][row] = m;
            named_bar_sync(1 + i, 256);
            m = fmaxf(m, pmax_s[i][half ^ 1][row]);
            // pass 2: p = exp(scale (s - m)), partial row sum, bf16 P into the K-major SWIZZLE_128B operand tile
            const float ms = m * sl2;
            float sum = 0.f;
            tmem_ld32(taddr + half * 32, ra);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                if (cc < my_chunks) {
                    const int c = half + 2 * cc;
                    uint32_t* cur = (cc & 1) ? rb : ra;
                    uint32_t* nxt = (cc & 1) ? ra : rb;
                    tmem_ld_wait();
                    if (cc + 1 < my_chunks) tmem_ld32(taddr + (c + 2) * 32, nxt);
                    const bool full = c * 32 + 32 <= T;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int col0 = c * 32 + g * 8;
                        if (col0 < NK) {                 // warp-uniform
                            float p[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float e = ex2_approx(__uint_as_float(cur[g * 8 + j]) * sl2 - ms);
                                p[j] = (full || col0 + j < T) ? e : 0.f;
                                sum += p[j];
                            }
                            const int atom = col0 >> 6, cj = (col0 & 63) >> 3;
                            st_shared_v4(sP + atom * TILE128_BYTES + ((cj ^ (row & 7)) << 4), pack8(p));
                        }
                    }
                }
            }
            psum_s[i][half][row] = sum;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[i]);
            named_bar_sync(1 + i, 256);
            sum += psum_s[i][half ^ 1][row];
            if (half == 0 && q < T) lse[((long long)b * H + h) * T + q] = m * scale + __logf(sum);
            // O = P V / rowsum: this warp stores 32 of the 64 output columns
            mbar_wait(&o_full[i], ph);
            tc_fence_after();
            const float inv = 1.f / sum;
            bf16* orow = out + ((long long)b * T + q) * (H * HD) + h * HD + half * 32;
            tmem_ld32(taddr + half * 32, ra);
            tmem_ld_wait();
            if (q < T) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(ra[g * 8 + j]) * inv;
                    st_global_v4(orow + g * 8, pack8(v));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int B_Q = 0, B_K = KV_BYTES, B_V = 2 * KV_BYTES, B_DO = 3 * KV_BYTES, B_PT = 4 * KV_BYTES,
              B_DST = B_PT + 2 * TILE128_BYTES, B_END = B_DST + 2 * TILE128_BYTES;
constexpr int BWD_SMEM = B_END + 1024;
// TMEM columns
constexpr int C_ST = 0, C_DPT = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
                bf16* __restrict__ dqkv, int B, int T, int H, float scale) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ops_full, ops_empty, st_full, tile_full, tile_free, acc_full, acc_empty, dq_full, dq_empty;
    __shared__ uint32_t tmem_slot;
    __shared__ float lse2_s[KV_ROWS], dsum_s[KV_ROWS];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NK = (T + 15) & ~15;
    const int items = B * H;
    const int D3 = 3 * H * HD;

    if (warp_idx == 0 && lane == 0) { prefetch_tmap(&map_qkv); prefetch_tmap(&map_do); }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(&ops_full, 1); mbar_init(&ops_empty, 1); mbar_init(&st_full, 1); mbar_init(&tile_full, 16);
        mbar_init(&tile_free, 1); mbar_init(&acc_full, 1); mbar_init(&acc_empty, 16); mbar_init(&dq_full, 1);
        mbar_init(&dq_empty, 16);
        fence_mbar_init();
    }
    if (warp_idx == 2) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp_idx == 0) {
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const int b = w / H, h = w - b * H;
                const int row0 = b * T;
                mbar_wait(&ops_empty, (it & 1) ^ 1);
                mbar_expect_tx(&ops_full, 4 * KV_BYTES);
                tma_load_2d(&map_qkv, &ops_full, smem + B_Q, h * HD, row0);
                tma_load_2d(&map_qkv, &ops_full, smem + B_K, (H + h) * HD, row0);
                tma_load_2d(&map_qkv, &ops_full, smem + B_V, (2 * H + h) * HD, row0);
                tma_load_2d(&map_do, &ops_full, smem + B_DO, h * HD, row0);
                if (w + (int)gridDim.x < items) {        // next head's operands -> L2 while this one is computed
                    const int wn = w + gridDim.x, bn = wn / H, hn = wn - bn * H;
                    tma_prefetch_2d(&map_qkv, hn * HD, bn * T);
                    tma_prefetch_2d(&map_qkv, (H + hn) * HD, bn * T);
                    tma_prefetch_2d(&map_qkv, (2 * H + hn) * HD, bn * T);
                    tma_prefetch_2d(&map_do, hn * HD, bn * T);
                }
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            const uint32_t sQ = smem_u32(smem + B_Q), sK = smem_u32(smem + B_K), sV = smem_u32(smem + B_V),
                           sDO = smem_u32(smem + B_DO), sPT = smem_u32(smem + B_PT), sDST = smem_u32(smem + B_DST);
            const uint32_t idesc_acc = make_idesc(128, HD, 0, 1);      // dV, dK: A K-major, B MN-major
            const uint32_t idesc_dq = make_idesc(128, HD, 1, 1);       // dQ: A (dS^T tile) MN-major, B MN-major
            int it = 0;
            uint32_t g = 0;        // step counter (phase of st_full / tile_full / tile_free)
            uint32_t ktg = 0;      // key-tile counter (phase of acc_full / acc_empty)
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                mbar_wait(&ops_full, it & 1);
                mbar_wait(&dq_empty, (it & 1) ^ 1);      // previous head's dQ has been read out of TMEM
                tc_fence_after();
                for (int kt = 0; kt < 2; ++kt, ++ktg) {
                    const int key_steps = (kt == 0 ? 128 : NK - 128) / 16;
                    for (int qc = 0; qc < 2; ++qc, ++g) {
                        const int nq = qc == 0 ? 128 : NK - 128;
                        const int q_steps = nq / 16;
                        const uint32_t idesc_s = make_idesc(128, nq, 0, 0);
                        // S^T = K_kt Q_qc^T and dP^T = V_kt dO_qc^T (the previous step's tiles were consumed in order)
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            tc_mma_f16(tmem + C_ST, make_desc(sK + kt * TILE128_BYTES + k * 32, 16, 1024),
                                       make_desc(sQ + qc * TILE128_BYTES + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            tc_mma_f16(tmem + C_DPT, make_desc(sV + kt * TILE128_BYTES + k * 32, 16, 1024),
                                       make_desc(sDO + qc * TILE128_BYTES + k * 32, 16, 1024), idesc_s, k > 0);
                        tc_commit(&st_full);
                        mbar_wait(&tile_full, g & 1);    // P^T and dS^T tiles are in shared memory
                        tc_fence_after();
                        if (qc == 0) {
                            mbar_wait(&acc_empty, (ktg & 1) ^ 1);   // previous key tile's dV/dK left TMEM
                            tc_fence_after();
                        }
                        for (int k = 0; k < q_steps; ++k) {
                            const uint64_t bdo = make_desc(sDO + qc * TILE128_BYTES + k * 2048, 8192, 1024);
                            const uint64_t bq = make_desc(sQ + qc * TILE128_BYTES + k * 2048, 8192, 1024);
                            const uint32_t aoff = (k >> 2) * TILE128_BYTES + (k & 3) * 32;
                            tc_mma_f16(tmem + C_DV, make_desc(sPT + aoff, 16, 1024), bdo, idesc_acc, (qc > 0 || k > 0));
                            tc_mma_f16(tmem + C_DK, make_desc(sDST + aoff, 16, 1024), bq, idesc_acc, (qc > 0 || k > 0));
                        }
                        for (int k = 0; k < key_steps; ++k)
                            tc_mma_f16(tmem + C_DQ + qc * HD, make_desc(sDST + k * 2048, TILE128_BYTES, 1024),
                                       make_desc(sK + kt * TILE128_BYTES + k * 2048, 8192, 1024), idesc_dq,
                                       (kt > 0 || k > 0));
                        tc_commit(&tile_free);
                    }
                    tc_commit(&acc_full);
                }
                tc_commit(&dq_full);
                tc_commit(&ops_empty);
            }
        }
    } else if (warp_idx >= 4) {
        // 16 warps: warp (quarter, cg) owns TMEM lanes 32*quarter..+31 and the 32-column chunk cg of each step.
        const int ew = warp_idx - 4;
        const int cg = ew >> 2;
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;
        const int ct = threadIdx.x - 128;                // 0..511 among the compute threads
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = scale * LOG2E;
        const uint32_t sPT = smem_u32(smem + B_PT) + row * 128;
        const uint32_t sDST = smem_u32(smem + B_DST) + row * 128;
        int it = 0;
        uint32_t g = 0, ktg = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const int b = w / H, h = w - b * H;
            // D[q] = sum_d dO[q,d] O[q,d] and lse[q] * log2(e) for this head (per-column constants of the steps below):
            // one thread per query row, 2 x 128 bytes each, all loads in flight at once
            named_bar_sync(1, 512);                      // everyone is done with the previous head's constants
            if (ct < NK) {
                float d = 0.f, l2 = 0.f;
                if (ct < T) {
                    const long long off = ((long long)b * T + ct) * (H * HD) + h * HD;
                    bf16x8 vo[8], vg[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { vo[k] = ld_global_v4(out + off + k * 8); vg[k] = ld_global_v4(dout + off + k * 8); }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float fo[8], fg[8];
                        unpack8(vo[k], fo);
                        unpack8(vg[k], fg);
#pragma unroll
                        for (int j = 0; j < 8; ++j) d += fo[j] * fg[j];
                    }
                    l2 = lse[((long long)b * H + h) * T + ct] * LOG2E;
                }
                dsum_s[ct] = d;
                lse2_s[ct] = l2;
            }
            named_bar_sync(1, 512);
            for (int kt = 0; kt < 2; ++kt, ++ktg) {
                const int key = kt * 128 + row;
                const bool key_ok = key < T;
                for (int qc = 0; qc < 2; ++qc, ++g) {
                    const int nq = qc == 0 ? 128 : NK - 128;
                    mbar_wait(&st_full, g & 1);
                    tc_fence_after();
                    if (cg * 32 < nq) {
                        uint32_t s[32], dp[32];
                        tmem_ld32(tlane + C_ST + cg * 32, s);
                        tmem_ld32(tlane + C_DPT + cg * 32, dp);
                        tmem_ld_wait();
                        mbar_wait(&tile_free, (g & 1) ^ 1);   // the previous step's MMAs have read the smem tiles
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            const int col0 = cg * 32 + gq * 8;
                            if (col0 < nq) {             // warp-uniform
                                float p[8], ds[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int q = qc * 128 + col0 + j;
                                    const float e = ex2_approx(__uint_as_float(s[gq * 8 + j]) * sl2 - lse2_s[q]);
                                    const bool ok = key_ok && (q < T);
                                    p[j] = ok ? e : 0.f;
                                    ds[j] = ok ? e * (__uint_as_float(dp[gq * 8 + j]) - dsum_s[q]) * scale : 0.f;
                                }
                                const uint32_t off = (col0 >> 6) * TILE128_BYTES + ((((col0 & 63) >> 3) ^ (row & 7)) << 4);
                                st_shared_v4(sPT + off, pack8(p));
                                st_shared_v4(sDST + off, pack8(ds));
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tile_full);
                }
                // dV (column groups 0,1) and dK (2,3) of this key tile: one 32-column chunk per warp
                mbar_wait(&acc_full, ktg & 1);
                tc_fence_after();
                {
                    const int which = cg >> 1, half = cg & 1;      // which: 0 = dV, 1 = dK
                    bf16* drow = dqkv + ((long long)b * T + key) * D3 + (which == 0 ? 2 * H : H) * HD + h * HD + half * 32;
                    uint32_t r[32];
                    tmem_ld32(tlane + (which == 0 ? C_DV : C_DK) + half * 32, r);
                    tmem_ld_wait();
                    if (key_ok) {
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
                            st_global_v4(drow + gq * 8, pack8(v));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty);
            }
            // dQ: column group (qt, half) stores 32 columns of query tile qt
            mbar_wait(&dq_full, it & 1);
            tc_fence_after();
            {
                const int qt = cg >> 1, half = cg & 1;
                const int q = qt * 128 + row;
                bf16* drow = dqkv + ((long long)b * T + q) * D3 + h * HD + half * 32;
                uint32_t r[32];
                tmem_ld32(tlane + C_DQ + qt * HD + half * 32, r);
                tmem_ld_wait();
                if (q < T) {
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
                        st_global_v4(drow + gq * 8, pack8(v));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}


// ------------------------------------------------------------------------------------------------ backward, pipelined
// Same mathematics as attn_bwd_kernel, cut into 2 key tiles x 4 query chunks of 64 (64, 64, 64, 16) so that the tensor
// core and the exp / dS warps overlap: S^T / dP^T of chunk c+1 are issued before the warps have finished chunk c (two
// TMEM buffers of 2 x 64 columns), the P^T / dS^T operand tiles are double-buffered in shared memory, dV / dK accumulate
// per chunk and dQ per chunk PAIR (M = 128 queries, the two dS^T buffers read as one MN-major operand).
constexpr int P_PT0 = 4 * KV_BYTES, P_DST0 = P_PT0 + 2 * TILE128_BYTES;      // PT[2] then DST[2], 16 KB each
constexpr int P_SD = 0;                                                      // TMEM: buffer b -> S^T at b*128, dP^T at b*128+64

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_pipe_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                     const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
                     bf16* __restrict__ dqkv, int B, int T, int H, float scale) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ops_full, ops_empty, sd_full[2], tile_full[2], tile_free[2], acc_full, acc_empty, dq_full, dq_empty;
    __shared__ uint64_t d_full[2], d_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float lse2_s[2][KV_ROWS], dsum_s[2][KV_ROWS];     // per-query constants, double-buffered across heads
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NK = (T + 15) & ~15;
    const int items = B * H;
    const int D3 = 3 * H * HD;

    if (warp_idx == 0 && lane == 0) { prefetch_tmap(&map_qkv); prefetch_tmap(&map_do); }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(&ops_full, 1); mbar_init(&ops_empty, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&sd_full[b], 1); mbar_init(&tile_full[b], 16); mbar_init(&tile_free[b], 1); }
        mbar_init(&acc_full, 1); mbar_init(&acc_empty, 16); mbar_init(&dq_full, 1); mbar_init(&dq_empty, 16);
        for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&d_empty[b], 16); }
        fence_mbar_init();
    }
    if (warp_idx == 2) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp_idx == 0) {
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const int b = w / H, h = w - b * H;
                const int row0 = b * T;
                mbar_wait(&ops_empty, (it & 1) ^ 1);
                mbar_expect_tx(&ops_full, 4 * KV_BYTES);
                tma_load_2d(&map_qkv, &ops_full, smem + B_Q, h * HD, row0);
                tma_load_2d(&map_qkv, &ops_full, smem + B_K, (H + h) * HD, row0);
                tma_load_2d(&map_qkv, &ops_full, smem + B_V, (2 * H + h) * HD, row0);
                tma_load_2d(&map_do, &ops_full, smem + B_DO, h * HD, row0);
                if (w + (int)gridDim.x < items) {
                    const int wn = w + gridDim.x, bn = wn / H, hn = wn - bn * H;
                    tma_prefetch_2d(&map_qkv, hn * HD, bn * T);
                    tma_prefetch_2d(&map_qkv, (H + hn) * HD, bn * T);
                    tma_prefetch_2d(&map_qkv, (2 * H + hn) * HD, bn * T);
                    tma_prefetch_2d(&map_do, hn * HD, bn * T);
                }
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            const uint32_t sQ = smem_u32(smem + B_Q), sK = smem_u32(smem + B_K), sV = smem_u32(smem + B_V),
                           sDO = smem_u32(smem + B_DO), sPT = smem_u32(smem + P_PT0), sDST = smem_u32(smem + P_DST0);
            const uint32_t idesc_acc = make_idesc(128, HD, 0, 1);      // dV, dK: A K-major, B MN-major
            const uint32_t idesc_dq = make_idesc(128, HD, 1, 1);       // dQ: A (dS^T buffers) MN-major, B MN-major
            const uint32_t idesc_s64 = make_idesc(128, 64, 0, 0), idesc_s16 = make_idesc(128, 16, 0, 0);
            // S^T = K_kt Q_qc^T and dP^T = V_kt dO_qc^T of chunk (kt, qc) into TMEM buffer qc & 1
            auto issue_sd = [&](int kt, int qc) {
                const uint32_t tb = tmem + P_SD + (qc & 1) * 128;
                const uint32_t id = qc == 3 ? idesc_s16 : idesc_s64;
                // the two accumulation chains are interleaved: successive MMAs into one accumulator serialise on the
                // tensor pipe's latency, which is what bounds these small (N <= 64) instructions
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    tc_mma_f16(tb, make_desc(sK + kt * TILE128_BYTES + k * 32, 16, 1024),
                               make_desc(sQ + qc * 8192 + k * 32, 16, 1024), id, k > 0);
                    tc_mma_f16(tb + 64, make_desc(sV + kt * TILE128_BYTES + k * 32, 16, 1024),
                               make_desc(sDO + qc * 8192 + k * 32, 16, 1024), id, k > 0);
                }
                tc_commit(&sd_full[qc & 1]);
            };
            int it = 0;
            uint32_t use[2] = {0, 0};      // completed uses of each buffer (phase of tile_full)
            uint32_t ktg = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                mbar_wait(&ops_full, it & 1);
                mbar_wait(&dq_empty, (it & 1) ^ 1);
                tc_fence_after();
                issue_sd(0, 0);
                for (int kt = 0; kt < 2; ++kt, ++ktg) {
                    const int key_steps = (kt == 0 ? 128 : NK - 128) / 16;
                    for (int qc = 0; qc < 4; ++qc) {
                        const int b = qc & 1;
                        // next chunk's S^T / dP^T: its TMEM buffer was read out before tile_full of the step two back
                        if (qc < 3) issue_sd(kt, qc + 1);
                        else if (kt == 0) issue_sd(1, 0);
                        mbar_wait(&tile_full[b], use[b] & 1);
                        ++use[b];
                        tc_fence_after();
                        if (qc == 0) {
                            mbar_wait(&acc_empty, (ktg & 1) ^ 1);
                            tc_fence_after();
                        }
                        const int q_steps = qc == 3 ? 1 : 4;
                        // dV, dK (K = this chunk's queries) and, on the second chunk of a pair, dQ of the query tile the
                        // pair covers (the two dS^T buffers as one MN-major operand, K = this tile's keys): three
                        // independent accumulation chains, issued round-robin
                        const int dq_steps = b == 1 ? key_steps : 0;
                        for (int k = 0; k < (q_steps > dq_steps ? q_steps : dq_steps); ++k) {
                            if (k < q_steps) {
                                const uint64_t bdo = make_desc(sDO + qc * 8192 + k * 2048, 8192, 1024);
                                const uint64_t bq = make_desc(sQ + qc * 8192 + k * 2048, 8192, 1024);
                                tc_mma_f16(tmem + C_DV, make_desc(sPT + b * TILE128_BYTES + k * 32, 16, 1024), bdo, idesc_acc,
                                           (qc > 0 || k > 0));
                                tc_mma_f16(tmem + C_DK, make_desc(sDST + b * TILE128_BYTES + k * 32, 16, 1024), bq, idesc_acc,
                                           (qc > 0 || k > 0));
                            }
                            if (k < dq_steps)
                                tc_mma_f16(tmem + C_DQ + (qc >> 1) * HD, make_desc(sDST + k * 2048, TILE128_BYTES, 1024),
                                           make_desc(sK + kt * TILE128_BYTES + k * 2048, 8192, 1024), idesc_dq, (kt > 0 || k > 0));
                        }
                        if (b == 1) {
                            tc_commit(&tile_free[0]);
                            tc_commit(&tile_free[1]);
                        }
                    }
                    tc_commit(&acc_full);
                }
                tc_commit(&dq_full);
                tc_commit(&ops_empty);
            }
        }
    } else if (warp_idx == 3) {
        // D[q] = sum_d dO[q,d] O[q,d] and lse[q] * log2(e) of the NEXT head while the other warps work on the current one
        int it = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const int b = w / H, h = w - b * H;
            const int buf = it & 1;
            mbar_wait(&d_empty[buf], ((it >> 1) & 1) ^ 1);
            for (int q = lane; q < KV_ROWS; q += 32) {
                float d = 0.f, l2 = 0.f;
                if (q < T) {
                    const long long off = ((long long)b * T + q) * (H * HD) + h * HD;
                    bf16x8 vo[8], vg[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { vo[k] = ld_global_v4(out + off + k * 8); vg[k] = ld_global_v4(dout + off + k * 8); }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float fo[8], fg[8];
                        unpack8(vo[k], fo);
                        unpack8(vg[k], fg);
#pragma unroll
                        for (int j = 0; j < 8; ++j) d += fo[j] * fg[j];
                    }
                    l2 = lse[((long long)b * H + h) * T + q] * LOG2E;
                }
                dsum_s[buf][q] = d;
                lse2_s[buf][q] = l2;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_full[buf]);
        }
    } else if (warp_idx >= 4) {
        // 16 warps: warp (quarter, cg) owns TMEM lanes 32*quarter..+31 and 16 of the 64 query columns of each chunk
        const int ew = warp_idx - 4;
        const int cg = ew >> 2;
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        const float sl2 = scale * LOG2E;
        const uint32_t sPT = smem_u32(smem + P_PT0) + row * 128;
        const uint32_t sDST = smem_u32(smem + P_DST0) + row * 128;
        int it = 0;
        uint32_t use[2] = {0, 0};
        uint32_t ktg = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const int b = w / H, h = w - b * H;
            const int dbuf = it & 1;
            mbar_wait(&d_full[dbuf], (it >> 1) & 1);         // this head's per-query constants (written by warp 3)
            const float* lse2c = lse2_s[dbuf];
            const float* dsumc = dsum_s[dbuf];
            for (int kt = 0; kt < 2; ++kt, ++ktg) {
                const int key = kt * 128 + row;
                const bool key_ok = key < T;
                for (int qc = 0; qc < 4; ++qc) {
                    const int bf = qc & 1;
                    mbar_wait(&sd_full[bf], use[bf] & 1);
                    tc_fence_after();
                    const bool active = qc < 3 || cg == 0;        // the last chunk holds 16 queries
                    float p[16], ds[16];
                    if (active) {
                        uint32_t s[16], dp[16];
                        tmem_ld16(tlane + P_SD + bf * 128 + cg * 16, s);
                        tmem_ld16(tlane + P_SD + bf * 128 + 64 + cg * 16, dp);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int q = qc * 64 + cg * 16 + j;
                            const float e = ex2_approx(__uint_as_float(s[j]) * sl2 - lse2c[q]);
                            const bool ok = key_ok && (q < T);
                            p[j] = ok ? e : 0.f;
                            ds[j] = ok ? e * (__uint_as_float(dp[j]) - dsumc[q]) * scale : 0.f;
                        }
                    }
                    // the MMAs of the pair that last used this smem buffer have finished reading it
                    mbar_wait(&tile_free[bf], (use[bf] & 1) ^ 1);
                    ++use[bf];
                    if (active) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t off = bf * TILE128_BYTES + (((cg * 2 + j) ^ (row & 7)) << 4);
                            st_shared_v4(sPT + off, pack8(p + j * 8));
                            st_shared_v4(sDST + off, pack8(ds + j * 8));
                        }
                    }
                    fence_proxy_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tile_full[bf]);
                }
                mbar_wait(&acc_full, ktg & 1);
                tc_fence_after();
                {
                    const int which = cg >> 1, half = cg & 1;      // which: 0 = dV, 1 = dK
                    bf16* drow = dqkv + ((long long)b * T + key) * D3 + (which == 0 ? 2 * H : H) * HD + h * HD + half * 32;
                    uint32_t r[32];
                    tmem_ld32(tlane + (which == 0 ? C_DV : C_DK) + half * 32, r);
                    tmem_ld_wait();
                    if (key_ok) {
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
                            st_global_v4(drow + gq * 8, pack8(v));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[dbuf]);
            mbar_wait(&dq_full, it & 1);
            tc_fence_after();
            {
                const int qt = cg >> 1, half = cg & 1;
                const int q = qt * 128 + row;
                bf16* drow = dqkv + ((long long)b * T + q) * D3 + h * HD + half * 32;
                uint32_t r[32];
                tmem_ld32(tlane + C_DQ + qt * HD + half * 32, r);
                tmem_ld_wait();
                if (q < T) {
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
                        st_global_v4(drow + gq * 8, pack8(v));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

bool g_attr_fwd = false, g_attr_bwd = false;
// DFU_ATTN_BWD=serial selects the unpipelined backward kernel
const bool g_bwd_pipe = []() { const char* e = getenv("DFU_ATTN_BWD"); return !(e && e[0] == 's'); }();

int check_shape(const char* what, int B, int T, int H) {
    DFU_CHECK_ARG(B > 0 && H > 0 && T > 128 && T <= KV_ROWS, "%s: supports 128 < T <= %d tokens (got T=%d, B=%d, H=%d)", what,
                  KV_ROWS, T, B, H);
    return DFU_OK;
}

}  // namespace

extern "C" int dfu_attention_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t T, int32_t H, float scale,
                                 void* stream) {
    DFU_CHECK_ARG(qkv && out && lse, "dfu_attention_fwd: null pointer");
    if (int rc = check_shape("dfu_attention_fwd", B, T, H)) return rc;
    DFU_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "dfu_attention_fwd: unaligned buffers");
    CUtensorMap map_q, map_kv;
    const long long rows = (long long)B * T, cols = 3LL * H * HD;
    if (int rc = make_map_2d(&map_q, qkv, cols, rows, cols, HD, 128)) return rc;
    if (int rc = make_map_2d(&map_kv, qkv, cols, rows, cols, HD, KV_ROWS)) return rc;
    if (!g_attr_fwd) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
        if (e != cudaSuccess) {
            dfu_set_error("dfu_attention_fwd: cannot reserve %d B shared memory: %s", FWD_SMEM, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return DFU_ERR_CUDA;
        }
        g_attr_fwd = true;
    }
    const int items = B * H;
    const int grid = items < dfu_num_sms() ? items : dfu_num_sms();
    dfu_launch(attn_fwd_kernel, grid, THREADS, FWD_SMEM, (cudaStream_t)stream, map_q, map_kv, (bf16*)out, lse, B, T, H, scale);
    DFU_CHECK_LAUNCH("attn_fwd_kernel");
    return DFU_OK;
}

extern "C" int dfu_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int32_t B,
                                 int32_t T, int32_t H, float scale, void* stream) {
    DFU_CHECK_ARG(qkv && out && dout && lse && dqkv, "dfu_attention_bwd: null pointer");
    if (int rc = check_shape("dfu_attention_bwd", B, T, H)) return rc;
    DFU_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)dout & 15) == 0 && ((uintptr_t)dqkv & 15) == 0 &&
                      ((uintptr_t)out & 3) == 0,
                  "dfu_attention_bwd: unaligned buffers");
    CUtensorMap map_qkv, map_do;
    const long long rows = (long long)B * T;
    if (int rc = make_map_2d(&map_qkv, qkv, 3LL * H * HD, rows, 3LL * H * HD, HD, KV_ROWS)) return rc;
    if (int rc = make_map_2d(&map_do, dout, (long long)H * HD, rows, (long long)H * HD, HD, KV_ROWS)) return rc;
    if (!g_attr_bwd) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
        if (e != cudaSuccess) {
            dfu_set_error("dfu_attention_bwd: cannot reserve %d B shared memory: %s", BWD_SMEM, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return DFU_ERR_CUDA;
        }
        g_attr_bwd = true;
    }
    const int items = B * H;
    const int grid = items < dfu_num_sms() ? items : dfu_num_sms();
    if (g_bwd_pipe)
        dfu_launch(attn_bwd_pipe_kernel, grid, THREADS, BWD_SMEM, (cudaStream_t)stream, map_qkv, map_do, (const bf16*)out,
                                                                                (const bf16*)dout, lse, (bf16*)dqkv, B, T, H, scale);
    else
        dfu_launch(attn_bwd_kernel, grid, THREADS, BWD_SMEM, (cudaStream_t)stream, map_qkv, map_do, (const bf16*)out, (const bf16*)dout,
                                                                           lse, (bf16*)dqkv, B, T, H, scale);
    DFU_CHECK_LAUNCH("attn_bwd_kernel");
    return DFU_OK;
}
