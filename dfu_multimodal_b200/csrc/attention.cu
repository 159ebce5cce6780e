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
// Warp roles (640 threads): warp 0 TMA producer, warp 1 MMA issuer (the whole warp walks the schedule, one elected lane
// issues), warp 2 TMEM allocator, warps 4-19 softmax / epilogue (warp w owns TMEM lanes 32*(w%4)..+31; the four warps of a
// lane quarter split the columns).  Everything that leaves the chip in rows -- O, dQ, dK, dV -- is staged in a shared-memory
// tile that is free at that moment and stored by TMA through a 3-D view that clips the rows >= T: a thread owns one row, so
// direct stores touch 32 different 128-byte lines per instruction.  How the time of a head splits over the phases of
// each kernel is measured by tools/gpu_attn_trace.py (a second build with -DDFU_ATTN_TRACE); DESIGN.md 5.2 has the numbers.
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

// Phase tracing for tools/gpu_attn_trace.py (a separate library built with -DDFU_ATTN_TRACE; compiled out otherwise):
// CTA 0 records clock64() at the phase boundaries of its MMA thread and of one thread per softmax group.
#ifdef DFU_ATTN_TRACE
__device__ long long g_attn_trace[4][16][12];
#define ATTN_TR(role, it, ph)                                                                   \
    do {                                                                                        \
        if (blockIdx.x == 0 && (it) < 16 && (threadIdx.x & 31) == 0) g_attn_trace[role][it][ph] = clock64(); \
    } while (0)
// accumulating variant for the backward kernel: time since the previous mark goes to category `cat` of role 1
#define ATTN_ACC(cat)                                                                           \
    do {                                                                                        \
        const long long now_ = clock64();                                                       \
        if (trb) tr_acc[cat] += now_ - tprev;     /* shared memory: a global read-modify-write would stall the warp */ \
        tprev = now_;                                                                           \
    } while (0)
#else
#define ATTN_TR(role, it, ph) ((void)0)
#define ATTN_ACC(cat) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------ forward
constexpr int F_Q0 = 0, F_Q1 = TILE128_BYTES, F_K = 2 * TILE128_BYTES, F_V = F_K + KV_BYTES, F_P0 = F_V + KV_BYTES,
              F_P1 = F_P0 + 4 * TILE128_BYTES, F_END = F_P1 + 4 * TILE128_BYTES;
constexpr int FWD_SMEM = F_END + 1024;

__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                const __grid_constant__ CUtensorMap map_o, float* __restrict__ lse, int B, int T, int H, float scale) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t qk_full, qk_empty, v_full, v_empty, s_full[2], p_full[2], o_full[2], o_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float pmax_s[2][2][128], psum_s[2][2][128];   // partial row max / sum exchanged by the two column halves
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NK = (T + 15) & ~15;            // key columns of S (multiple of 16)
    const int items = B * H;

    if (warp_idx == 0 && lane == 0) { prefetch_tmap(&map_q); prefetch_tmap(&map_kv); prefetch_tmap(&map_o); }
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
        // The whole warp walks the loop and waits; one elected lane issues the MMAs and their commits.
        const uint32_t idesc_s = make_idesc(128, NK, 0, 0);
        const uint32_t idesc_o = make_idesc(128, HD, 0, 1);
        const uint32_t sQ = smem_u32(smem + F_Q0), sK = smem_u32(smem + F_K), sV = smem_u32(smem + F_V),
                       sP = smem_u32(smem + F_P0);
        const int nks = NK / 16;
        const uint64_t dQ0 = make_desc(sQ, 16, 1024), dK0 = make_desc(sK, 16, 1024);
        const uint64_t dP0 = make_desc(sP, 16, 1024), dV0 = make_desc(sV, 8192, 1024);
        auto issue_s = [&](int i) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc_mma_f16(tmem + i * 256, dQ0 + (uint64_t)((i * TILE128_BYTES + k * 32) >> 4), dK0 + (uint64_t)((k * 32) >> 4),
                               idesc_s, k > 0);
                tc_commit(&s_full[i]);
                if (i == 1) tc_commit(&qk_empty);        // Q and K may be overwritten once both S tiles are done
            }
            __syncwarp();
        };
        auto issue_o = [&](int i) {
            if (elect_one()) {
                // 13 MMAs of 32 tensor clocks each: descriptors = base + compile-time offset, nothing else per issue
#pragma unroll
                for (int k = 0; k < KV_ROWS / 16; ++k)
                    if (k < nks)
                        tc_mma_f16(tmem + i * 256,
                                   dP0 + (uint64_t)((i * 4 * TILE128_BYTES + (k >> 2) * TILE128_BYTES + (k & 3) * 32) >> 4),
                                   dV0 + (uint64_t)((k * 2048) >> 4), idesc_o, k > 0);
                tc_commit(&o_full[i]);
                if (i == 1) tc_commit(&v_empty);         // V may be overwritten once both O tiles are done
            }
            __syncwarp();
        };
        // The two query tiles run half a head apart: S0(h), O1(h-1), S1(h), O0(h), ...  While one softmax group is in its
        // exp pass (SFU-bound) the other is waiting for an MMA, reading O out or reducing its row maxima, instead of both
        // groups hitting the same phase together and leaving the SFU idle in between.
        int it = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            mbar_wait(&qk_full, ph);
            mbar_wait(&o_empty[0], ph ^ 1);              // the previous head's O_0 has been read out of TMEM
            tc_fence_after();
            ATTN_TR(0, it, 0);
            issue_s(0);
            ATTN_TR(0, it, 1);
            if (it > 0) {
                mbar_wait(&p_full[1], ph ^ 1);           // previous head: softmax warps wrote P_1 (and finished reading S_1)
                tc_fence_after();
                ATTN_TR(0, it, 2);
                issue_o(1);
            }
            ATTN_TR(0, it, 3);
            mbar_wait(&o_empty[1], ph ^ 1);
            tc_fence_after();
            ATTN_TR(0, it, 4);
            issue_s(1);
            ATTN_TR(0, it, 5);
            mbar_wait(&v_full, ph);
            mbar_wait(&p_full[0], ph);
            tc_fence_after();
            ATTN_TR(0, it, 6);
            issue_o(0);
            ATTN_TR(0, it, 7);
        }
        if (it > 0) {                                    // drain: O_1 of the last head
            mbar_wait(&p_full[1], (it - 1) & 1);
            tc_fence_after();
            issue_o(1);
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
        const bool storer = (ew & 7) == 0 && lane == 0;  // issues the group's TMA store of O
        int it = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            const int b = w / H, h = w - b * H;
            const bool tr = (ew & 7) == 0 && lane == 0;  // one tracing thread per group
            if (tr) ATTN_TR(1 + i, it, 0);
            mbar_wait(&s_full[i], ph);
            tc_fence_after();
            if (tr) ATTN_TR(1 + i, it, 1);
            uint32_t ra[32], rb[32];
            // pass 1: row maximum of the raw scores (the next chunk's TMEM load is in flight while this one is reduced;
            // four independent chains so that the reduction is not one 104-deep dependency)
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            // tcgen05.wait::ld waits for every outstanding load, so "load the next chunk while this one is reduced" leaves the
            // ~300-clock TMEM latency exposed once per chunk (the reduction is 32 FMNMX).  Two chunks per round trip instead.
#pragma unroll
            for (int cc = 0; cc < 4; cc += 2) {
                if (cc < my_chunks) {
                    tmem_ld32(taddr + (half + 2 * cc) * 32, ra);
                    if (cc + 1 < my_chunks) tmem_ld32(taddr + (half + 2 * cc + 2) * 32, rb);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (cc + u < my_chunks) {
                            const int c = half + 2 * (cc + u);
                            const uint32_t* cur = u ? rb : ra;
                            if (c * 32 + 32 <= T) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(cur[j]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (c * 32 + j < T) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(cur[j]));
                            }
                        }
                    }
                }
            }
            float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            pmax_s[i][half][row] = m;
            if (tr) ATTN_TR(1 + i, it, 2);
            if (storer) bulk_wait_read0();               // the previous head's O tile has left the P_i buffer (see below)
            named_bar_sync(1 + i, 256);
            if (tr) ATTN_TR(1 + i, it, 3);
            m = fmaxf(m, pmax_s[i][half ^ 1][row]);
            // pass 2: p = exp(scale (s - m)), partial row sum, bf16 P into the K-major SWIZZLE_128B operand tile.  Only the
            // chunk that straddles T pays for the key mask; the others are FFMA + EX2 + FADD + half a pack per score.
            const float ms = m * sl2;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
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
                        for (int g = 0; g < 4; ++g) {
                            const int col0 = c * 32 + g * 8;
                            float p[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float x = __uint_as_float(cur[g * 8 + j]) * sl2 - ms;
#if defined(DFU_ATTN_EXP_NONE)      // timing experiment of tools/gpu_attn_trace.py (wrong results): the pass without the SFU
                                p[j] = x * 0.5f;
#else
                                p[j] = ex2_approx(x);
#endif
                                s4[j & 3] += p[j];
                            }
                            const int atom = col0 >> 6, cj = (col0 & 63) >> 3;
#ifndef DFU_ATTN_NO_PSTORE
                            st_shared_v4(sP + atom * TILE128_BYTES + ((cj ^ (row & 7)) << 4), pack8(p));
#endif
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int col0 = c * 32 + g * 8;
                            if (col0 < NK) {             // warp-uniform
                                float p[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float e = ex2_approx(__uint_as_float(cur[g * 8 + j]) * sl2 - ms);
                                    p[j] = col0 + j < T ? e : 0.f;
                                    s4[j & 3] += p[j];
                                }
                                const int atom = col0 >> 6, cj = (col0 & 63) >> 3;
                                st_shared_v4(sP + atom * TILE128_BYTES + ((cj ^ (row & 7)) << 4), pack8(p));
                            }
                        }
                    }
                }
            }
            float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            psum_s[i][half][row] = sum;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[i]);
            if (tr) ATTN_TR(1 + i, it, 4);
            named_bar_sync(1 + i, 256);
            if (tr) ATTN_TR(1 + i, it, 5);
            sum += psum_s[i][half ^ 1][row];
            if (half == 0 && q < T) lse[((long long)b * H + h) * T + q] = m * scale + __logf(sum);
            // O = P V / rowsum.  A thread owns one row, so direct global stores would touch 32 different 128-byte lines per
            // instruction (2 048 L1 wavefronts per head -- measured 2.4 k clocks); instead the bf16 rows are staged in the
            // first atom of P_i (free once o_full fires: the MMA has read P) and leave through one TMA store per tile, which
            // also clips the rows >= T of the second tile.
            mbar_wait(&o_full[i], ph);
            tc_fence_after();
            if (tr) ATTN_TR(1 + i, it, 6);
            const float inv = 1.f / sum;
            tmem_ld32(taddr + half * 32, ra);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[i]);     // TMEM columns of tile i are free for the next head's S
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(ra[g * 8 + j]) * inv;
                st_shared_v4(sP + (((half * 4 + g) ^ (row & 7)) << 4), pack8(v));
            }
            fence_proxy_async_smem();
            named_bar_sync(1 + i, 256);
            if (storer) {
                tma_store_3d(&map_o, smem + (i ? F_P1 : F_P0), h * HD, i * 128, b);
                bulk_commit();
            }
            if (tr) ATTN_TR(1 + i, it, 7);
        }
        if (storer) bulk_wait0();                        // the last O tile is in global memory before the CTA retires
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
constexpr int P_O = P_DST0 + 2 * TILE128_BYTES;                              // O rows of the head (for D = rowsum(dO * O))
// second home of the rows 0..63 of Q / dO / O (8 KB each): odd heads of a CTA keep their first query chunk here, so that it
// can be loaded a head ahead while the previous head still uses rows 0..63 of B_Q / B_DO / P_O
constexpr int P_XQ = P_O + KV_BYTES, P_XDO = P_XQ + 64 * 128, P_XO = P_XDO + 64 * 128;
constexpr int PIPE_SMEM = P_XO + 64 * 128 + 1024;
// Gradient tiles on their way out are staged in the P^T / dS^T operand tiles: dV / dK of a key tile in PT[1] / DST[1], dQ in
// PT[0] / DST[0].  They are free from acc_full / dq_full (every MMA that read them has completed) until the exp warps write
// that buffer again -- chunk 1 of the next key tile resp. chunk 0 of the next head -- which first waits for the TMA stores
// to have read them.
constexpr int P_STAGE0 = P_PT0 + TILE128_BYTES, P_STAGE1 = P_DST0 + TILE128_BYTES;

// TMA views of the pipelined backward: loads of qkv / dO / O in boxes of 64, 144, 128 and 80 rows (the order the operands are
// needed in), the clipped 3-D store view of dqkv.
struct BwdMaps {
    CUtensorMap qkv_64, qkv_144, qkv_128, qkv_80, do_64, do_144, out_64, out_144, dqkv;
};

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_pipe_kernel(const __grid_constant__ BwdMaps maps, const float* __restrict__ lse, int B, int T, int H, float scale) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_a, full_b, full_kv0, ops_empty, kv0_empty, sd_full[2], tile_full[2], tile_free[2], acc_full, acc_empty,
        dq_full, dq_empty;
    __shared__ uint32_t tmem_slot;
    // per-query constants of the head: lse * log2(e) and D = scale * sum_d dO O
    __shared__ __align__(16) float lse2_s[KV_ROWS], dsum_s[KV_ROWS];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NK = (T + 15) & ~15;
    const int items = B * H;

    if (warp_idx == 0 && lane == 0) {
        prefetch_tmap(&maps.qkv_64); prefetch_tmap(&maps.qkv_144); prefetch_tmap(&maps.qkv_128); prefetch_tmap(&maps.qkv_80);
        prefetch_tmap(&maps.do_64); prefetch_tmap(&maps.do_144); prefetch_tmap(&maps.out_64); prefetch_tmap(&maps.out_144);
        prefetch_tmap(&maps.dqkv);
    }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(&full_a, 1); mbar_init(&full_b, 1); mbar_init(&full_kv0, 1);
        mbar_init(&ops_empty, 1); mbar_init(&kv0_empty, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&sd_full[b], 1); mbar_init(&tile_full[b], 16); mbar_init(&tile_free[b], 1); }
        mbar_init(&acc_full, 1); mbar_init(&acc_empty, 16); mbar_init(&dq_full, 1); mbar_init(&dq_empty, 16);
        fence_mbar_init();
    }
    if (warp_idx == 2) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp_idx == 0) {
        if (lane == 0) {
            // The 133 KB of a head's operands arrive in the order they are needed, on three barriers, and as early as their
            // buffers allow:
            //   kv0  K and V of key tile 0 -- dead once the tile-0 MMAs have completed (kv0_empty), so the NEXT head's are
            //        loaded half a head early, under this head's second key tile
            //   a    Q / dO / O rows 0..63: with kv0, all the first chunk needs.  Two homes (in place for even heads, P_X* for
            //        odd ones), so the next head's are loaded half a head early too
            //   b    the rest (Q / dO / O rows 64.., K and V of key tile 1), once the previous head has completed; it lands
            //        while the first chunk is being processed.  (One barrier per query chunk instead was slower: 142.7 vs
            //        139.1 us per layer -- more, smaller TMA requests and three more CTA-wide barriers.)
            // With one barrier for everything the exp warps idled ~5 k clocks of a ~22 k-clock head at every head boundary.
            auto load_kv0 = [&](int w) {
                const int b = w / H, h = w - b * H;
                mbar_expect_tx(&full_kv0, 2 * TILE128_BYTES);
                tma_load_2d(&maps.qkv_128, &full_kv0, smem + B_K, (H + h) * HD, b * T);
                tma_load_2d(&maps.qkv_128, &full_kv0, smem + B_V, (2 * H + h) * HD, b * T);
            };
            auto load_a = [&](int w, int par) {
                const int b = w / H, h = w - b * H;
                mbar_expect_tx(&full_a, 3 * 64 * 128);
                tma_load_2d(&maps.qkv_64, &full_a, smem + (par ? P_XQ : B_Q), h * HD, b * T);
                tma_load_2d(&maps.do_64, &full_a, smem + (par ? P_XDO : B_DO), h * HD, b * T);
                tma_load_2d(&maps.out_64, &full_a, smem + (par ? P_XO : P_O), h * HD, b * T);
            };
            if ((int)blockIdx.x < items) {
                load_kv0(blockIdx.x);
                load_a(blockIdx.x, 0);
            }
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const int b = w / H, h = w - b * H;
                const int row0 = b * T;
                const bool more = w + (int)gridDim.x < items;
                mbar_wait(&ops_empty, (it & 1) ^ 1);
                mbar_expect_tx(&full_b, 3 * 144 * 128 + 2 * 80 * 128);
                tma_load_2d(&maps.qkv_144, &full_b, smem + B_Q + 64 * 128, h * HD, row0 + 64);
                tma_load_2d(&maps.do_144, &full_b, smem + B_DO + 64 * 128, h * HD, row0 + 64);
                tma_load_2d(&maps.out_144, &full_b, smem + P_O + 64 * 128, h * HD, row0 + 64);
                tma_load_2d(&maps.qkv_80, &full_b, smem + B_K + TILE128_BYTES, (H + h) * HD, row0 + 128);
                tma_load_2d(&maps.qkv_80, &full_b, smem + B_V + TILE128_BYTES, (2 * H + h) * HD, row0 + 128);
                if (w + 2 * (int)gridDim.x < items) {            // the head after next -> L2
                    const int wn = w + 2 * gridDim.x, bn = wn / H, hn = wn - bn * H;
                    tma_prefetch_2d(&maps.qkv_64, hn * HD, bn * T);
                    tma_prefetch_2d(&maps.qkv_144, hn * HD, bn * T + 64);
                    tma_prefetch_2d(&maps.qkv_128, (H + hn) * HD, bn * T);
                    tma_prefetch_2d(&maps.qkv_80, (H + hn) * HD, bn * T + 128);
                    tma_prefetch_2d(&maps.qkv_128, (2 * H + hn) * HD, bn * T);
                    tma_prefetch_2d(&maps.qkv_80, (2 * H + hn) * HD, bn * T + 128);
                    tma_prefetch_2d(&maps.do_64, hn * HD, bn * T);
                    tma_prefetch_2d(&maps.do_144, hn * HD, bn * T + 64);
                    tma_prefetch_2d(&maps.out_64, hn * HD, bn * T);
                    tma_prefetch_2d(&maps.out_144, hn * HD, bn * T + 64);
                }
                if (more) {
                    // This head's tile-0 MMAs have completed: K / V rows 0..127 are dead, and every consumer has passed its
                    // wait on this head's phase of full_a / full_kv0 -- re-arming a barrier before its waiters have seen the
                    // phase complete would let two phases go by and leave them waiting for the wrong parity.
                    mbar_wait(&kv0_empty, it & 1);
                    load_kv0(w + gridDim.x);
                    load_a(w + gridDim.x, (it + 1) & 1);         // its home was last used by the head before this one
                }
            }
        }
    } else if (warp_idx == 1) {
        // The whole warp walks the (fully unrolled) schedule and waits; one elected lane issues.  The ~140 MMAs of a head
        // are 16-32 tensor clocks each, so an issue must cost a few instructions: every descriptor is a precomputed base
        // plus a compile-time offset (elect_one() in warp-uniform flow keeps them in uniform registers).
        const uint32_t sQ = smem_u32(smem + B_Q), sK = smem_u32(smem + B_K), sV = smem_u32(smem + B_V),
                       sDO = smem_u32(smem + B_DO), sPT = smem_u32(smem + P_PT0), sDST = smem_u32(smem + P_DST0);
        const uint32_t idesc_acc = make_idesc(128, HD, 0, 1);      // dV, dK: A K-major, B MN-major
        const uint32_t idesc_dq = make_idesc(128, HD, 1, 1);       // dQ: A (dS^T buffers) MN-major, B MN-major
        const uint32_t idesc_s64 = make_idesc(128, 64, 0, 0), idesc_s16 = make_idesc(128, 16, 0, 0);
        const uint64_t kK = make_desc(sK, 16, 1024), kQ = make_desc(sQ, 16, 1024), kV = make_desc(sV, 16, 1024),
                       kDO = make_desc(sDO, 16, 1024), kPT = make_desc(sPT, 16, 1024), kDST = make_desc(sDST, 16, 1024);
        const uint64_t mDO = make_desc(sDO, 8192, 1024), mQ = make_desc(sQ, 8192, 1024), mK = make_desc(sK, 8192, 1024),
                       mDST = make_desc(sDST, TILE128_BYTES, 1024);
        // the first query chunk of odd heads lives in P_XQ / P_XDO
        const uint64_t kXQ = make_desc(smem_u32(smem + P_XQ), 16, 1024), kXDO = make_desc(smem_u32(smem + P_XDO), 16, 1024),
                       mXQ = make_desc(smem_u32(smem + P_XQ), 8192, 1024), mXDO = make_desc(smem_u32(smem + P_XDO), 8192, 1024);
        uint64_t kQ0 = kQ, kDO0 = kDO, mQ0 = mQ, mDO0 = mDO;      // this head's chunk-0 bases
        // S^T = K_kt Q_qc^T and dP^T = V_kt dO_qc^T of chunk (kt, qc) into TMEM buffer qc & 1
        auto issue_sd = [&](int kt, int qc) {
            if (elect_one()) {
                const uint32_t tb = tmem + P_SD + (qc & 1) * 128;
                const uint32_t id = qc == 3 ? idesc_s16 : idesc_s64;
                // the two accumulation chains are interleaved: successive MMAs into one accumulator serialise on the
                // tensor pipe's latency, which is what bounds these small (N <= 64) instructions
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    tc_mma_f16(tb, kK + (uint64_t)((kt * TILE128_BYTES + k * 32) >> 4),
                               (qc == 0 ? kQ0 : kQ + (uint64_t)((qc * 8192) >> 4)) + (uint64_t)((k * 32) >> 4), id, k > 0);
                    tc_mma_f16(tb + 64, kV + (uint64_t)((kt * TILE128_BYTES + k * 32) >> 4),
                               (qc == 0 ? kDO0 : kDO + (uint64_t)((qc * 8192) >> 4)) + (uint64_t)((k * 32) >> 4), id, k > 0);
                }
                tc_commit(&sd_full[qc & 1]);
            }
            __syncwarp();
        };
        int it = 0;
        uint32_t use[2] = {0, 0};      // completed uses of each buffer (phase of tile_full)
        uint32_t ktg = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            kQ0 = (it & 1) ? kXQ : kQ; kDO0 = (it & 1) ? kXDO : kDO; mQ0 = (it & 1) ? mXQ : mQ; mDO0 = (it & 1) ? mXDO : mDO;
            mbar_wait(&full_kv0, it & 1);
            mbar_wait(&full_a, it & 1);
            mbar_wait(&dq_empty, (it & 1) ^ 1);
            tc_fence_after();
            issue_sd(0, 0);
            mbar_wait(&full_b, it & 1);                          // everything after the first chunk needs the rest
            tc_fence_after();
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                const int key_steps = (kt == 0 ? 128 : NK - 128) / 16;
#pragma unroll
                for (int qc = 0; qc < 4; ++qc) {
                    const int b = qc & 1;
                    // next chunk's S^T / dP^T: its TMEM buffer was read out before tile_full of the step two back
                    if (qc < 3) issue_sd(kt, qc + 1);
                    else if (kt == 0) issue_sd(1, 0);
                    mbar_wait(&tile_full[b], use[b] & 1);
                    ++use[b];
                    if (qc == 0) mbar_wait(&acc_empty, (ktg & 1) ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
                        // dV, dK (K = this chunk's queries) and, on the second chunk of a pair, dQ of the query tile the
                        // pair covers (the two dS^T buffers as one MN-major operand, K = this tile's keys): three
                        // independent accumulation chains, issued round-robin
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (k < (qc == 3 ? 1 : 4)) {
                                tc_mma_f16(tmem + C_DV, kPT + (uint64_t)((b * TILE128_BYTES + k * 32) >> 4),
                                           (qc == 0 ? mDO0 : mDO + (uint64_t)((qc * 8192) >> 4)) + (uint64_t)((k * 2048) >> 4),
                                           idesc_acc, (qc > 0 || k > 0));
                                tc_mma_f16(tmem + C_DK, kDST + (uint64_t)((b * TILE128_BYTES + k * 32) >> 4),
                                           (qc == 0 ? mQ0 : mQ + (uint64_t)((qc * 8192) >> 4)) + (uint64_t)((k * 2048) >> 4),
                                           idesc_acc, (qc > 0 || k > 0));
                            }
                            if (b == 1 && k < key_steps)
                                tc_mma_f16(tmem + C_DQ + (qc >> 1) * HD, mDST + (uint64_t)((k * 2048) >> 4),
                                           mK + (uint64_t)((kt * TILE128_BYTES + k * 2048) >> 4), idesc_dq, (kt > 0 || k > 0));
                        }
                        if (b == 1) {
                            tc_commit(&tile_free[0]);
                            tc_commit(&tile_free[1]);
                        }
                        if (qc == 3) {
                            tc_commit(&acc_full);
                            if (kt == 0) tc_commit(&kv0_empty);      // K / V rows 0..127 are dead: the next head's may land
                            if (kt == 1) {
                                tc_commit(&dq_full);
                                tc_commit(&ops_empty);
                            }
                        }
                    }
                    __syncwarp();
                }
                ++ktg;
            }
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
        const bool storer = ew == 0 && lane == 0;            // issues the TMA stores of the staged gradient tiles
        const int tid = threadIdx.x - 128;                   // 0..511 among the exp warps
        float lse_next = 0.f;                                // lse[tid] of the next head (0 for the padding rows)
        if (tid < T && (int)blockIdx.x < items)
            lse_next = __ldg(lse + (long long)blockIdx.x * T + tid);       // lse is [B*H][T] and w = b * H + h
        int it = 0;
        uint32_t use[2] = {0, 0};
        uint32_t ktg = 0;
        [[maybe_unused]] const bool trb = blockIdx.x == 0 && ew == 0 && lane == 0;      // the tracing thread (tracing builds)
#ifdef DFU_ATTN_TRACE
        __shared__ long long tr_acc[12];
        if (trb)
            for (int c = 0; c < 12; ++c) tr_acc[c] = 0;
        long long tprev = clock64();
#endif
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const int b = w / H, h = w - b * H;
            ATTN_ACC(9);
            // The head's per-query constants.  dO is in shared memory for the MMAs and O arrives with it, so D costs each of
            // the 512 threads a few 16-byte chunk pairs and three shuffles: rows 0..63 as soon as barrier `a` fires (the
            // first S^T / dP^T chunk is still in the tensor core), the rest before the second chunk.  (Two spare warps
            // summing the rows from global memory -- or from shared memory, a row per lane -- took 3.5-20 k clocks per head
            // next to sixteen polling warps, and the exp warps waited on them.)  lse is prefetched a head ahead into one
            // register of the first 208 threads.
            auto d_rows = [&](int q, const uint8_t* so, const uint8_t* sg) {      // 8 lanes per row
                const int c = tid & 7;
                float d = 0.f;
                if (q < T) {
                    const uint32_t off = q * 128 + ((c ^ (q & 7)) << 4);
                    const uint4 uo = *reinterpret_cast<const uint4*>(so + off);
                    const uint4 ug = *reinterpret_cast<const uint4*>(sg + off);
                    float fo[8], fg[8];
                    unpack8(*reinterpret_cast<const bf16x8*>(&uo), fo);
                    unpack8(*reinterpret_cast<const bf16x8*>(&ug), fg);
#pragma unroll
                    for (int e = 0; e < 8; ++e) d += fo[e] * fg[e];
                }
                d += __shfl_xor_sync(0xffffffffu, d, 1);
                d += __shfl_xor_sync(0xffffffffu, d, 2);
                d += __shfl_xor_sync(0xffffffffu, d, 4);
                if (c == 0 && q < KV_ROWS) dsum_s[q] = d * scale;
            };
            mbar_wait(&full_a, it & 1);
            d_rows(tid >> 3, smem + ((it & 1) ? P_XO : P_O), smem + ((it & 1) ? P_XDO : B_DO));
            if (tid < KV_ROWS) lse2_s[tid] = lse_next * LOG2E;
            named_bar_sync(1, 512);
            if (tid < T && w + (int)gridDim.x < items) lse_next = __ldg(lse + (long long)(w + gridDim.x) * T + tid);
            ATTN_ACC(0);
            const float* lse2c = lse2_s;
            const float* dsumc = dsum_s;
            for (int kt = 0; kt < 2; ++kt, ++ktg) {
                const int key = kt * 128 + row;
                const bool key_ok = key < T;
                for (int qc = 0; qc < 4; ++qc) {
                    const int bf = qc & 1;
                    if (kt == 0 && qc == 1) {                    // D of the queries 64.. (their rows came with barrier `b`)
                        mbar_wait(&full_b, it & 1);
#pragma unroll
                        for (int i = 1; i < 4; ++i) d_rows((tid >> 3) + 64 * i, smem + P_O, smem + B_DO);
                        named_bar_sync(1, 512);
                    }
                    mbar_wait(&sd_full[bf], use[bf] & 1);
                    tc_fence_after();
                    ATTN_ACC(1);
                    const bool active = qc < 3 || cg == 0;        // the last chunk holds 16 queries
                    float p[16], ds[16];
                    if (active) {
                        uint32_t s[16], dp[16];
                        tmem_ld16(tlane + P_SD + bf * 128 + cg * 16, s);
                        tmem_ld16(tlane + P_SD + bf * 128 + 64 + cg * 16, dp);
                        const int q0 = qc * 64 + cg * 16;
                        float lq[16], dq[16];                     // the 16 queries' constants: 8 broadcast LDS.128
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 l = *reinterpret_cast<const float4*>(lse2c + q0 + 4 * j);
                            const float4 d = *reinterpret_cast<const float4*>(dsumc + q0 + 4 * j);
                            lq[4 * j] = l.x; lq[4 * j + 1] = l.y; lq[4 * j + 2] = l.z; lq[4 * j + 3] = l.w;
                            dq[4 * j] = d.x; dq[4 * j + 1] = d.y; dq[4 * j + 2] = d.z; dq[4 * j + 3] = d.w;
                        }
                        tmem_ld_wait();
                        if (kt == 0 && q0 + 16 <= T) {            // every key and query of this thread's slice is real: no masks
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float e = ex2_approx(__uint_as_float(s[j]) * sl2 - lq[j]);
                                p[j] = e;
                                ds[j] = e * (__uint_as_float(dp[j]) * scale - dq[j]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float e = ex2_approx(__uint_as_float(s[j]) * sl2 - lq[j]);
                                const bool ok = key_ok && (q0 + j < T);
                                p[j] = ok ? e : 0.f;
                                ds[j] = ok ? e * (__uint_as_float(dp[j]) * scale - dq[j]) : 0.f;
                            }
                        }
                    }
                    ATTN_ACC(2);
                    // the MMAs of the pair that last used this smem buffer have finished reading it
                    mbar_wait(&tile_free[bf], (use[bf] & 1) ^ 1);
                    ATTN_ACC(3);
                    ++use[bf];
                    if (qc == kt) {                              // the tiles about to be written doubled as gradient staging
                        if (storer) bulk_wait_read0();           // (dQ in PT[0] / DST[0], dV / dK in PT[1] / DST[1]): the TMA
                        named_bar_sync(1, 512);                  // stores issued since have read them
                    }
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
                    ATTN_ACC(4);
                }
                mbar_wait(&acc_full, ktg & 1);
                tc_fence_after();
                ATTN_ACC(5);
                // dV / dK of this key tile leave through shared memory and TMA: a thread owns one row, so direct global
                // stores would touch 32 different lines per instruction (measured: 3 k clocks per tile).  Column group
                // (which, half) stages 32 columns of dV (which 0) or dK (which 1); the TMA store clips the keys >= T.
                {
                    const int which = cg >> 1, half = cg & 1;
                    uint32_t r[32];
                    tmem_ld32(tlane + (which == 0 ? C_DV : C_DK) + half * 32, r);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty);   // the accumulators are free for the next key tile
                    const uint32_t sST = smem_u32(smem + (which == 0 ? P_STAGE0 : P_STAGE1)) + row * 128;
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
                        st_shared_v4(sST + (((half * 4 + gq) ^ (row & 7)) << 4), pack8(v));
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 512);
                    if (storer) {
                        tma_store_3d(&maps.dqkv, smem + P_STAGE0, (2 * H + h) * HD, kt * 128, b);
                        tma_store_3d(&maps.dqkv, smem + P_STAGE1, (H + h) * HD, kt * 128, b);
                        bulk_commit();
                    }
                }
                ATTN_ACC(6);
            }
            mbar_wait(&dq_full, it & 1);
            tc_fence_after();
            ATTN_ACC(7);
            {
                const int qt = cg >> 1, half = cg & 1;          // column group (qt, half): 32 columns of query tile qt
                uint32_t r[32];
                tmem_ld32(tlane + C_DQ + qt * HD + half * 32, r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dq_empty);
                const uint32_t sST = smem_u32(smem + (qt == 0 ? P_PT0 : P_DST0)) + row * 128;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
                    st_shared_v4(sST + (((half * 4 + gq) ^ (row & 7)) << 4), pack8(v));
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 512);
                if (storer) {
                    tma_store_3d(&maps.dqkv, smem + P_PT0, h * HD, 0, b);
                    tma_store_3d(&maps.dqkv, smem + P_DST0, h * HD, 128, b);
                    bulk_commit();
                }
            }
            ATTN_ACC(8);
#ifdef DFU_ATTN_TRACE
            if (trb && it < 16)
                for (int c = 0; c < 12; ++c) g_attn_trace[3][it][c] = tr_acc[c];     // running sums up to this head
#endif
        }
        if (storer) bulk_wait0();                            // the last gradients are in global memory before the CTA retires
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
    CUtensorMap map_q, map_kv, map_o;
    const long long rows = (long long)B * T, cols = 3LL * H * HD;
    if (int rc = make_map_2d(&map_q, qkv, cols, rows, cols, HD, 128)) return rc;
    if (int rc = make_map_2d(&map_kv, qkv, cols, rows, cols, HD, KV_ROWS)) return rc;
    if (int rc = make_map_3d(&map_o, out, (long long)H * HD, T, B, (long long)H * HD, HD, 128)) return rc;
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
    dfu_launch(attn_fwd_kernel, grid, THREADS, FWD_SMEM, (cudaStream_t)stream, map_q, map_kv, map_o, lse, B, T, H, scale);
    DFU_CHECK_LAUNCH("attn_fwd_kernel");
    return DFU_OK;
}

extern "C" int dfu_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int32_t B,
                                 int32_t T, int32_t H, float scale, void* stream) {
    DFU_CHECK_ARG(qkv && out && dout && lse && dqkv, "dfu_attention_bwd: null pointer");
    if (int rc = check_shape("dfu_attention_bwd", B, T, H)) return rc;
    DFU_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)dout & 15) == 0 && ((uintptr_t)dqkv & 15) == 0 &&
                      ((uintptr_t)out & 15) == 0,
                  "dfu_attention_bwd: unaligned buffers");
    const long long rows = (long long)B * T, c3 = 3LL * H * HD, c1 = (long long)H * HD;
    CUtensorMap map_qkv, map_do;                                 // serial kernel: whole-head boxes
    BwdMaps maps;
    if (g_bwd_pipe) {
        if (int rc = make_map_2d(&maps.qkv_64, qkv, c3, rows, c3, HD, 64)) return rc;
        if (int rc = make_map_2d(&maps.qkv_144, qkv, c3, rows, c3, HD, KV_ROWS - 64)) return rc;
        if (int rc = make_map_2d(&maps.qkv_128, qkv, c3, rows, c3, HD, 128)) return rc;
        if (int rc = make_map_2d(&maps.qkv_80, qkv, c3, rows, c3, HD, KV_ROWS - 128)) return rc;
        if (int rc = make_map_2d(&maps.do_64, dout, c1, rows, c1, HD, 64)) return rc;
        if (int rc = make_map_2d(&maps.do_144, dout, c1, rows, c1, HD, KV_ROWS - 64)) return rc;
        if (int rc = make_map_2d(&maps.out_64, out, c1, rows, c1, HD, 64)) return rc;
        if (int rc = make_map_2d(&maps.out_144, out, c1, rows, c1, HD, KV_ROWS - 64)) return rc;
        if (int rc = make_map_3d(&maps.dqkv, dqkv, c3, T, B, c3, HD, 128)) return rc;
    } else {
        if (int rc = make_map_2d(&map_qkv, qkv, c3, rows, c3, HD, KV_ROWS)) return rc;
        if (int rc = make_map_2d(&map_do, dout, c1, rows, c1, HD, KV_ROWS)) return rc;
    }
    if (!g_attr_bwd) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM);
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
        dfu_launch(attn_bwd_pipe_kernel, grid, THREADS, PIPE_SMEM, (cudaStream_t)stream, maps, lse, B, T, H, scale);
    else
        dfu_launch(attn_bwd_kernel, grid, THREADS, BWD_SMEM, (cudaStream_t)stream, map_qkv, map_do, (const bf16*)out, (const bf16*)dout,
                                                                           lse, (bf16*)dqkv, B, T, H, scale);
    DFU_CHECK_LAUNCH("attn_bwd_kernel");
    return DFU_OK;
}

#ifdef DFU_ATTN_TRACE
// Tracing build only (not declared in include/dfu_b200.h): copies CTA 0's phase clocks, 4 x 16 x 12 int64, to the host.
extern "C" int dfu_attn_trace_read_host(long long* out) {
    return cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(g_attn_trace)) == cudaSuccess ? DFU_OK : DFU_ERR_CUDA;
}
#endif
