// tcgen05 / TMEM / TMA GEMM for sm_100a: the one dense-contraction kernel behind every Linear and
// (im2col'ed) convolution of the fusion step, forward, dgrad and wgrad.
//
//   D[z][m][n] = act(alpha * sum_k A[z][m][k] * B[z][n][k] + bias[n]) + res[z][m][n]
//
// Persistent, warp-specialised:
//   warp 0      TMA producer (one lane): 4-stage ring of {A 128x64, B block_n x64} bf16 tiles, SWIZZLE_128B
//   warp 1      MMA issuer (one lane): tcgen05.mma cta_group::1 kind::f16, M=128, N=block_n, K=16 x4 per stage
//   warp 2      TMEM allocator (512 columns = two 128 x 256 fp32 accumulators, double buffered)
//   warps 4-11  epilogue: tcgen05.ld 32x32b.x32 -> bias / act / residual -> 128-bit global stores
// Both operands may be K-major or MN-major (UMMA descriptor + instruction-descriptor major bits), so
// dgrad (B MN-major) and wgrad (A and B MN-major) need no transposed copies.  Split-K (fp32 atomic
// accumulation) keeps all SMs busy on the skinny wgrad shapes.
#include "common.cuh"
#include "../../include/dfu_b200.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int MAX_BLOCK_N = 256;
constexpr int MAX_STAGES = 8;         // barrier array size (narrow column tiles run up to 8 x 24 KB stages)
constexpr int STAGES = 4;            // ring depth (3 when a second staging slab per epilogue warp is needed for aux)
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB
constexpr int B_TILE_BYTES = MAX_BLOCK_N * BLOCK_K * 2;    // 32 KB
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;   // 48 KB
constexpr int STG_BYTES = 4096;                             // per epilogue warp: 32 rows x 128 B, SWIZZLE_128B
constexpr int EPI_BUFS = 1;                                 // staging slabs per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 8 * EPI_BUFS * STG_BYTES + 1024;    // + alignment slack
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;

struct GemmParams {
    int M, N, K;
    int nb0, nb1;
    int block_n;
    int a_mn, b_mn;
    int split_k, num_kb, kb_per_split;
    int m_tiles, n_tiles;
    long long total_work;
    void* D;
    long long ldd, sd0, sd1;
    int d_fp32, atomic;
    bf16* aux;
    const float* bias;
    const bf16* res;
    long long ldr, sr0, sr1;
    float alpha;
    int act;
    int vec_ok;
    int tma_store;     // epilogue goes through swizzled smem + TMA store / reduce-add
    int stages;        // smem ring depth actually used (<= MAX_STAGES)
    int cg;            // 1: one CTA per tile; 2: CTA pair per 256-row tile (cta_group::2)
    int stage_bytes;   // stride of one ring stage (A tile + this CTA's part of the B tile)
    int epi_bufs;      // staging slab sets per epilogue warp (2: the TMA store of slab i overlaps the fill of slab i+1)
    // implicit-GEMM convolution (dfu_conv_gemm): the A (fwd, dgrad) or A and B (wgrad) tiles are gathered by TMA from the
    // NHWC activation tensor as boxes of pixels shifted by the filter tap; out-of-image pixels read as zero.
    int cv_mode;              // 0 plain GEMM, 1 forward, 2 dgrad (stride 1), 3 wgrad
    int cv_wb, cv_hb, cv_bb;  // pixel box of one tile (128 output pixels, or 64 pixels per k-block for wgrad)
    int cv_tw, cv_th;         // boxes per image row / column
    int cv_lw, cv_lh;         // log2(cv_wb), log2(cv_hb): the box sides are powers of two
    int cv_stride;            // convolution stride (forward / wgrad): the gathered box walks the input with this step
    // dgrad (mode 2) taps of this launch: r = cv_r0 + i * cv_tstep (i < cv_nr), s = cv_s0 + j * cv_tstep (j < cv_ns); the
    // gathered dy pixel of output pixel (h, w) of the launch's output view is (h + (cv_ph + pad - r) / cv_tstep, ...).
    // Stride 1: one launch, tstep 1, all taps.  Stride 2: one launch per parity class (cv_ph, cv_pw) of the dx pixels.
    int cv_r0, cv_s0, cv_nr, cv_ns, cv_tstep, cv_ph, cv_pw;
    int cv_S, cv_pad;         // filter width, padding
    int cv_cchunks;           // 64-channel chunks of the gathered tensor (fwd: Cin/64, dgrad: Cout/64)
    int cv_C;                 // channels of the tap-major weight columns (Cin)
    int cv_sh, cv_sb;         // epilogue slab = 32 tile rows = box (cv_wb, cv_sh, cv_sb)
    int cv_ow, cv_oh, cv_nb;  // extent of the output image batch (tile rows beyond it are clipped by the TMA store)
    float* colstats;   // per-column sum (and sum of squares) of D accumulated here when non-null
    int colstats_mode; // 1: sums[N]; 2: sums[N] then sums of squares[N]
    int res_vec;
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp, the same one every time (see tc_common.cuh)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// the staging slab about to be refilled was handed to TMA EPI_BUFS stores ago
__device__ __forceinline__ void bulk_wait_read_prev() {
    if (EPI_BUFS == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
// ---- cta_group::2 (CTA pair on one 256-row tile): forms taken from CUTLASS cute/arch/{copy_sm100_tma,mma_sm100_umma}.hpp
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even (leader) CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory, transaction bytes reported to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// MMA completion -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 remAddr32;\n\t"
        "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ bf16x8 ld_nc16(const bf16* p) {
    uint4 u;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    return *reinterpret_cast<bf16x8*>(&u);
}
__device__ __forceinline__ void sts16(uint32_t saddr, const void* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// ----------------------------------------------------------------------------- the kernel
// CS: compile the per-column statistics epilogue in (kept out of the default instantiation: it costs registers).
// EM: epilogue specialisation.  The fully general epilogue (EM 0: every feature behind a runtime flag, plus the
// direct-store path for outputs TMA cannot describe) is ~11k SASS instructions and its warps stall on instruction
// fetch (ncu: no_inst); the shapes that carry the step get instantiations with the flags folded at compile time:
//   1 bf16 store   2 bf16 + bias   3 bf16 + bias + residual   4 bf16 + bias + GELU + pre-activation copy   5 fp32 store
//   6 bf16 x residual operand (stored GELU')   7 bf16 + bias + GELU, aux = GELU'   8 bf16 + residual (no bias)
//   9 bf16 + bias + ReLU   10 bf16 + bias + residual, ReLU after the add (9 / 10: eval-mode conv with BatchNorm folded in)
// CG: 1 = one CTA per 128-row tile; 2 = a CTA pair (cluster of 2) shares a 256-row tile: tcgen05.mma.cta_group::2 issued by
//     the even CTA, each CTA staging its 128 rows of A and HALF of the B tile, so every SM pulls a third less operand data
//     per FLOP through its L2 port -- the limit of the 128x256 single-CTA kernel on the K >= 768 shapes.
// Work item -> (n_tile, m_tile, k split, z0, z1), n_tile fastest.  32-bit arithmetic, and no division at all for the
// factors that are 1 (the decode runs once per tile in every role; 64-bit runtime divisions were ~15% of the epilogue
// instructions of the short-K shapes).
struct WorkItem { int n_tile, m_tile, ks, z0, z1; };
__device__ __forceinline__ WorkItem decode_work(uint32_t w, const GemmParams& p) {
    WorkItem o;
    uint32_t t = w;
    if (p.n_tiles == 1) { o.n_tile = 0; } else { o.n_tile = (int)(t % (uint32_t)p.n_tiles); t /= (uint32_t)p.n_tiles; }
    if (p.split_k == 1 && p.nb0 == 1 && p.nb1 == 1) {
        o.m_tile = (int)t; o.ks = 0; o.z0 = 0; o.z1 = 0;
        return o;
    }
    o.m_tile = (int)(t % (uint32_t)p.m_tiles); t /= (uint32_t)p.m_tiles;
    o.ks = (int)(t % (uint32_t)p.split_k); t /= (uint32_t)p.split_k;
    o.z0 = (int)(t % (uint32_t)p.nb0);
    o.z1 = (int)(t / (uint32_t)p.nb0);
    return o;
}

template <bool CS, int EM, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dfu_tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_x,
                   const GemmParams p) {
    pdl_trigger();        // the next kernel of the stream may be scheduled as soon as this grid is resident (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[MAX_STAGES];
    __shared__ uint64_t empty_bar[MAX_STAGES];
    __shared__ uint64_t tmem_full_bar[2];
    __shared__ uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    if (warp_idx == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    }
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;      // 0 = leader of the pair (issues the MMAs)
    const int work0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int work_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS * CG);       // the epilogue warps of both CTAs release the leader
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_idx == 2) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&tmem_base_slot)),
                         "r"((uint32_t)TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&tmem_base_slot)),
                         "r"((uint32_t)TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (CS && p.colstats) {        // CTA-wide per-column accumulators, flushed with one global atomic per column at the end
        float* colacc0 = reinterpret_cast<float*>(smem + p.stages * p.stage_bytes + 8 * p.epi_bufs * (p.aux ? 2 : 1) * STG_BYTES);
        for (int i = threadIdx.x; i < p.N * p.colstats_mode; i += NUM_THREADS) colacc0[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();      // the peer's barriers are initialised before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    // barriers, TMEM and tensor-map prefetch are set up: only now wait for the previous kernel's results
    pdl_wait();

    // bytes landing per ring stage: both CTAs' tiles are counted on the leader's barrier
    const int bn_cta = p.block_n / CG;                              // B-tile rows (N columns) staged by this CTA
    const uint32_t stage_tx = (uint32_t)CG * (A_TILE_BYTES + (uint32_t)bn_cta * BLOCK_K * 2);

    if (warp_idx == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (uint32_t w = work0; w < (uint32_t)p.total_work; w += work_step) {
                const WorkItem wi = decode_work(w, p);
                const int n_tile = wi.n_tile, m_tile = wi.m_tile, ks = wi.ks, z0 = wi.z0, z1 = wi.z1;
                const int m0 = (m_tile * CG + (int)cta_rank) * BLOCK_M, n0 = n_tile * p.block_n + (int)cta_rank * bn_cta;
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
                // implicit conv: the tile's pixel-box origin once per tile, (filter row, filter column, channel chunk) of
                // the k-block as running counters -- no division on the per-k-block path of the issuing thread
                int cvw0 = 0, cvh0 = 0, cvb0 = 0, cv_r = 0, cv_s = 0, cv_cc = 0;
                if (p.cv_mode == 1 || p.cv_mode == 2) {
                    const int wt = m_tile % p.cv_tw, hb_ = m_tile / p.cv_tw;
                    cvw0 = wt * p.cv_wb; cvh0 = (hb_ % p.cv_th) * p.cv_hb; cvb0 = (hb_ / p.cv_th) * p.cv_bb;
                    const int tap0 = kb0 / p.cv_cchunks;
                    cv_cc = kb0 - tap0 * p.cv_cchunks; cv_r = tap0 / p.cv_S; cv_s = tap0 - cv_r * p.cv_S;
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * p.stage_bytes;
                    uint8_t* sb = sa + A_TILE_BYTES;
                    if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
                    const int k0 = kb * BLOCK_K;
                    if (CG == 2) {
                        // each CTA of the pair: its 128 rows of A, its half of the B tile; bytes counted by the leader
                        if (!p.a_mn) {
                            tma_load_4d_2sm(&map_a, &full_bar[stage], sa, k0, m0, z0, z1);
                        } else {
                            tma_load_4d_2sm(&map_a, &full_bar[stage], sa, m0, k0, z0, z1);
                            tma_load_4d_2sm(&map_a, &full_bar[stage], sa + 8192, m0 + 64, k0, z0, z1);
                        }
                        if (!p.b_mn) {
                            tma_load_4d_2sm(&map_b, &full_bar[stage], sb, k0, n0, z0, z1);
                        } else {
                            for (int j = 0; j < bn_cta / 64; ++j)
                                tma_load_4d_2sm(&map_b, &full_bar[stage], sb + j * 8192, n0 + j * 64, k0, z0, z1);
                        }
                    } else
                    if (p.cv_mode == 1 || p.cv_mode == 2) {
                        // tile = box of output pixels; k-block = (filter tap, 64-channel chunk) of the gathered tensor
                        const bool fwd = p.cv_mode == 1;
                        const int r = fwd ? cv_r : p.cv_r0 + cv_r * p.cv_tstep, sx = fwd ? cv_s : p.cv_s0 + cv_s * p.cv_tstep;
                        const int cc = cv_cc, tap = r * p.cv_S + sx;
                        if (++cv_cc == p.cv_cchunks) { cv_cc = 0; if (++cv_s == (fwd ? p.cv_S : p.cv_ns)) { cv_s = 0; ++cv_r; } }
                        // dgrad: exact division (the class's taps have the parity of ph + pad), possibly of a negative number
                        const int dw = fwd ? sx - p.cv_pad : (p.cv_pw + p.cv_pad - sx) >> (p.cv_tstep - 1);
                        const int dh = fwd ? r - p.cv_pad : (p.cv_ph + p.cv_pad - r) >> (p.cv_tstep - 1);
                        tma_load_4d(&map_a, &full_bar[stage], sa, cc * 64, cvw0 * p.cv_stride + dw, cvh0 * p.cv_stride + dh, cvb0);
                        if (p.cv_mode == 1) {
                            tma_load_4d(&map_b, &full_bar[stage], sb, k0, n0, 0, 0);
                        } else {                         // weights [Cout, (tap, cin)] read as an MN-major operand
                            for (int j = 0; j < p.block_n / 64; ++j)
                                tma_load_4d(&map_b, &full_bar[stage], sb + j * 8192, tap * p.cv_C + n0 + j * 64, cc * 64, 0, 0);
                        }
                    } else if (p.cv_mode == 3) {
                        // k-block = box of 64 pixels; A = dy (channels contiguous), B = x shifted by the tap of each
                        // 64-column group of the (tap, cin) output columns
                        const int wt = kb % p.cv_tw, ht = (kb / p.cv_tw) % p.cv_th, bt = kb / (p.cv_tw * p.cv_th);
                        const int w0 = wt * p.cv_wb, h0 = ht * p.cv_hb, b0 = bt * p.cv_bb;
                        tma_load_4d(&map_a, &full_bar[stage], sa, m0, w0, h0, b0);
                        tma_load_4d(&map_a, &full_bar[stage], sa + 8192, m0 + 64, w0, h0, b0);
                        for (int j = 0; j < p.block_n / 64; ++j) {
                            const int ncol = n0 + j * 64;
                            const int tap = ncol / p.cv_C, c0 = ncol - tap * p.cv_C;
                            const int r = tap / p.cv_S, sx = tap - r * p.cv_S;
                            tma_load_4d(&map_b, &full_bar[stage], sb + j * 8192, c0, w0 * p.cv_stride + sx - p.cv_pad,
                                        h0 * p.cv_stride + r - p.cv_pad, b0);
                        }
                    } else if (!p.a_mn) {
                        tma_load_4d(&map_a, &full_bar[stage], sa, k0, m0, z0, z1);
                    } else {
                        tma_load_4d(&map_a, &full_bar[stage], sa, m0, k0, z0, z1);
                        tma_load_4d(&map_a, &full_bar[stage], sa + 8192, m0 + 64, k0, z0, z1);
                    }
                    if (CG == 2 || p.cv_mode) {
                        // B already issued above
                    } else if (!p.b_mn) {
                        tma_load_4d(&map_b, &full_bar[stage], sb, k0, n0, z0, z1);
                    } else {
                        for (int j = 0; j < p.block_n / 64; ++j)
                            tma_load_4d(&map_b, &full_bar[stage], sb + j * 8192, n0 + j * 64, k0, z0, z1);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (cta_rank == 0) {
            // ------------------------------------------------------------------ MMA issuer (leader CTA only when CG == 2)
            // The whole warp walks the schedule and waits; one elected lane issues.  (Under a `lane == 0` branch ptxas wraps
            // every tcgen05.mma in a vote / elect loop with R2UR moves, ~17 instructions per MMA -- more than a 128 x 64 x 16
            // MMA takes on the narrow tiles of the ResNet branch; measured on the attention kernels)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) |
                                   ((uint32_t)p.b_mn << 16) | ((uint32_t)(p.block_n >> 3) << 17) |
                                   ((uint32_t)((BLOCK_M * CG) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (uint32_t w = work0; w < (uint32_t)p.total_work; w += work_step, ++it) {
                const int ks = decode_work(w, p).ks;
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * MAX_BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
                    const uint32_t sb = sa + A_TILE_BYTES;
                    if (elect_one()) {
                        const uint64_t adesc0 = p.a_mn ? make_desc(sa, 8192, 1024) : make_desc(sa, 16, 1024);
                        const uint64_t bdesc0 = p.b_mn ? make_desc(sb, 8192, 1024) : make_desc(sb, 16, 1024);
                        const uint64_t astep = p.a_mn ? (2048 >> 4) : (32 >> 4), bstep = p.b_mn ? (2048 >> 4) : (32 >> 4);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k) {
                            if (CG == 2) tc_mma_f16_2sm(tmem_d, adesc0 + k * astep, bdesc0 + k * bstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                            else tc_mma_f16(tmem_d, adesc0 + k * astep, bdesc0 + k * bstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
                        if (CG == 2) tc_commit_2sm(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
                        // accumulator complete -> epilogue warps (of both CTAs)
                        if (kb + 1 == kb1) {
                            if (CG == 2) tc_commit_2sm(&tmem_full_bar[acc]); else tc_commit(&tmem_full_bar[acc]);
                        }
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx >= EPI_WARP0) {
        // ---------------------------------------------------------------------- epilogue
        const int ew = warp_idx - EPI_WARP0;
        constexpr bool GEN = (EM == 0);
        const bool f_tma = GEN ? (p.tma_store != 0) : true;
        const bool f_fp32 = GEN ? (p.d_fp32 != 0) : (EM == 5);
        const bool f_bias = GEN ? (p.bias != nullptr) : ((EM >= 2 && EM <= 4) || EM == 7 || EM == 9 || EM == 10);
        const int f_act = GEN ? p.act : (EM == 4 ? 2 : EM == 6 ? 5 : EM == 7 ? 4 : EM == 9 ? 1 : EM == 10 ? 6 : 0);
        const bool f_res = GEN ? (p.res != nullptr) : (EM == 3 || EM == 6 || EM == 8 || EM == 10);
        const bool f_aux = GEN ? (p.aux != nullptr) : (EM == 4 || EM == 7);
        const float alpha = GEN ? p.alpha : 1.f;
        const int quarter = warp_idx & 3;        // TMEM lane quarter this warp may touch
        const int group = ew >> 2;               // two warps share a quarter, alternate 32-column chunks
        const int n_chunks = p.block_n / 32;
        // column statistics: each lane keeps (sum, sum of squares) of its two columns of the warp's two slabs of the
        // current column tile in registers across work items; they go to the CTA's shared accumulators only when the
        // column tile changes (the host sizes the grid so that it rarely does) -- shared float atomics are CAS loops
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int cs_tile = -1;
        int bufsel = 0;            // staging set in use (epi_bufs == 2): alternates per slab ACROSS tiles, so that "at most
                                   // one bulk store still reading" always refers to the other set
        float* const cs_acc = reinterpret_cast<float*>(smem + p.stages * p.stage_bytes + 8 * p.epi_bufs * (f_aux ? 2 : 1) * STG_BYTES);
        auto cs_flush = [&]() {
            if (cs_tile < 0) return;
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) {
                const int col = cs_tile * p.block_n + (group + 2 * ai) * 64 + 2 * lane;
                if ((group + 2 * ai) * 64 < p.block_n) {
                    if (col < p.N) {
                        atomicAdd(cs_acc + col, cs[ai * 4]);
                        if (p.colstats_mode == 2) atomicAdd(cs_acc + p.N + col, cs[ai * 4 + 2]);
                    }
                    if (col + 1 < p.N) {
                        atomicAdd(cs_acc + col + 1, cs[ai * 4 + 1]);
                        if (p.colstats_mode == 2) atomicAdd(cs_acc + p.N + col + 1, cs[ai * 4 + 3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) cs[i] = 0.f;
        };
        int it = 0;
        for (uint32_t w = work0; w < (uint32_t)p.total_work; w += work_step, ++it) {
            const WorkItem wi = decode_work(w, p);
            const int n_tile = wi.n_tile, z0 = wi.z0, z1 = wi.z1;
            const int m_tile = wi.m_tile * CG + (int)cta_rank;                                // this CTA's 128-row tile
            if (CS && p.colstats && n_tile != cs_tile) { cs_flush(); cs_tile = n_tile; }
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const int m = m_tile * BLOCK_M + quarter * 32 + lane;
            const bool row_ok = m < p.M;
            int cv_w0 = 0, cv_h0 = 0, cv_b0 = 0;      // implicit conv: origin of this tile's pixel box in the output image
            if (p.cv_mode == 1 || p.cv_mode == 2) {
                const int wt = m_tile % p.cv_tw, hb_ = m_tile / p.cv_tw;
                cv_w0 = wt * p.cv_wb; cv_h0 = (hb_ % p.cv_th) * p.cv_hb; cv_b0 = (hb_ / p.cv_th) * p.cv_bb;
            }
            uint32_t cs_rows = 0;            // which of this warp's 32 tile rows are samples of the column statistics
            if (CS && p.colstats) {
                bool row_in = row_ok;
                if (p.cv_mode) {
                    // a tile row outside the image still holds a value (its taps reach inside): not a sample
                    const int idx = quarter * 32 + lane;
                    row_in = cv_w0 + (idx & (p.cv_wb - 1)) < p.cv_ow &&
                             cv_h0 + ((idx >> p.cv_lw) & (p.cv_hb - 1)) < p.cv_oh &&
                             cv_b0 + (idx >> (p.cv_lw + p.cv_lh)) < p.cv_nb;
                }
                cs_rows = __ballot_sync(0xffffffffu, row_in);
            }
            const long long d_off = (long long)z1 * p.sd1 + (long long)z0 * p.sd0 + (long long)m * p.ldd;
            const long long r_off = (long long)z1 * p.sr1 + (long long)z0 * p.sr0 + (long long)m * p.ldr;
            if (f_tma) {
                // TMEM -> registers -> bias / act / residual -> swizzled smem slab (32 rows x 128 B) -> TMA store.
                // The two warps of a lane quarter alternate slabs; TMA clips rows >= M and columns >= N.  Everything
                // with latency is issued ahead of its use: the next 32-column TMEM load while the current chunk is in
                // the FP32 pipes, and the next slab's residual row (8 x 16 B per thread) one slab early.
                const int set_bytes = (f_aux ? 2 : 1) * STG_BYTES;          // one staging set: out slab (+ aux slab)
                uint8_t* const stg_base = smem + p.stages * p.stage_bytes + ew * p.epi_bufs * set_bytes;
                const uint32_t sa_base = smem_u32(stg_base) + lane * 128;
                const int slab_cols = f_fp32 ? 32 : 64;
                const int n_slabs = p.block_n / slab_cols;
                const int row0 = m_tile * BLOCK_M + quarter * 32;
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * MAX_BLOCK_N;
                const bool use_res = f_res;
                const bf16* const rrow = p.res + r_off;
                bf16x8 rres[8];

                auto load_res = [&](int sl) {
                    const int ncol = n_tile * p.block_n + sl * slab_cols;
                    const int nq = slab_cols / 8;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q < nq) {
                            const int n = ncol + q * 8;
                            if (row_ok && p.res_vec && n + 8 <= p.N) {
                                rres[q] = ld_nc16(rrow + n);
                            } else {
                                alignas(16) bf16 t[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    t[j] = (row_ok && n + j < p.N) ? rrow[n + j] : __float2bfloat16(0.f);
                                rres[q] = *reinterpret_cast<const bf16x8*>(t);
                            }
                        }
                    }
                };
                // one 32-column chunk: r = raw accumulators of columns [nbase, nbase + 32) of this thread's row
                auto process = [&](uint32_t* r, int sl, int ch, int roff) {
                    const int nbase = n_tile * p.block_n + sl * slab_cols + ch * 32;
                    float v[32];
                    if (f_bias) {
                        if (nbase + 32 <= p.N) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase) + q);
                                v[q * 4] = fmaf(__uint_as_float(r[q * 4]), alpha, b4.x);
                                v[q * 4 + 1] = fmaf(__uint_as_float(r[q * 4 + 1]), alpha, b4.y);
                                v[q * 4 + 2] = fmaf(__uint_as_float(r[q * 4 + 2]), alpha, b4.z);
                                v[q * 4 + 3] = fmaf(__uint_as_float(r[q * 4 + 3]), alpha, b4.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = fmaf(__uint_as_float(r[j]), alpha, (nbase + j < p.N) ? __ldg(p.bias + nbase + j) : 0.f);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
                    }
                    const uint32_t sa_out = sa_base + bufsel * set_bytes, sa_aux = sa_out + STG_BYTES;
                    if (ch == 0) {                           // the TMA stores that last used this staging set have read it
                        if (lane == 0) {
                            if (p.epi_bufs == 2) bulk_wait_read1(); else bulk_wait_read0();
                        }
                        __syncwarp();
                    }
                    if (f_aux) {                             // pre-activation copy (bf16 outputs only)
                        const uint32_t rowp = sa_aux;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            bf16x8 pk;
                            if (f_act == 4) {                // aux = gelu'(pre-activation), v = gelu(pre-activation)
                                float dv[8];
#pragma unroll
                                for (int j = 0; j < 8; j += 2)
                                    gelu_both_fast2(v[q * 8 + j], v[q * 8 + j + 1], v[q * 8 + j], v[q * 8 + j + 1], dv[j], dv[j + 1]);
                                pk = pack8(dv);
                            } else {
                                pk = pack8(v + q * 8);
                            }
                            sts16(rowp + (((ch * 4 + q) ^ (lane & 7)) << 4), &pk);
                        }
                    }
                    if (f_act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (f_act == 2) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) gelu_fast2(v[j], v[j + 1], v[j], v[j + 1]);
                    }
                    if (use_res) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float rv[8];
                            unpack8(rres[roff + q], rv);
                            if (f_act == 3) {                // res holds the pre-activation h: D *= gelu'(h)
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[q * 8 + j] *= gelu_grad_fast(rv[j]);
                            } else if (f_act == 5) {         // res holds the activation derivative: D *= res
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[q * 8 + j] *= rv[j];
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[q * 8 + j] += rv[j];
                            }
                        }
                        if (f_act == 6) {                    // ReLU after the residual add (Bottleneck tail)
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                        }
                    }
                    const uint32_t rowp = sa_out;
                    if (f_fp32) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) sts16(rowp + ((q ^ (lane & 7)) << 4), v + q * 4);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const bf16x8 pk = pack8(v + q * 8);
                            sts16(rowp + (((ch * 4 + q) ^ (lane & 7)) << 4), &pk);
                        }
                    }
                };
                auto store_slab = [&](int sl) {
                    const int ncol0 = n_tile * p.block_n + sl * slab_cols;
                    uint8_t* const stg_out = stg_base + bufsel * set_bytes;
                    if (CS && p.colstats && !f_fp32) {
                        // column statistics of the slab just staged (32 rows x 64 bf16 columns): lane L walks column pair
                        // (2L, 2L+1) down the swizzled slab (conflict-free), then one shared-memory atomic per column
                        __syncwarp();
                        const uint32_t sbase = smem_u32(stg_out) + (lane & 3) * 4;
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                        if (cs_rows == 0xffffffffu) {
                            uint64_t s2 = f2_pack(0.f, 0.f), q2 = s2;       // both columns per FADD2 / FFMA2
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                uint32_t w;
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(sbase + r * 128 + (((lane >> 2) ^ (r & 7)) << 4)) : "memory");
                                const uint64_t f = f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
                                s2 = f2_add(s2, f);
                                q2 = f2_fma(f, f, q2);
                            }
                            f2_unpack(s2, s0, s1);
                            f2_unpack(q2, q0, q1);
                        } else {
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                if ((cs_rows >> r) & 1u) {           // warp-uniform
                                    uint32_t w;
                                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(sbase + r * 128 + (((lane >> 2) ^ (r & 7)) << 4)) : "memory");
                                    const float2 f = __bfloat1622float2(*reinterpret_cast<const bf162*>(&w));
                                    s0 += f.x; s1 += f.y;
                                    q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
                                }
                            }
                        }
                        if (((sl >> 1) & 1) == 0) { cs[0] += s0; cs[1] += s1; cs[2] += q0; cs[3] += q1; }
                        else                      { cs[4] += s0; cs[5] += s1; cs[6] += q0; cs[7] += q1; }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.cv_mode == 1 || p.cv_mode == 2) {
                            // this warp's 32 tile rows are the pixel box (cv_wb, cv_sh, cv_sb) of the output image
                            const int idx0 = quarter * 32;
                            tma_store_4d(&map_d, stg_out, ncol0, cv_w0, cv_h0 + ((idx0 >> p.cv_lw) & (p.cv_hb - 1)),
                                         cv_b0 + (idx0 >> (p.cv_lw + p.cv_lh)));
                        } else {
                            if (f_aux) tma_store_4d(&map_x, stg_out + STG_BYTES, ncol0, row0, z0, z1);
                            if (p.atomic) tma_reduce_add_4d(&map_d, stg_out, ncol0, row0, z0, z1);
                            else tma_store_4d(&map_d, stg_out, ncol0, row0, z0, z1);
                        }
                        bulk_commit();
                    }
                    if (p.epi_bufs == 2) bufsel ^= 1;
                };
                auto slab_ok = [&](int sl) { return sl < n_slabs && n_tile * p.block_n + sl * slab_cols < p.N; };

                uint32_t ra[32], rb[32];
                int sl = group;
                if (slab_ok(sl)) {
                    tmem_ld32(tbase + sl * slab_cols, ra);
                    if (use_res) load_res(sl);
                }
                if (!f_fp32) {
                    while (slab_ok(sl)) {
                        const bool more = slab_ok(sl + 2);
                        tmem_ld_wait();
                        tmem_ld32(tbase + sl * 64 + 32, rb);
                        process(ra, sl, 0, 0);
                        tmem_ld_wait();
                        if (more) tmem_ld32(tbase + (sl + 2) * 64, ra);
                        process(rb, sl, 1, 4);
                        if (more && use_res) load_res(sl + 2);
                        store_slab(sl);
                        sl += 2;
                    }
                } else {
                    while (slab_ok(sl)) {
                        bool more = slab_ok(sl + 2);
                        tmem_ld_wait();
                        if (more) tmem_ld32(tbase + (sl + 2) * 32, rb);
                        process(ra, sl, 0, 0);
                        if (more && use_res) load_res(sl + 2);
                        store_slab(sl);
                        sl += 2;
                        if (!more) break;
                        more = slab_ok(sl + 2);
                        tmem_ld_wait();
                        if (more) tmem_ld32(tbase + (sl + 2) * 32, ra);
                        process(rb, sl, 0, 0);
                        if (more && use_res) load_res(sl + 2);
                        store_slab(sl);
                        sl += 2;
                    }
                }
            } else if (EM == 0)
            for (int c = group; c < n_chunks; c += 2) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * MAX_BLOCK_N + c * 32;
                tmem_ld32(taddr, r);
                tmem_ld_wait();
                const int nbase = n_tile * p.block_n + c * 32;
                if (nbase >= p.N) continue;      // warp-uniform
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int n = nbase + g * 8;
                    if (n >= p.N) break;         // warp-uniform
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * p.alpha;
                    const bool full8 = (n + 8 <= p.N);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (n + j < p.N) v[j] += __ldg(p.bias + n + j);
                    }
                    if (!row_ok) continue;
                    if (p.aux) {
                        bf16* ap = p.aux + d_off + n;
                        float av[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            av[j] = v[j];
                            if (p.act == 4) gelu_both_fast(v[j], v[j], av[j]);
                        }
                        if (full8 && p.vec_ok) {
                            st8(ap, pack8(av));
                        } else {
                            for (int j = 0; j < 8; ++j)
                                if (n + j < p.N) ap[j] = __float2bfloat16(av[j]);
                        }
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act == 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
                    }
                    if (p.res) {
                        const bf16* rp = p.res + r_off + n;
                        float rv[8];
                        if (full8 && p.vec_ok) {
                            unpack8(ld8(rp), rv);
                        } else {
                            for (int j = 0; j < 8; ++j) rv[j] = (n + j < p.N) ? __bfloat162float(rp[j]) : 0.f;
                        }
                        if (p.act == 3) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] *= gelu_grad_fast(rv[j]);
                        } else if (p.act == 5) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] *= rv[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] += rv[j];
                        }
                        if (p.act == 6) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                        }
                    }
                    if (p.d_fp32) {
                        float* dp = reinterpret_cast<float*>(p.D) + d_off + n;
                        if (p.atomic == 2) {          // split-K partial sums
                            if (full8 && p.vec_ok) {
                                red_add_v4(dp, v[0], v[1], v[2], v[3]);
                                red_add_v4(dp + 4, v[4], v[5], v[6], v[7]);
                            } else {
                                for (int j = 0; j < 8; ++j)
                                    if (n + j < p.N) atomicAdd(dp + j, v[j]);
                            }
                        } else if (p.atomic == 1) {   // sole owner of the tile: plain read-modify-write
                            if (full8 && p.vec_ok) {
                                float4 o0 = *reinterpret_cast<float4*>(dp), o1 = *reinterpret_cast<float4*>(dp + 4);
                                *reinterpret_cast<float4*>(dp) = make_float4(o0.x + v[0], o0.y + v[1], o0.z + v[2], o0.w + v[3]);
                                *reinterpret_cast<float4*>(dp + 4) = make_float4(o1.x + v[4], o1.y + v[5], o1.z + v[6], o1.w + v[7]);
                            } else {
                                for (int j = 0; j < 8; ++j)
                                    if (n + j < p.N) dp[j] += v[j];
                            }
                        } else if (full8 && p.vec_ok) {
                            *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
                            *reinterpret_cast<float4*>(dp + 4) = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
                            for (int j = 0; j < 8; ++j)
                                if (n + j < p.N) dp[j] = v[j];
                        }
                    } else {
                        bf16* dp = reinterpret_cast<bf16*>(p.D) + d_off + n;
                        if (full8 && p.vec_ok) {
                            st8(dp, pack8(v));
                        } else {
                            for (int j = 0; j < 8; ++j)
                                if (n + j < p.N) dp[j] = __float2bfloat16(v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cta(&tmem_empty_bar[acc], 0); else mbar_arrive(&tmem_empty_bar[acc]);
            }
        }
        if (CS && p.colstats) cs_flush();
        if (f_tma && lane == 0) bulk_wait0();   // all bulk stores of this warp have completed
    }

    tc_fence_before();
    __syncthreads();
    if (CS && p.colstats) {
        const float* colacc0 = reinterpret_cast<const float*>(smem + p.stages * p.stage_bytes + 8 * p.epi_bufs * (p.aux ? 2 : 1) * STG_BYTES);
        for (int i = threadIdx.x; i < p.N * p.colstats_mode; i += NUM_THREADS) {
            const float sv = colacc0[i];
            if (sv != 0.f) atomicAdd(p.colstats + i, sv);
        }
    }
    if (CG == 2) cluster_sync_all();      // neither CTA leaves (or frees TMEM) while its peer may still address it
    if (warp_idx == 2) {
        tc_fence_after();
        if (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                         : "memory");
    }
}

// ----------------------------------------------------------------------------- SIMT check kernel
// Straightforward fp32-accumulate GEMM over the same argument block: validates the tcgen05 kernel on the
// device and serves shapes TMA cannot describe (row strides that are not multiples of 16 bytes, N < 32).
struct RefParams {
    int M, N, K, nb0, nb1;
    const bf16* A; int a_mn; long long lda, sa0, sa1;
    const bf16* B; int b_mn; long long ldb, sb0, sb1;
    void* D; int d_fp32, atomic; long long ldd, sd0, sd1;
    bf16* aux; const float* bias; const bf16* res; long long ldr, sr0, sr1;
    float alpha; int act;
};

__global__ void __launch_bounds__(256) dfu_ref_gemm_kernel(const RefParams p) {
    pdl_sync();
    __shared__ float As[16][17];
    __shared__ float Bs[16][17];
    const int z = blockIdx.z;
    const int z0 = z % p.nb0, z1 = z / p.nb0;
    const bf16* A = p.A + z1 * p.sa1 + z0 * p.sa0;
    const bf16* B = p.B + z1 * p.sb1 + z0 * p.sb0;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        {
            const int am = blockIdx.y * 16 + ty, ak = k0 + tx;
            float v = 0.f;
            if (am < p.M && ak < p.K)
                v = __bfloat162float(p.a_mn ? A[(long long)ak * p.lda + am] : A[(long long)am * p.lda + ak]);
            As[ty][tx] = v;
            const int bn = blockIdx.x * 16 + ty, bk = k0 + tx;
            float u = 0.f;
            if (bn < p.N && bk < p.K)
                u = __bfloat162float(p.b_mn ? B[(long long)bk * p.ldb + bn] : B[(long long)bn * p.ldb + bk]);
            Bs[ty][tx] = u;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += As[ty][k] * Bs[tx][k];
        __syncthreads();
    }
    if (m >= p.M || n >= p.N) return;
    float v = acc * p.alpha;
    if (p.bias) v += p.bias[n];
    const long long d_off = z1 * p.sd1 + z0 * p.sd0 + (long long)m * p.ldd + n;
    if (p.aux) p.aux[d_off] = __float2bfloat16(p.act == 4 ? gelu_erf_grad(v) : v);
    if (p.act == 1) v = fmaxf(v, 0.f);
    else if (p.act == 2 || p.act == 4) v = gelu_erf(v);
    if (p.res) {
        const float rv = __bfloat162float(p.res[z1 * p.sr1 + z0 * p.sr0 + (long long)m * p.ldr + n]);
        if (p.act == 3) v *= gelu_erf_grad(rv); else if (p.act == 5) v *= rv; else v += rv;
        if (p.act == 6) v = fmaxf(v, 0.f);
    }
    if (p.d_fp32) {
        float* d = reinterpret_cast<float*>(p.D) + d_off;
        if (p.atomic) atomicAdd(d, v); else *d = v;
    } else {
        reinterpret_cast<bf16*>(p.D)[d_off] = __float2bfloat16(v);
    }
}

// ----------------------------------------------------------------------------- host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
bool g_encode_tried = false;

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    if (!g_encode_tried) {
        g_encode_tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
        (void)cudaGetLastError();
    }
    return g_encode;
}

// 4-D bf16 tensor map {inner, rows, z0, z1}; strides in elements.
int make_map(CUtensorMap* map, const void* base, long long inner, long long rows, long long ld, int nb0, long long s0,
             int nb1, long long s1, int box_inner, int box_rows, int esz = 2) {
    auto enc = get_encode();
    if (!enc) { dfu_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return DFU_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nb0, (cuuint64_t)nb1};
    // strides of dims 1..3 in bytes (multiples of 16); a batch dim of extent 1 only ever sees coordinate 0
    cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)(nb0 > 1 ? s0 : ld) * esz,
                             (cuuint64_t)(nb1 > 1 ? s1 : ld) * esz};
    cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dfu_set_error("cuTensorMapEncodeTiled failed (%d): inner=%lld rows=%lld ld=%lld box=%dx%d", (int)r, inner, rows,
                      ld, box_inner, box_rows);
        return DFU_ERR_CUDA;
    }
    return DFU_OK;
}

// General 4-D bf16/fp32 map: dims/strides in elements (stride[0] is implicit 1), box and traversal strides per dim.
int make_map_nd(CUtensorMap* map, const void* base, const long long dims_[4], const long long strides_[3], const int box_[4],
                int esz = 2, int walk = 1) {
    // walk > 1: the box steps through dims 1 and 2 (image width / height) with that stride; box_ then counts tensor
    // elements spanned, ceil(box / walk) of them land in shared memory (a strided convolution's gathered pixels)
    auto enc = get_encode();
    if (!enc) { dfu_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return DFU_ERR_CUDA; }
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], estr[4] = {1, (cuuint32_t)walk, (cuuint32_t)walk, 1};
    for (int i = 0; i < 4; ++i) { dims[i] = (cuuint64_t)dims_[i]; box[i] = (cuuint32_t)box_[i]; }
    for (int i = 0; i < 3; ++i) strides[i] = (cuuint64_t)strides_[i] * esz;
    CUresult r = enc(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dfu_set_error("cuTensorMapEncodeTiled(4d) failed (%d): dims=%lldx%lldx%lldx%lld box=%dx%dx%dx%d", (int)r, dims_[0], dims_[1],
                      dims_[2], dims_[3], box_[0], box_[1], box_[2], box_[3]);
        return DFU_ERR_CUDA;
    }
    return DFU_OK;
}

// Pixel box (wb, hb, bb) with wb*hb*bb == total (powers of two) that wastes the fewest tile rows on a W x H x B image
// batch; ties go to the widest box (longest contiguous TMA segments).
void pick_pixel_box(int W, int H, int B, int total, int* wb, int* hb, int* bb) {
    double best = -1.0;
    for (int w = 1; w <= 32 && w <= total; w *= 2)
        for (int h = 1; w * h <= total; h *= 2) {
            const int b = total / (w * h);
            const double covered = (double)dfu_cdiv(W, w) * w * dfu_cdiv(H, h) * h * dfu_cdiv(B, b) * b;
            const double eff = (double)W * H * B / covered + 1e-6 * w + 1e-8 * h;
            if (eff > best) { best = eff; *wb = w; *hb = h; *bb = b; }
        }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int pick_block_n(int N, bool mn_major) {
    const int step = mn_major ? 64 : 32;
    int best = step, best_cost = 1 << 30;
    for (int bn = step; bn <= MAX_BLOCK_N; bn += step) {
        const int tiles = (N + bn - 1) / bn;
        // cost ~ MMA columns issued (padding waste) + a per-tile overhead favouring fewer, larger tiles
        const int cost = tiles * bn + tiles * 24;
        if (cost <= best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

// Split-K factor for an accumulating GEMM: minimise (waves of work items) x (k-blocks per item + fixed per-item cost),
// i.e. account for the wave quantisation of tiles*split over the CTAs instead of just "enough items to fill the GPU".
int pick_split_k(long long out_tiles, int num_kb, int ctas) {
    int best = 1;
    long long best_cost = -1;
    const int max_split = max(1, min(256, num_kb / 4));
    for (int sk = 1; sk <= max_split; ++sk) {
        const int kbs = dfu_cdiv(num_kb, sk);
        const int real = dfu_cdiv(num_kb, kbs);                 // splits that actually get k-blocks
        const long long waves = (out_tiles * real + ctas - 1) / ctas;
        const long long cost = waves * (kbs + 6);                // ~6 k-blocks of pipeline fill + epilogue per item
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sk; }
    }
    return best;
}

// Operand ring: a stage holds the A tile plus this CTA's share of the B tile (block_n / cg rows), and the ring is as deep
// as the shared memory left by the epilogue staging and the column-statistics accumulators allows (narrow column tiles
// get up to 8 stages: their loads are latency-bound, not capacity-bound).
void plan_ring(GemmParams& p, bool aux, long long colacc_floats) {
    p.stage_bytes = A_TILE_BYTES + (p.block_n / p.cg) * BLOCK_K * 2;
    const long long staging = 8ll * p.epi_bufs * (aux ? 2 : 1) * STG_BYTES;
    const long long avail = SMEM_BYTES - 1024 - staging - colacc_floats * 4;
    static const int cap = []() { const char* e = getenv("DFU_GEMM_MAX_STAGES"); return e ? atoi(e) : MAX_STAGES; }();
    p.stages = (int)min((long long)min(cap, MAX_STAGES), avail / p.stage_bytes);
}

bool g_attr_set = false;
// CTA-pair (cta_group::2) kernel for the large GEMMs; DFU_GEMM_PAIR=0 keeps every launch on the single-CTA kernel
const bool g_pair_enabled = []() { const char* e = getenv("DFU_GEMM_PAIR"); return !(e && e[0] == '0'); }();

cudaError_t set_smem_attrs() {
    cudaError_t e = cudaSuccess;
#define DFU_SET_ATTR(CS, EM, CG)                                                                                     \
    if (e == cudaSuccess)                                                                                            \
        e = cudaFuncSetAttribute(dfu_tc_gemm_kernel<CS, EM, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)
    DFU_SET_ATTR(true, 0, 1); DFU_SET_ATTR(true, 1, 1); DFU_SET_ATTR(false, 0, 1); DFU_SET_ATTR(false, 1, 1); DFU_SET_ATTR(false, 2, 1);
    DFU_SET_ATTR(false, 3, 1); DFU_SET_ATTR(false, 4, 1); DFU_SET_ATTR(false, 5, 1);
    DFU_SET_ATTR(false, 1, 2); DFU_SET_ATTR(false, 2, 2); DFU_SET_ATTR(false, 3, 2); DFU_SET_ATTR(false, 4, 2);
    DFU_SET_ATTR(false, 5, 2);
    DFU_SET_ATTR(false, 6, 1); DFU_SET_ATTR(false, 6, 2); DFU_SET_ATTR(true, 6, 1); DFU_SET_ATTR(true, 6, 2);
    DFU_SET_ATTR(false, 7, 1); DFU_SET_ATTR(false, 7, 2);
    DFU_SET_ATTR(false, 8, 1); DFU_SET_ATTR(false, 8, 2);
    DFU_SET_ATTR(false, 9, 1); DFU_SET_ATTR(false, 9, 2); DFU_SET_ATTR(false, 10, 1); DFU_SET_ATTR(false, 10, 2);
#undef DFU_SET_ATTR
    return e;
}

// Which epilogue instantiation serves this launch (-1: the one with the column-statistics code).
int pick_epilogue_mode(const GemmParams& p) {
    if (p.colstats) {    // -1: general epilogue + column statistics; -2: plain bf16 store; -6: bf16 (x) res
        if (p.tma_store && p.alpha == 1.f && !p.d_fp32 && !p.bias && !p.aux) {
            if (!p.res && p.act == 0) return -2;
            if (p.res && p.act == 5) return -6;
        }
        return -1;
    }
    if (!p.tma_store || p.alpha != 1.f) return 0;
    const bool bias = p.bias != nullptr, res = p.res != nullptr, aux = p.aux != nullptr;
    if (p.d_fp32) return (!bias && !res && !aux && p.act == 0) ? 5 : 0;
    if (!bias && !res && !aux && p.act == 0) return 1;
    if (bias && !res && !aux && p.act == 0) return 2;
    if (bias && res && !aux && p.act == 0) return 3;
    if (bias && !res && aux && p.act == 2) return 4;
    if (!bias && res && !aux && p.act == 5) return 6;
    if (bias && !res && aux && p.act == 4) return 7;
    if (!bias && res && !aux && p.act == 0) return 8;      // D = AB + res: the skip gradient joins in the dgrad epilogue
    if (bias && !res && !aux && p.act == 1) return 9;      // eval-mode conv + folded BatchNorm + ReLU
    if (bias && res && !aux && p.act == 6) return 10;      // ... + identity, ReLU after the add
    return 0;
}

template <int EM, bool CS = false>
void launch_pair(int grid, cudaStream_t stream, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                 const CUtensorMap& mx, const GemmParams& p) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = dfu_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    (void)cudaLaunchKernelEx(&cfg, dfu_tc_gemm_kernel<CS, EM, 2>, ma, mb, md, mx, p);
}

void launch_gemm(int em, int grid, cudaStream_t stream, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                 const CUtensorMap& mx, const GemmParams& p) {
    if (p.cg == 2) {
        switch (em) {
            case 1: launch_pair<1>(grid, stream, ma, mb, md, mx, p); return;
            case 2: launch_pair<2>(grid, stream, ma, mb, md, mx, p); return;
            case 3: launch_pair<3>(grid, stream, ma, mb, md, mx, p); return;
            case 4: launch_pair<4>(grid, stream, ma, mb, md, mx, p); return;
            case 6: launch_pair<6>(grid, stream, ma, mb, md, mx, p); return;
            case -6: launch_pair<6, true>(grid, stream, ma, mb, md, mx, p); return;
            case 7: launch_pair<7>(grid, stream, ma, mb, md, mx, p); return;
            case 8: launch_pair<8>(grid, stream, ma, mb, md, mx, p); return;
            case 9: launch_pair<9>(grid, stream, ma, mb, md, mx, p); return;
            case 10: launch_pair<10>(grid, stream, ma, mb, md, mx, p); return;
            default: launch_pair<5>(grid, stream, ma, mb, md, mx, p); return;
        }
    }
    switch (em) {
        case -1: dfu_launch(dfu_tc_gemm_kernel<true, 0, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case -2: dfu_launch(dfu_tc_gemm_kernel<true, 1, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 1: dfu_launch(dfu_tc_gemm_kernel<false, 1, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 2: dfu_launch(dfu_tc_gemm_kernel<false, 2, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 3: dfu_launch(dfu_tc_gemm_kernel<false, 3, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 4: dfu_launch(dfu_tc_gemm_kernel<false, 4, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 5: dfu_launch(dfu_tc_gemm_kernel<false, 5, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 6: dfu_launch(dfu_tc_gemm_kernel<false, 6, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case -6: dfu_launch(dfu_tc_gemm_kernel<true, 6, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 7: dfu_launch(dfu_tc_gemm_kernel<false, 7, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 8: dfu_launch(dfu_tc_gemm_kernel<false, 8, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 9: dfu_launch(dfu_tc_gemm_kernel<false, 9, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        case 10: dfu_launch(dfu_tc_gemm_kernel<false, 10, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
        default: dfu_launch(dfu_tc_gemm_kernel<false, 0, 1>, grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, md, mx, p); break;
    }
}

}  // namespace

extern "C" int dfu_gemm(const dfu_gemm_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DFU_CHECK_ARG(a != nullptr, "dfu_gemm: null args");
    DFU_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0 && a->nb0 > 0 && a->nb1 > 0, "dfu_gemm: bad sizes M=%d N=%d K=%d",
                  a->M, a->N, a->K);
    DFU_CHECK_ARG(a->A && a->B && a->D, "dfu_gemm: null operand");
    DFU_CHECK_ARG(!(a->accumulate && a->d_dtype != 1), "dfu_gemm: accumulate needs fp32 output");
    DFU_CHECK_ARG(!(a->accumulate && (a->aux || a->act || a->res || a->bias || a->colstats)),
                  "dfu_gemm: accumulate mode takes no bias/act/res/aux/colstats");
    DFU_CHECK_ARG(a->act >= 0 && a->act <= 6 && !((a->act == 3 || a->act == 5 || a->act == 6) && !a->res) && !(a->act == 4 && !a->aux),
                  "dfu_gemm: act=%d (3, 5 and 6 need res, 4 needs aux)", a->act);
    DFU_CHECK_ARG(!a->colstats || (a->colstats_mode >= 1 && a->colstats_mode <= 2 && (long long)a->N * a->colstats_mode <= 4096 &&
                                   a->nb0 * a->nb1 == 1 && a->d_dtype == 0),
                  "dfu_gemm: colstats needs mode 1|2, N x mode <= 4096, a bf16 output and no batching");

    const bool batched = (a->nb0 * a->nb1) > 1;
    bool tma_ok = aligned16(a->A) && aligned16(a->B) && (a->lda % 8 == 0) && (a->ldb % 8 == 0);
    if (batched)
        tma_ok = tma_ok && (a->nb0 == 1 || (a->sa0 % 8 == 0 && a->sb0 % 8 == 0)) &&
                 (a->nb1 == 1 || (a->sa1 % 8 == 0 && a->sb1 % 8 == 0));
    const bool tiny = a->N < 16;
    int impl = a->impl;
    if (impl == 0 && (!tma_ok || tiny)) impl = 1;   // shape-specialised SIMT kernel, same semantics
    DFU_CHECK_ARG(!(impl == 1 && a->colstats), "dfu_gemm: colstats is only implemented by the tcgen05 kernel");

    if (impl == 1) {
        RefParams p;
        p.M = a->M; p.N = a->N; p.K = a->K; p.nb0 = a->nb0; p.nb1 = a->nb1;
        p.A = (const bf16*)a->A; p.a_mn = a->a_major; p.lda = a->lda; p.sa0 = a->sa0; p.sa1 = a->sa1;
        p.B = (const bf16*)a->B; p.b_mn = a->b_major; p.ldb = a->ldb; p.sb0 = a->sb0; p.sb1 = a->sb1;
        p.D = a->D; p.d_fp32 = a->d_dtype; p.atomic = a->accumulate; p.ldd = a->ldd; p.sd0 = a->sd0; p.sd1 = a->sd1;
        p.aux = (bf16*)a->aux; p.bias = a->bias; p.res = (const bf16*)a->res; p.ldr = a->ldr; p.sr0 = a->sr0;
        p.sr1 = a->sr1; p.alpha = a->alpha; p.act = a->act;
        dim3 grid(dfu_cdiv(a->N, 16), dfu_cdiv(a->M, 16), a->nb0 * a->nb1);
        DFU_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "dfu_gemm(simt): grid too large (M=%d, batch=%d)", a->M,
                      a->nb0 * a->nb1);
        dfu_launch(dfu_ref_gemm_kernel, grid, 256, 0, stream, p);
        DFU_CHECK_LAUNCH("dfu_ref_gemm_kernel");
        return DFU_OK;
    }

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = a->M; p.N = a->N; p.K = a->K; p.nb0 = a->nb0; p.nb1 = a->nb1;
    p.a_mn = a->a_major ? 1 : 0;
    p.b_mn = a->b_major ? 1 : 0;
    const int esz = a->d_dtype ? 4 : 2;
    bool d_tma = aligned16(a->D) && ((a->ldd * esz) % 16 == 0) && (a->nb0 == 1 || (a->sd0 * esz) % 16 == 0) &&
                 (a->nb1 == 1 || (a->sd1 * esz) % 16 == 0);
    if (a->aux) d_tma = d_tma && aligned16(a->aux) && !a->d_dtype;
    p.tma_store = d_tma ? 1 : 0;
    // Shared memory plan: 4 ring stages + one staging set per epilogue warp, or 3 stages + 64 KB of staging when the
    // epilogue needs a second slab (aux) or the per-column accumulators (colstats).
    p.stages = STAGES;
    p.epi_bufs = 1;
    // (measured: a second staging set does not speed up the short-K shapes -- l1 1x1 77 -> 79 us -- and the lost ring
    // stage slows the K=768 ones, so it is only taken when the third stage is given up anyway)
    if (p.tma_store && (a->aux || a->colstats)) {
        p.stages = 3;
        p.epi_bufs = a->aux ? 1 : 2;     // an aux slab doubles the set size: 3 x 48 KB + 8 x 2 x 4 KB either way
    }
    p.colstats = a->colstats; p.colstats_mode = a->colstats_mode;
    if (a->colstats && !p.tma_store) {
        dfu_set_error("dfu_gemm: colstats needs a 16-byte aligned output (TMA-store epilogue)");
        return DFU_ERR_UNSUPPORTED;
    }
    // bf16 TMA-store slabs are 64 columns wide, so block_n must be a multiple of 64 there
    p.block_n = pick_block_n(a->N, p.b_mn || (d_tma && !a->d_dtype));
    p.stage_bytes = STAGE_BYTES;
    p.cg = 1;
    {
        // CTA-pair kernel for the big unbatched shapes whose N tiles are full 256-wide ones
        GemmParams probe = p;
        probe.tma_store = p.tma_store; probe.alpha = a->alpha; probe.bias = a->bias; probe.res = (const bf16*)a->res;
        probe.aux = (bf16*)a->aux; probe.act = a->act; probe.d_fp32 = a->d_dtype; probe.colstats = a->colstats;
        const int em = pick_epilogue_mode(probe);
        if (g_pair_enabled && (em >= 1 || em == -6) && !batched && p.block_n == 256 && a->N % 256 == 0 && a->M >= 512 && a->K >= 256 &&
            a->max_ctas == 0) {
            // a pair finishes a 256-row tile in ~0.88 of the time two CTAs need for two 128-row tiles (measured), but it
            // also halves the number of schedulable units: take it only when the wave count does not eat the gain
            // (split-K launches choose their split for whatever unit count they get)
            const int ctas = dfu_num_sms();
            const long long nt = dfu_cdiv(a->N, 256);
            const long long w1 = (dfu_cdiv(a->M, BLOCK_M) * nt + ctas - 1) / ctas;
            const long long w2 = (dfu_cdiv(a->M, 2 * BLOCK_M) * nt + ctas / 2 - 1) / (ctas / 2);
            if (a->accumulate || w2 * 88 < w1 * 100) p.cg = 2;
        }
    }
    if (p.cg == 2) p.epi_bufs = 1;
    // block_n = 256: 4 x 48 KB stages (3 next to aux / colstats staging); a CTA pair: 6 x 32 KB (5); narrower tiles deeper
    plan_ring(p, a->aux != nullptr, a->colstats ? (long long)a->N * a->colstats_mode : 0);
    p.m_tiles = dfu_cdiv(a->M, BLOCK_M * p.cg);
    p.n_tiles = dfu_cdiv(a->N, p.block_n);
    p.num_kb = dfu_cdiv(a->K, BLOCK_K);
    const int num_sms = dfu_num_sms();
    const int max_ctas = a->max_ctas > 0 ? a->max_ctas : num_sms;
    const long long out_tiles = (long long)p.m_tiles * p.n_tiles * a->nb0 * a->nb1;
    int split = 1;
    if (a->accumulate) split = pick_split_k(out_tiles, p.num_kb, max_ctas / p.cg);
    p.kb_per_split = dfu_cdiv(p.num_kb, split);
    p.split_k = dfu_cdiv(p.num_kb, p.kb_per_split);
    p.total_work = out_tiles * p.split_k;
    DFU_CHECK_ARG(p.total_work < (1ll << 31) - 4096, "dfu_gemm: %lld work items (32-bit tile index)", p.total_work);
    p.D = a->D; p.ldd = a->ldd; p.sd0 = a->sd0; p.sd1 = a->sd1;
    p.d_fp32 = a->d_dtype; p.atomic = a->accumulate ? (p.split_k > 1 ? 2 : 1) : 0;
    p.aux = (bf16*)a->aux; p.bias = a->bias; p.res = (const bf16*)a->res;
    p.ldr = a->ldr; p.sr0 = a->sr0; p.sr1 = a->sr1;
    p.alpha = a->alpha; p.act = a->act;
    bool vec = aligned16(a->D) && ((a->ldd * esz) % 16 == 0) && ((a->sd0 * esz) % 16 == 0) && ((a->sd1 * esz) % 16 == 0);
    if (a->aux) vec = vec && aligned16(a->aux) && (a->ldd % 8 == 0) && (a->sd0 % 8 == 0) && (a->sd1 % 8 == 0);
    if (a->res) vec = vec && aligned16(a->res) && (a->ldr % 8 == 0) && (a->sr0 % 8 == 0) && (a->sr1 % 8 == 0);
    p.vec_ok = vec ? 1 : 0;
    p.res_vec = (a->res && aligned16(a->res) && (a->ldr % 8 == 0) && (a->sr0 % 8 == 0) && (a->sr1 % 8 == 0)) ? 1 : 0;

    CUtensorMap map_a, map_b;
    int rc;
    if (!p.a_mn) rc = make_map(&map_a, a->A, a->K, a->M, a->lda, a->nb0, a->sa0, a->nb1, a->sa1, BLOCK_K, BLOCK_M);
    else         rc = make_map(&map_a, a->A, a->M, a->K, a->lda, a->nb0, a->sa0, a->nb1, a->sa1, 64, BLOCK_K);
    if (rc) return rc;
    if (!p.b_mn) rc = make_map(&map_b, a->B, a->K, a->N, a->ldb, a->nb0, a->sb0, a->nb1, a->sb1, BLOCK_K, p.block_n / p.cg);
    else         rc = make_map(&map_b, a->B, a->N, a->K, a->ldb, a->nb0, a->sb0, a->nb1, a->sb1, 64, BLOCK_K);
    if (rc) return rc;

    CUtensorMap map_d, map_x;
    memset(&map_d, 0, sizeof(map_d));
    memset(&map_x, 0, sizeof(map_x));
    if (p.tma_store) {
        const int slab = a->d_dtype ? 32 : 64;
        rc = make_map(&map_d, a->D, a->N, a->M, a->ldd, a->nb0, a->sd0, a->nb1, a->sd1, slab, 32, esz);
        if (rc) return rc;
        if (a->aux) {
            rc = make_map(&map_x, a->aux, a->N, a->M, a->ldd, a->nb0, a->sd0, a->nb1, a->sd1, 64, 32, 2);
            if (rc) return rc;
        }
    }
    if (!g_attr_set) {
        cudaError_t e = set_smem_attrs();
        if (e != cudaSuccess) {
            dfu_set_error("dfu_gemm: cannot reserve %d B shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return DFU_ERR_CUDA;
        }
        g_attr_set = true;
    }
    int units = (int)min((long long)(max_ctas / p.cg), p.total_work);
    // column statistics live in registers while a CTA stays on one column tile: with the column tile the fastest-varying
    // index of the work order that holds for the whole launch when the number of CTAs is a multiple of n_tiles
    if (p.colstats && p.n_tiles > 1 && units > p.n_tiles) units -= units % p.n_tiles;
    const int grid = p.cg * units;
    launch_gemm(pick_epilogue_mode(p), grid, stream, map_a, map_b, map_d, map_x, p);
    DFU_CHECK_LAUNCH("dfu_tc_gemm_kernel");
    return DFU_OK;
}


// ----------------------------------------------------------------------------- implicit-GEMM convolution
namespace {
int conv_gemm_launch(const dfu_conv_args* a, cudaStream_t stream, int ph, int pw);
}

extern "C" int dfu_conv_gemm(const dfu_conv_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DFU_CHECK_ARG(a != nullptr, "dfu_conv_gemm: null args");
    if (a->mode == 2 && a->stride == 2) {
        // input gradient of a stride-2 convolution: the dx pixels of one parity class (h % 2, w % 2) see a fixed subset of
        // the filter taps at unit-stride dy offsets, so each class is an ordinary implicit GEMM whose output view is dx
        // with doubled pixel strides.  Classes without taps (1x1 filters: every odd row / column) are zero.
        bool any_empty = false;
        for (int c = 0; c < 4; ++c) {
            const int r0 = ((c >> 1) + a->pad) & 1, s0 = ((c & 1) + a->pad) & 1;
            if (r0 >= a->R || s0 >= a->S) any_empty = true;
        }
        if (any_empty) {
            cudaError_t e = cudaMemsetAsync(a->out, 0, (size_t)a->B * a->H * a->W * a->C * 2, stream);
            if (e != cudaSuccess) { dfu_set_error("dfu_conv_gemm: memset failed: %s", cudaGetErrorString(e)); return DFU_ERR_CUDA; }
        }
        for (int c = 0; c < 4; ++c) {
            const int ph = c >> 1, pw = c & 1;
            const int r0 = (ph + a->pad) & 1, s0 = (pw + a->pad) & 1;
            if (r0 >= a->R || s0 >= a->S || ph >= a->H || pw >= a->W) continue;
            const int rc = conv_gemm_launch(a, stream, ph, pw);
            if (rc) return rc;
        }
        return DFU_OK;
    }
    return conv_gemm_launch(a, stream, 0, 0);
}

namespace {
int conv_gemm_launch(const dfu_conv_args* a, cudaStream_t stream, int ph, int pw) {
    DFU_CHECK_ARG(a != nullptr, "dfu_conv_gemm: null args");
    DFU_CHECK_ARG(a->mode >= 1 && a->mode <= 3, "dfu_conv_gemm: mode %d", a->mode);
    DFU_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0 && a->R > 0 && a->S > 0 && a->pad >= 0, "dfu_conv_gemm: bad geometry");
    if ((a->stride != 1 && a->stride != 2) || a->C % 64 != 0 || a->Cout % 64 != 0) {
        dfu_set_error("dfu_conv_gemm: needs stride 1 or 2 and channel counts that are multiples of 64 (C=%d, Cout=%d, stride=%d)",
                      a->C, a->Cout, a->stride);
        return DFU_ERR_UNSUPPORTED;
    }
    const int Ho = (a->H + 2 * a->pad - a->R) / a->stride + 1, Wo = (a->W + 2 * a->pad - a->S) / a->stride + 1;
    DFU_CHECK_ARG(Ho > 0 && Wo > 0, "dfu_conv_gemm: empty output");
    DFU_CHECK_ARG(a->out && (a->mode == 2 || a->x) && (a->mode == 3 || a->w) && (a->mode == 1 || a->dy),
                  "dfu_conv_gemm: null operand for mode %d", a->mode);
    DFU_CHECK_ARG(aligned16(a->out) && aligned16(a->x) && aligned16(a->w) && aligned16(a->dy), "dfu_conv_gemm: unaligned operand");
    DFU_CHECK_ARG(a->mode == 3 || a->ldw % 8 == 0, "dfu_conv_gemm: weight row stride must be a multiple of 8");
    DFU_CHECK_ARG(a->mode != 3 || a->ldo % 4 == 0, "dfu_conv_gemm: dW row stride must be a multiple of 4");

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.nb0 = p.nb1 = 1;
    p.alpha = 1.f;
    p.tma_store = 1;
    p.stages = STAGES;
    p.stage_bytes = STAGE_BYTES;
    p.cg = 1;
    p.epi_bufs = 1;
    p.vec_ok = 1;
    p.cv_mode = a->mode;
    if (a->colstats) {
        DFU_CHECK_ARG(a->mode == 1 && a->Cout <= 2048, "dfu_conv_gemm: colstats is a forward-mode feature (Cout <= 2048)");
        p.colstats = a->colstats;
        p.colstats_mode = 2;
    }
    if (a->bias || a->act) {
        DFU_CHECK_ARG(a->mode == 1 && (a->act == 0 || a->act == 1) && !a->colstats,
                      "dfu_conv_gemm: bias / ReLU belong to the forward mode (act 0|1) and exclude colstats");
        p.bias = a->bias;
        p.act = a->act;
    }
    p.cv_S = a->S;
    p.cv_pad = a->pad;
    p.cv_C = a->C;
    p.cv_stride = a->mode == 2 ? 1 : a->stride;      // dgrad gathers dy at unit stride (per parity class at stride 2)
    const int taps = a->R * a->S;
    const int num_sms = dfu_num_sms();
    const int max_ctas = a->max_ctas > 0 ? a->max_ctas : num_sms;

    const long long xd[4] = {a->C, a->W, a->H, a->B};
    const long long xs[3] = {a->C, (long long)a->W * a->C, (long long)a->H * a->W * a->C};
    const long long yd[4] = {a->Cout, Wo, Ho, a->B};
    const long long ys[3] = {a->Cout, (long long)Wo * a->Cout, (long long)Ho * Wo * a->Cout};
    CUtensorMap map_a, map_b, map_d, map_x;
    memset(&map_x, 0, sizeof(map_x));
    int rc;
    if (a->mode == 1 || a->mode == 2) {
        const bool fwd = a->mode == 1;
        const int tstep = fwd ? 1 : a->stride;
        // image the output pixels live in (dgrad at stride 2: the dx pixels of parity class (ph, pw))
        const int OW = fwd ? Wo : (a->W - pw + tstep - 1) / tstep, OH = fwd ? Ho : (a->H - ph + tstep - 1) / tstep;
        p.cv_tstep = tstep; p.cv_ph = ph; p.cv_pw = pw;
        p.cv_r0 = fwd ? 0 : (ph + a->pad) % tstep; p.cv_s0 = fwd ? 0 : (pw + a->pad) % tstep;
        p.cv_nr = fwd ? a->R : (a->R - p.cv_r0 + tstep - 1) / tstep;
        p.cv_ns = fwd ? a->S : (a->S - p.cv_s0 + tstep - 1) / tstep;
        const int ntaps = p.cv_nr * p.cv_ns;
        pick_pixel_box(OW, OH, a->B, BLOCK_M, &p.cv_wb, &p.cv_hb, &p.cv_bb);
        for (p.cv_lw = 0; (1 << p.cv_lw) < p.cv_wb; ++p.cv_lw) {}
        for (p.cv_lh = 0; (1 << p.cv_lh) < p.cv_hb; ++p.cv_lh) {}
        p.cv_tw = dfu_cdiv(OW, p.cv_wb);
        p.cv_th = dfu_cdiv(OH, p.cv_hb);
        p.cv_ow = OW; p.cv_oh = OH; p.cv_nb = a->B;
        p.m_tiles = p.cv_tw * p.cv_th * dfu_cdiv(a->B, p.cv_bb);
        p.M = p.m_tiles * BLOCK_M;
        p.N = fwd ? a->Cout : a->C;
        const int gathered_c = fwd ? a->C : a->Cout;
        p.K = ntaps * gathered_c;
        p.cv_cchunks = gathered_c / 64;
        p.num_kb = ntaps * p.cv_cchunks;
        p.a_mn = 0;
        p.b_mn = fwd ? 0 : 1;
        p.block_n = pick_block_n(p.N, true);
        plan_ring(p, false, p.colstats ? 2ll * p.N : 0);
        p.n_tiles = dfu_cdiv(p.N, p.block_n);
        p.split_k = 1;
        p.kb_per_split = p.num_kb;
        const int rem = 32 / p.cv_wb;
        p.cv_sh = p.cv_hb < rem ? p.cv_hb : rem;
        p.cv_sb = rem / p.cv_sh;
        const int walk = fwd ? a->stride : 1;                     // forward: the gathered box steps through x
        const int abox[4] = {64, p.cv_wb * walk, p.cv_hb * walk, p.cv_bb};
        const int dbox[4] = {64, p.cv_wb, p.cv_sh, p.cv_sb};
        rc = make_map_nd(&map_a, fwd ? a->x : a->dy, fwd ? xd : yd, fwd ? xs : ys, abox, 2, walk);
        if (rc) return rc;
        if (fwd) rc = make_map(&map_b, a->w, p.K, p.N, a->ldw, 1, 0, 1, 0, BLOCK_K, p.block_n);
        else     rc = make_map(&map_b, a->w, (long long)taps * a->C, a->Cout, a->ldw, 1, 0, 1, 0, 64, BLOCK_K);
        if (rc) return rc;
        if (fwd || tstep == 1) {
            rc = make_map_nd(&map_d, a->out, fwd ? yd : xd, fwd ? ys : xs, dbox);
        } else {
            // output view = the (ph, pw) sub-grid of dx: pixel strides doubled, origin at pixel (ph, pw)
            const long long vd[4] = {a->C, OW, OH, a->B};
            const long long vs[3] = {(long long)tstep * a->C, (long long)tstep * a->W * a->C, (long long)a->H * a->W * a->C};
            rc = make_map_nd(&map_d, (const bf16*)a->out + ((long long)ph * a->W + pw) * a->C, vd, vs, dbox);
        }
        if (rc) return rc;
        p.D = a->out;
        p.d_fp32 = 0;
    } else {
        pick_pixel_box(Wo, Ho, a->B, BLOCK_K, &p.cv_wb, &p.cv_hb, &p.cv_bb);
        for (p.cv_lw = 0; (1 << p.cv_lw) < p.cv_wb; ++p.cv_lw) {}
        for (p.cv_lh = 0; (1 << p.cv_lh) < p.cv_hb; ++p.cv_lh) {}
        p.cv_tw = dfu_cdiv(Wo, p.cv_wb);
        p.cv_th = dfu_cdiv(Ho, p.cv_hb);
        p.num_kb = p.cv_tw * p.cv_th * dfu_cdiv(a->B, p.cv_bb);
        p.M = a->Cout;
        p.N = taps * a->C;
        p.K = p.num_kb * BLOCK_K;
        p.a_mn = p.b_mn = 1;
        p.block_n = pick_block_n(p.N, true);
        plan_ring(p, false, p.colstats ? 2ll * p.N : 0);
        p.m_tiles = dfu_cdiv(p.M, BLOCK_M);
        p.n_tiles = dfu_cdiv(p.N, p.block_n);
        const long long out_tiles = (long long)p.m_tiles * p.n_tiles;
        int split = 1;
        if (a->accumulate) split = pick_split_k(out_tiles, p.num_kb, max_ctas);
        p.kb_per_split = dfu_cdiv(p.num_kb, split);
        p.split_k = dfu_cdiv(p.num_kb, p.kb_per_split);
        p.atomic = a->accumulate ? (p.split_k > 1 ? 2 : 1) : 0;
        const int box[4] = {64, p.cv_wb, p.cv_hb, p.cv_bb};
        const int xbox[4] = {64, p.cv_wb * a->stride, p.cv_hb * a->stride, p.cv_bb};
        rc = make_map_nd(&map_a, a->dy, yd, ys, box);
        if (rc) return rc;
        rc = make_map_nd(&map_b, a->x, xd, xs, xbox, 2, a->stride);
        if (rc) return rc;
        rc = make_map(&map_d, a->out, p.N, p.M, a->ldo, 1, 0, 1, 0, 32, 32, 4);
        if (rc) return rc;
        p.D = a->out;
        p.ldd = a->ldo;
        p.d_fp32 = 1;
    }
    p.total_work = (long long)p.m_tiles * p.n_tiles * p.split_k;
    if (!g_attr_set) {
        cudaError_t e = set_smem_attrs();
        if (e != cudaSuccess) {
            dfu_set_error("dfu_conv_gemm: cannot reserve %d B shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return DFU_ERR_CUDA;
        }
        g_attr_set = true;
    }
    int grid = (int)min((long long)max_ctas, p.total_work);
    if (p.colstats && p.n_tiles > 1 && grid > p.n_tiles) grid -= grid % p.n_tiles;     // see dfu_gemm
    launch_gemm(pick_epilogue_mode(p), grid, stream, map_a, map_b, map_d, map_x, p);
    DFU_CHECK_LAUNCH("dfu_tc_gemm_kernel(conv)");
    return DFU_OK;
}
}  // namespace
