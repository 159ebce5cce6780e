// BatchNorm2d (training + eval, forward + backward) over NHWC bf16 [M, C] matrices, LayerNorm over rows,
// and row softmax.  All are HBM-bound: 128-bit accesses along C / D, fp32 statistics, warp-shuffle reductions.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/dfu_b200.h"
#include <stdlib.h>

namespace {

inline int grid_cap(long long blocks, int per_sm) {
    long long cap = (long long)dfu_num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ------------------------------------------------------------------------------------------ BatchNorm
constexpr int BN_UNROLL = 4;
// Thread layout: 256 threads = (256/cvec) row lanes x cvec channel vectors (8 channels each).
// MODE 0: sums of y and y^2.  MODE 1 (backward): sums of dpre and dpre*xhat.
template <int MODE>
__global__ void __launch_bounds__(256, 2) bn_reduce_kernel(const bf16* __restrict__ a, const bf16* __restrict__ z,
                                                        const bf16* __restrict__ y, const float* __restrict__ mean_invstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ sums, long long M, int C, int relu,
                                                        const unsigned char* __restrict__ zmask) {
    pdl_sync();
    __shared__ float red[2][256][9];   // padded against bank conflicts
    const int cvec = C >> 3;
    const int rows_per_iter = 256 / cvec;
    const int cv = threadIdx.x % cvec;
    const int rl = threadIdx.x / cvec;
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    // xhat = xa*y + xb;  mask (no saved z): mp*y + mq > 0
    float xa[8], xb[8], mp[8], mq[8];
    const bool mask_from_y = MODE == 1 && relu && z == nullptr && zmask == nullptr;
    const bool use_bits = MODE == 1 && relu && zmask != nullptr;     // 1 bit per element saved by the forward pass
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float mean = mean_invstd[cv * 8 + i], invstd = mean_invstd[C + cv * 8 + i];
            xa[i] = invstd;
            xb[i] = -mean * invstd;
            mp[i] = mask_from_y ? gamma[cv * 8 + i] * invstd : 0.f;
            mq[i] = mask_from_y ? beta[cv * 8 + i] - gamma[cv * 8 + i] * invstd * mean : 1.f;
        }
    }
    // BN_UNROLL rows per thread per trip: all 128-bit loads are issued before any is consumed (memory-level parallelism)
    const long long rstride = (long long)gridDim.x * rows_per_iter;
    for (long long r0 = (long long)blockIdx.x * rows_per_iter + rl; r0 < M; r0 += rstride * BN_UNROLL) {
        bf16x8 va[BN_UNROLL], vy[BN_UNROLL], vz[BN_UNROLL];
        unsigned int mb[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * rstride;
            if (r < M) {
                const long long off = r * C + cv * 8;
                va[u] = ld8(a + off);
                if (MODE == 1) {
                    vy[u] = ld8(y + off);
                    if (use_bits) mb[u] = zmask[r * cvec + cv];
                    else if (relu && !mask_from_y) vz[u] = ld8(z + off);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * rstride;
            if (r >= M) break;
            float f[8];
            unpack8(va[u], f);
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { s1[i] += f[i]; s2[i] += f[i] * f[i]; }
            } else {
                float zz[8], yy[8];
                unpack8(vy[u], yy);
                if (use_bits) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (!((mb[u] >> i) & 1u)) f[i] = 0.f;
                } else if (relu && !mask_from_y) {
                    unpack8(vz[u], zz);
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (!(zz[i] > 0.f)) f[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (mask_from_y && !(fmaf(mp[i], yy[i], mq[i]) > 0.f)) f[i] = 0.f;
                    s1[i] += f[i];
                    s2[i] = fmaf(f[i], fmaf(xa[i], yy[i], xb[i]), s2[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[0][threadIdx.x][i] = s1[i]; red[1][threadIdx.x][i] = s2[i]; }
    __syncthreads();
    // threads 0..cvec*8-1 each finish one channel (loop if C > 256)
    for (int c = threadIdx.x; c < C; c += 256) {
        const int v = c >> 3, i = c & 7;
        float t1 = 0.f, t2 = 0.f;
        for (int k = 0; k < rows_per_iter; ++k) { t1 += red[0][k * cvec + v][i]; t2 += red[1][k * cvec + v][i]; }
        atomicAdd(sums + c, t1);
        atomicAdd(sums + C + c, t2);
    }
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, float* __restrict__ mean_invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, long long M, int C, float eps, float momentum) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= C) return;
    const float inv_m = 1.f / (float)M;
    const float mean = sums[c] * inv_m;
    float var = sums[C + c] * inv_m - mean * mean;
    var = fmaxf(var, 0.f);
    mean_invstd[c] = mean;
    mean_invstd[C + c] = rsqrtf(var + eps);
    if (running_mean) {
        const float unbiased = (M > 1) ? var * ((float)M / (float)(M - 1)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv,
                                     float* __restrict__ mean_invstd, int C, float eps) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean_invstd[c] = rm[c];
    mean_invstd[C + c] = rsqrtf(rv[c] + eps);
}

// `sums` non-null: training-mode call that also does the finalize step -- every thread derives mean / invstd of its 8
// channels from the batch sums, block 0 stores them (for the backward pass) and updates the running statistics.
__global__ void __launch_bounds__(256) bn_apply_kernel(const bf16* __restrict__ y, float* __restrict__ mean_invstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const bf16* __restrict__ res, bf16* __restrict__ z, long long M,
                                                       int C, int relu, const float* __restrict__ sums,
                                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                                       long long* __restrict__ nbt, float eps, float momentum,
                                                       unsigned char* __restrict__ zmask) {
    pdl_sync();
    const int cvec = C >> 3;
    const int rows_per_iter = 256 / cvec;
    const int cv = threadIdx.x % cvec;
    const int rl = threadIdx.x / cvec;
    float sc[8], sh[8];
    if (sums) {
        const float inv_m = 1.f / (float)M;
        if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cv * 8 + i;
            const float mean = sums[c] * inv_m;
            const float var = fmaxf(sums[C + c] * inv_m - mean * mean, 0.f);
            const float invstd = rsqrtf(var + eps);
            if (blockIdx.x == 0 && rl == 0) {
                mean_invstd[c] = mean;
                mean_invstd[C + c] = invstd;
                if (running_mean) {
                    const float unbiased = (M > 1) ? var * ((float)M / (float)(M - 1)) : var;
                    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
                }
            }
            sc[i] = gamma[c] * invstd;
            sh[i] = beta[c] - mean * sc[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cv * 8 + i;
            sc[i] = gamma[c] * mean_invstd[C + c];
            sh[i] = beta[c] - mean_invstd[c] * sc[i];
        }
    }
    const long long rstride = (long long)gridDim.x * rows_per_iter;
    for (long long r0 = (long long)blockIdx.x * rows_per_iter + rl; r0 < M; r0 += rstride * BN_UNROLL) {
        bf16x8 vy[BN_UNROLL], vr[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * rstride;
            if (r < M) {
                const long long off = r * C + cv * 8;
                vy[u] = ld8(y + off);
                if (res) vr[u] = ld8(res + off);
            }
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * rstride;
            if (r >= M) break;
            float f[8];
            unpack8(vy[u], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = f[i] * sc[i] + sh[i];
            if (res) {
                float rr[8];
                unpack8(vr[u], rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += rr[i];
            }
            if (relu) {
                if (zmask) {                 // ReLU mask for the backward pass: 1 bit per element instead of re-reading z
                    unsigned int bits = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) bits |= (f[i] > 0.f ? 1u : 0u) << i;
                    zmask[r * cvec + cv] = (unsigned char)bits;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            st8(z + r * C + cv * 8, pack8(f));
        }
    }
}

__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(const bf16* __restrict__ dz, const bf16* __restrict__ z,
                                                              const bf16* __restrict__ y, const float* __restrict__ mean_invstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ sums,
                                                              bf16* __restrict__ dy, bf16* __restrict__ dres,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              long long M, int C, int relu, int train,
                                                              const unsigned char* __restrict__ zmask) {
    pdl_sync();
    const int cvec = C >> 3;
    const int rows_per_iter = 256 / cvec;
    const int cv = threadIdx.x % cvec;
    const int rl = threadIdx.x / cvec;
    // Per channel everything is folded into five coefficients (keeps the kernel at 2 blocks/SM without spilling the
    // in-flight loads):  dy = ca*d + cb*y + cc  with d = dz masked;  mask (no saved z): cp*y + cq > 0.
    float ca[8], cb[8], cc[8], cp[8], cq[8];
    const float inv_m = 1.f / (float)M;
    const bool use_bits = relu && zmask != nullptr;
    const bool mask_from_y = relu && z == nullptr && !use_bits;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        const float mean = mean_invstd[c], invstd = mean_invstd[C + c];
        const float g = gamma[c] * invstd;
        const float k1 = train ? sums[c] * inv_m : 0.f;
        const float k2 = train ? sums[C + c] * inv_m : 0.f;
        ca[i] = g;
        cb[i] = -g * k2 * invstd;
        cc[i] = g * (k2 * invstd * mean - k1);
        cp[i] = mask_from_y ? gamma[c] * invstd : 0.f;
        cq[i] = mask_from_y ? beta[c] - gamma[c] * invstd * mean : 1.f;
    }
    if (blockIdx.x == 0 && rl == 0 && dgamma) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dbeta[cv * 8 + i] = sums[cv * 8 + i];
            dgamma[cv * 8 + i] = sums[C + cv * 8 + i];
        }
    }
    const bool use_z = relu && !mask_from_y && !use_bits;
    const long long rstride = (long long)gridDim.x * rows_per_iter;
    for (long long r0 = (long long)blockIdx.x * rows_per_iter + rl; r0 < M; r0 += rstride * BN_UNROLL) {
        bf16x8 vd[BN_UNROLL], vy[BN_UNROLL], vz[BN_UNROLL];
        unsigned int mb[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * rstride;
            if (r < M) {
                const long long off = r * C + cv * 8;
                vd[u] = ld8(dz + off);
                vy[u] = ld8(y + off);
                if (use_z) vz[u] = ld8(z + off);
                if (use_bits) mb[u] = zmask[r * cvec + cv];
            }
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * rstride;
            if (r >= M) break;
            const long long off = r * C + cv * 8;
            float d[8], yy[8];
            unpack8(vd[u], d);
            unpack8(vy[u], yy);
            if (use_bits) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (!((mb[u] >> i) & 1u)) d[i] = 0.f;
            } else if (use_z) {
                float zz[8];
                unpack8(vz[u], zz);
#pragma unroll
                for (int i = 0; i < 8; ++i) if (!(zz[i] > 0.f)) d[i] = 0.f;
            } else if (mask_from_y) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (!(fmaf(cp[i], yy[i], cq[i]) > 0.f)) d[i] = 0.f;
            }
            if (dres) st8(dres + off, pack8(d));
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaf(ca[i], d[i], fmaf(cb[i], yy[i], cc[i]));
            st8(dy + off, pack8(o));
        }
    }
}

// ------------------------------------------------------------------------------------------ streaming variants
// ncu on the kernels above (profiles/r02b_ncu_full_bandwidth_kernels.txt): 24 % occupancy, issue slots a third busy, DRAM at
// 50-62 %: they are bound by the bytes a block can keep in flight out of its registers.  The streaming variants take the
// loads out of the register file: each block walks a CONTIGUOUS range of 4 KB chunks (256 vectors of 8 channels: the channel
// vector of a thread stays fixed, 256 being a multiple of C/8) and one thread keeps ST_DEPTH chunks per input tensor in
// flight with cp.async.bulk into a shared-memory ring (mbarrier transaction counts); the 256 threads consume a chunk with
// one ld.shared.v4 each.  In flight per SM: blocks x ST_DEPTH x 4 KB per tensor, independent of registers.
// Measured (profiles/r02k_bn_stream_sweep.txt, r02l_kernel_breakdown_*): bn_apply 20-30 % faster at every layer size (1.85 -> 1.49 ms
// per step).  The same ring under the two BACKWARD kernels was slower in the step (reduce 1.45 -> 2.11 ms, apply 2.26 -> 2.45 ms:
// their per-chunk arithmetic no longer overlaps the loads at 2-3 blocks per SM) and was taken out again.
constexpr int ST_DEPTH = 6;
constexpr int ST_CHUNK = 256 * 16;           // bytes of one tensor chunk

__device__ __forceinline__ bf16x8 lds8(const void* p) {
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(tc::smem_u32(p)));
    return *reinterpret_cast<bf16x8*>(&u);
}

// NT: number of streamed input tensors (1: y; 2: y + res, or dz + y).  issue(chunk, slot) is called by thread 0.
template <int NT>
struct StreamRing {
    uint8_t* smem;                // [ST_DEPTH][NT][ST_CHUNK]
    uint64_t* full;               // [ST_DEPTH]
    const uint8_t* src[NT];
    long long nvec, c_first, c_step, c_count;      // this block's chunks: c_first + i * c_step, i < c_count
    __device__ __forceinline__ long long chunk(long long i) const { return c_first + i * c_step; }
    // interleaved = 1: at any moment the blocks read neighbouring chunks (a front moving through the tensor);
    // interleaved = 0: every block owns one contiguous range
    __device__ __forceinline__ void plan(long long nvec_, int interleaved) {
        nvec = nvec_;
        const long long nchunks = (nvec + 255) / 256;
        if (interleaved) {
            c_first = blockIdx.x;
            c_step = gridDim.x;
            c_count = nchunks > (long long)blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        } else {
            const long long per = (nchunks + gridDim.x - 1) / gridDim.x;
            c_first = min(nchunks, per * blockIdx.x);
            c_step = 1;
            c_count = min(nchunks, c_first + per) - c_first;
        }
    }
    __device__ __forceinline__ void issue(long long c, int slot) {
        const long long v0 = c * 256;
        const uint32_t bytes = (uint32_t)(min(256ll, nvec - v0) * 16);
        tc::mbar_expect_tx(&full[slot], bytes * NT);
#pragma unroll
        for (int t = 0; t < NT; ++t) tc::bulk_load_1d(smem + ((size_t)slot * NT + t) * ST_CHUNK, src[t] + v0 * 16, bytes, &full[slot]);
    }
    __device__ __forceinline__ void start() {
        if (threadIdx.x == 0) {
            for (int s = 0; s < ST_DEPTH; ++s) tc::mbar_init(&full[s], 1);
            tc::fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int s = 0; s < ST_DEPTH && s < c_count; ++s) issue(chunk(s), s);
    }
    // data of chunk number i (0-based within this block's range) for this thread, tensor t
    __device__ __forceinline__ const void* slot_ptr(long long i, int t) const {
        return smem + ((size_t)(i % ST_DEPTH) * NT + t) * ST_CHUNK + threadIdx.x * 16;
    }
    __device__ __forceinline__ void wait(long long i) { tc::mbar_wait(&full[i % ST_DEPTH], (uint32_t)((i / ST_DEPTH) & 1)); }
    // all threads have read chunk i out of shared memory: its slot takes chunk i + ST_DEPTH
    __device__ __forceinline__ void release(long long i) {
        __syncthreads();
        if (threadIdx.x == 0 && i + ST_DEPTH < c_count) issue(chunk(i + ST_DEPTH), (int)(i % ST_DEPTH));
    }
};

template <bool HAS_RES>
__global__ void __launch_bounds__(256) bn_apply_stream_kernel(const bf16* __restrict__ y, float* __restrict__ mean_invstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const bf16* __restrict__ res, bf16* __restrict__ z, long long M, int C,
                                                              int relu, const float* __restrict__ sums, float* __restrict__ running_mean,
                                                              float* __restrict__ running_var, long long* __restrict__ nbt, float eps,
                                                              float momentum, unsigned char* __restrict__ zmask, int interleaved) {
    pdl_sync();
    extern __shared__ __align__(128) uint8_t st_smem[];
    __shared__ uint64_t st_full[ST_DEPTH];
    constexpr int NT = HAS_RES ? 2 : 1;
    const int cvec = C >> 3;
    const int cv = threadIdx.x & (cvec - 1);
    StreamRing<NT> ring;
    ring.smem = st_smem;
    ring.full = st_full;
    ring.src[0] = reinterpret_cast<const uint8_t*>(y);
    if (HAS_RES) ring.src[NT - 1] = reinterpret_cast<const uint8_t*>(res);
    ring.plan(M * cvec, interleaved);
    ring.start();

    float sc[8], sh[8];
    if (sums) {
        const float inv_m = 1.f / (float)M;
        if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cv * 8 + i;
            const float mean = sums[c] * inv_m;
            const float var = fmaxf(sums[C + c] * inv_m - mean * mean, 0.f);
            const float invstd = rsqrtf(var + eps);
            if (blockIdx.x == 0 && threadIdx.x < cvec) {
                mean_invstd[c] = mean;
                mean_invstd[C + c] = invstd;
                if (running_mean) {
                    const float unbiased = (M > 1) ? var * ((float)M / (float)(M - 1)) : var;
                    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
                }
            }
            sc[i] = gamma[c] * invstd;
            sh[i] = beta[c] - mean * sc[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cv * 8 + i;
            sc[i] = gamma[c] * mean_invstd[C + c];
            sh[i] = beta[c] - mean_invstd[c] * sc[i];
        }
    }
    for (long long i = 0; i < ring.c_count; ++i) {
        const long long v = ring.chunk(i) * 256 + threadIdx.x;
        ring.wait(i);
        float f[8];
        if (v < ring.nvec) {
            unpack8(lds8(ring.slot_ptr(i, 0)), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = f[k] * sc[k] + sh[k];
            if (HAS_RES) {
                float rr[8];
                unpack8(lds8(ring.slot_ptr(i, NT - 1)), rr);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] += rr[k];
            }
        }
        ring.release(i);
        if (v < ring.nvec) {
            if (relu) {
                if (zmask) {
                    unsigned int bits = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) bits |= (f[k] > 0.f ? 1u : 0u) << k;
                    zmask[v] = (unsigned char)bits;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
            }
            st8(z + v * 8, pack8(f));
        }
    }
}

// ------------------------------------------------------------------------------------------ LayerNorm
// One warp per row, D <= 1024, D % 8 == 0: row kept in registers (<= 4 vectors of 8 per lane).
constexpr int LN_MAXV = 4;

template <int NV>
__global__ void __launch_bounds__(256, 3) layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, bf16* __restrict__ y,
                                                           float* __restrict__ stats, long long rows, int D, float eps) {
    pdl_sync();
    // one warp per row, persistent: the next row's 128-bit loads are in flight while the current row is reduced and
    // written; the row stays packed (4 registers per vector) and is unpacked in each pass; gamma / beta come through
    // the read-only path (L1 hits after the first row)
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int nvec = D >> 3;
    const long long stride = (long long)gridDim.x * warps_per_block;
    long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    bf16x8 raw[NV], nxt[NV];
    if (row < rows) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + k * 32;
            if (vi < nvec) raw[k] = ld8(x + row * D + vi * 8);
        }
    }
    for (; row < rows; row += stride) {
        const long long nrow = row + stride;
        if (nrow < rows) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int vi = lane + k * 32;
                if (vi < nvec) nxt[k] = ld8(x + nrow * D + vi * 8);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + k * 32;
            if (vi < nvec) {
                float f[8];
                unpack8(raw[k], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) s += f[i];
            }
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + k * 32;
            if (vi < nvec) {
                float f[8];
                unpack8(raw[k], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float d = f[i] - mean; q += d * d; }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        if (lane == 0 && stats) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + k * 32;
            if (vi < nvec) {
                float o[8], f[8];
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                unpack8(raw[k], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = (f[i] - mean) * rstd * gg[i] + bb[i];
                st8(y + row * D + vi * 8, pack8(o));
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) raw[k] = nxt[k];
    }
}

// Backward: one warp per row, NV vectors of 8 per lane (D = 256*NV max).  dx = rstd*(dy*g - c1 - xhat*c2) (+ dres: the
// skip-connection gradient of the pre-norm residual block, fused here instead of a separate add pass).
// dgamma/dbeta partials stay in registers across the warp's rows, are combined across the block's warps through a
// conflict-free shared layout (no shared atomics) and leave the block as one global atomic per column.
template <int NV>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                               const float* __restrict__ gamma, const float* __restrict__ stats,
                                                               const bf16* __restrict__ dres, bf16* __restrict__ dx,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               long long rows, int D) {
    pdl_sync();
    extern __shared__ float sh[];   // [2][8][D/8]: element (col = vi*8+i) of plane p lives at p*D + i*(D/8) + vi; then gamma[D]
    float* sgam = sh + 2 * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) sgam[i] = gamma[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int nvec = D >> 3;
    float ag[NV][8], ab[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) ag[k][i] = ab[k][i] = 0.f;
    for (long long row = (long long)blockIdx.x * warps_per_block + warp; row < rows;
         row += (long long)gridDim.x * warps_per_block) {
        bf16x8 xr[NV], dr[NV], rr[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + k * 32;
            if (vi < nvec) {
                xr[k] = ld8(x + row * D + vi * 8);
                dr[k] = ld8(dy + row * D + vi * 8);
                if (dres) rr[k] = ld8(dres + row * D + vi * 8);
            }
        }
        const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + k * 32;
            if (vi < nvec) {
                float xv[8], dv[8];
                unpack8(xr[k], xv);
                unpack8(dr[k], dv);
                const float4 g0 = *reinterpret_cast<const float4*>(sgam + vi * 8);
                const float4 g1 = *reinterpret_cast<const float4*>(sgam + vi * 8 + 4);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xh = (xv[i] - mean) * rstd;
                    const float dg = dv[i] * gg[i];
                    c1 += dg;
                    c2 += dg * xh;
                    ag[k][i] += dv[i] * xh;
                    ab[k][i] += dv[i];
                }
            }
        }
        c1 = warp_sum(c1) / (float)D;
        c2 = warp_sum(c2) / (float)D;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + k * 32;
            if (vi < nvec) {
                float xv[8], dv[8], o[8];
                unpack8(xr[k], xv);
                unpack8(dr[k], dv);
                const float4 g0 = *reinterpret_cast<const float4*>(sgam + vi * 8);
                const float4 g1 = *reinterpret_cast<const float4*>(sgam + vi * 8 + 4);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = rstd * (dv[i] * gg[i] - c1 - (xv[i] - mean) * rstd * c2);
                if (dres) {
                    float rv[8];
                    unpack8(rr[k], rv);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] += rv[i];
                }
                st8(dx + row * D + vi * 8, pack8(o));
            }
        }
    }
    for (int w = 0; w < warps_per_block; ++w) {
        if (warp == w) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int vi = lane + k * 32;
                if (vi < nvec) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float* pg = sh + i * nvec + vi;
                        float* pb = sh + D + i * nvec + vi;
                        if (w == 0) { *pg = ag[k][i]; *pb = ab[k][i]; }
                        else { *pg += ag[k][i]; *pb += ab[k][i]; }
                    }
                }
            }
        }
        __syncthreads();
    }
    // thread t owns column t: consecutive lanes hit consecutive addresses, so a warp's 32 atomics cover 4 sectors
    for (int t = threadIdx.x; t < D; t += blockDim.x) {
        const int src = (t & 7) * nvec + (t >> 3);
        atomicAdd(dgamma + t, sh[src]);
        atomicAdd(dbeta + t, sh[D + src]);
    }
}

// ------------------------------------------------------------------------------------------ row softmax
// One warp per row of n <= 256 scores (row stride ld, pad columns [n, ld) written as zero).
__global__ void __launch_bounds__(256) softmax_rows_kernel(const bf16* __restrict__ s, bf16* __restrict__ p, long long rows,
                                                           int n, long long ld) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        float v[8];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + k * 32;
            v[k] = (c < n) ? __bfloat162float(s[row * ld + c]) : -INFINITY;
            mx = fmaxf(mx, v[k]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + k * 32;
            v[k] = (c < n) ? __expf(v[k] - mx) : 0.f;
            sum += v[k];
        }
        const float inv = 1.f / warp_sum(sum);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + k * 32;
            if (c < ld) p[row * ld + c] = __float2bfloat16(v[k] * inv);
        }
    }
}
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const bf16* __restrict__ p, const bf16* __restrict__ dp,
                                                               bf16* __restrict__ ds, long long rows, int n, long long ld,
                                                               float scale) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        float pv[8], dv[8];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + k * 32;
            pv[k] = (c < n) ? __bfloat162float(p[row * ld + c]) : 0.f;
            dv[k] = (c < n) ? __bfloat162float(dp[row * ld + c]) : 0.f;
            dot += pv[k] * dv[k];
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + k * 32;
            if (c < ld) ds[row * ld + c] = __float2bfloat16(pv[k] * (dv[k] - dot) * scale);
        }
    }
}

bool bn_shape_ok(int C) { return C >= 8 && C % 8 == 0 && (C / 8) <= 256 && 256 % (C / 8) == 0; }
// DFU_BN_STREAM=0: register-staged kernels only
const bool g_bn_stream = []() { const char* e = getenv("DFU_BN_STREAM"); return !(e && e[0] == '0'); }();
// blocks of a streaming kernel: `per_sm` resident blocks, each with a contiguous range of 4 KB chunks
inline int stream_grid(long long nvec, int per_sm) { return grid_cap((nvec + 255) / 256, per_sm); }
const int g_stream_interleaved = []() { const char* e = getenv("DFU_STREAM_INTERLEAVED"); return e ? atoi(e) : 1; }();
const int g_stream_per_sm = []() { const char* e = getenv("DFU_STREAM_PER_SM"); return e ? atoi(e) : 0; }();
// resident blocks per SM (measured, profiles/r02j_bn_stream_sweep.txt): 2 for the >= 64 MB tensors of layer1 / layer2, 4 below
inline int stream_per_sm(long long nvec) { return g_stream_per_sm > 0 ? g_stream_per_sm : (nvec >= (4ll << 20) ? 2 : 4); }
template <typename K>
inline void stream_smem_attr(K kernel, size_t bytes) {
    if (bytes > 40 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);   // + static barriers
}


// Eval-mode BatchNorm folded into the preceding convolution's weights: one block per output channel.
__global__ void __launch_bounds__(128) fold_bn_weights_kernel(const float* __restrict__ w, int K, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ mean,
                                                              const float* __restrict__ var, float eps, bf16* __restrict__ w_out,
                                                              long long ldo, float* __restrict__ bias_out) {
    pdl_sync();
    const int co = blockIdx.x;
    const float scale = gamma[co] / sqrtf(var[co] + eps);
    const float* wr = w + (long long)co * K;
    bf16* orow = w_out + (long long)co * ldo;
    for (int k = threadIdx.x; k < (int)ldo; k += blockDim.x) orow[k] = __float2bfloat16(k < K ? wr[k] * scale : 0.f);
    if (threadIdx.x == 0) bias_out[co] = beta[co] - mean[co] * scale;
}

}  // namespace

extern "C" int dfu_bn_stats(const void* y, float* sums, int64_t M, int32_t C, void* stream) {
    DFU_CHECK_ARG(y && sums && M > 0 && bn_shape_ok(C), "dfu_bn_stats: unsupported shape M=%lld C=%d", (long long)M, C);
    const int rpi = 256 / (C / 8);
    // every block ends with 2C global atomics: one resident wave (launch bounds: 2 blocks per SM) keeps them at
    // 296 x 2C -- with 8 blocks per SM the atomics of the wide, short layers (C >= 512) cost more than their loads
    dfu_launch(bn_reduce_kernel<0>, grid_cap((M + rpi - 1) / rpi, 2), 256, 0, (cudaStream_t)stream, (const bf16*)y, nullptr, nullptr, nullptr, nullptr, nullptr, sums, M, C, 0, nullptr);
    DFU_CHECK_LAUNCH("bn_stats");
    return DFU_OK;
}
extern "C" int dfu_bn_finalize(const float* sums, float* mean_invstd, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, int64_t M, int32_t C, float eps, float momentum, void* stream) {
    DFU_CHECK_ARG(sums && mean_invstd && M > 0 && C > 0, "dfu_bn_finalize: bad args");
    dfu_launch(bn_finalize_kernel, dfu_cdiv(C, 128), 128, 0, (cudaStream_t)stream, sums, mean_invstd, running_mean, running_var,
                                                                          (long long*)num_batches_tracked, M, C, eps, momentum);
    DFU_CHECK_LAUNCH("bn_finalize");
    return DFU_OK;
}
extern "C" int dfu_bn_eval_stats(const float* running_mean, const float* running_var, float* mean_invstd, int32_t C,
                                 float eps, void* stream) {
    DFU_CHECK_ARG(running_mean && running_var && mean_invstd && C > 0, "dfu_bn_eval_stats: bad args");
    dfu_launch(bn_eval_stats_kernel, dfu_cdiv(C, 128), 128, 0, (cudaStream_t)stream, running_mean, running_var, mean_invstd, C, eps);
    DFU_CHECK_LAUNCH("bn_eval_stats");
    return DFU_OK;
}
extern "C" int dfu_bn_apply(const void* y, const float* mean_invstd, const float* gamma, const float* beta, const void* res,
                            void* z, uint8_t* zmask, int64_t M, int32_t C, int32_t relu, void* stream) {
    DFU_CHECK_ARG(y && mean_invstd && gamma && beta && z && M > 0 && bn_shape_ok(C), "dfu_bn_apply: unsupported shape M=%lld C=%d",
                  (long long)M, C);
    const int rpi = 256 / (C / 8);
    if (g_bn_stream) {
        const long long nvec = M * (C / 8);
        const int grid = stream_grid(nvec, stream_per_sm(nvec));
        if (res) {
            const size_t smem = (size_t)ST_DEPTH * 2 * ST_CHUNK;
            stream_smem_attr(bn_apply_stream_kernel<true>, smem);
            dfu_launch(bn_apply_stream_kernel<true>, grid, 256, smem, (cudaStream_t)stream, (const bf16*)y, const_cast<float*>(mean_invstd), gamma, beta, (const bf16*)res, (bf16*)z, M, C, relu,
                       (const float*)nullptr, (float*)nullptr, (float*)nullptr, (long long*)nullptr, 0.f, 0.f, zmask, g_stream_interleaved);
        } else {
            const size_t smem = (size_t)ST_DEPTH * ST_CHUNK;
            dfu_launch(bn_apply_stream_kernel<false>, grid, 256, smem, (cudaStream_t)stream, (const bf16*)y, const_cast<float*>(mean_invstd), gamma, beta, (const bf16*)res, (bf16*)z, M, C, relu,
                       (const float*)nullptr, (float*)nullptr, (float*)nullptr, (long long*)nullptr, 0.f, 0.f, zmask, g_stream_interleaved);
        }
    } else
    dfu_launch(bn_apply_kernel, grid_cap((M + rpi - 1) / rpi, 8), 256, 0, (cudaStream_t)stream, (const bf16*)y, const_cast<float*>(mean_invstd), gamma, beta, (const bf16*)res, (bf16*)z, M, C, relu, nullptr, nullptr,
        nullptr, nullptr, 0.f, 0.f, zmask);
    DFU_CHECK_LAUNCH("bn_apply");
    return DFU_OK;
}
extern "C" int dfu_bn_apply_train(const void* y, const float* sums, const float* gamma, const float* beta, const void* res, void* z,
                                  uint8_t* zmask, float* mean_invstd, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                  int64_t M, int32_t C, int32_t relu, float eps, float momentum, void* stream) {
    DFU_CHECK_ARG(y && sums && gamma && beta && z && mean_invstd && M > 0 && bn_shape_ok(C),
                  "dfu_bn_apply_train: unsupported shape M=%lld C=%d", (long long)M, C);
    DFU_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "dfu_bn_apply_train: running_mean/var go together");
    const int rpi = 256 / (C / 8);
    if (g_bn_stream) {
        const long long nvec = M * (C / 8);
        const int grid = stream_grid(nvec, stream_per_sm(nvec));
        if (res) {
            const size_t smem = (size_t)ST_DEPTH * 2 * ST_CHUNK;
            stream_smem_attr(bn_apply_stream_kernel<true>, smem);
            dfu_launch(bn_apply_stream_kernel<true>, grid, 256, smem, (cudaStream_t)stream, (const bf16*)y, mean_invstd, gamma, beta, (const bf16*)res, (bf16*)z, M, C, relu,
                       sums, running_mean, running_var, (long long*)num_batches_tracked, eps, momentum, zmask, g_stream_interleaved);
        } else {
            const size_t smem = (size_t)ST_DEPTH * ST_CHUNK;
            dfu_launch(bn_apply_stream_kernel<false>, grid, 256, smem, (cudaStream_t)stream, (const bf16*)y, mean_invstd, gamma, beta, (const bf16*)res, (bf16*)z, M, C, relu,
                       sums, running_mean, running_var, (long long*)num_batches_tracked, eps, momentum, zmask, g_stream_interleaved);
        }
    } else
    dfu_launch(bn_apply_kernel, grid_cap((M + rpi - 1) / rpi, 8), 256, 0, (cudaStream_t)stream, (const bf16*)y, mean_invstd, gamma, beta, (const bf16*)res, (bf16*)z, M, C, relu, sums, running_mean, running_var,
        (long long*)num_batches_tracked, eps, momentum, zmask);
    DFU_CHECK_LAUNCH("bn_apply_train");
    return DFU_OK;
}
extern "C" int dfu_bn_bwd_reduce(const void* dz, const void* z, const uint8_t* zmask, const void* y, const float* mean_invstd,
                                 const float* gamma, const float* beta, float* sums, int64_t M, int32_t C, int32_t relu,
                                 void* stream) {
    DFU_CHECK_ARG(dz && y && mean_invstd && sums && M > 0 && bn_shape_ok(C) && (!relu || z || zmask || (gamma && beta)),
                  "dfu_bn_bwd_reduce: bad args");
    const int rpi = 256 / (C / 8);
    dfu_launch(bn_reduce_kernel<1>, grid_cap((M + rpi - 1) / rpi, 2), 256, 0, (cudaStream_t)stream, (const bf16*)dz, (const bf16*)z, (const bf16*)y, mean_invstd, gamma, beta, sums, M, C, relu, zmask);
    DFU_CHECK_LAUNCH("bn_bwd_reduce");
    return DFU_OK;
}
extern "C" int dfu_bn_bwd_apply(const void* dz, const void* z, const uint8_t* zmask, const void* y, const float* mean_invstd, const float* gamma,
                                const float* beta, const float* sums, void* dy, void* dres, float* dgamma, float* dbeta,
                                int64_t M, int32_t C, int32_t relu, int32_t train, void* stream) {
    DFU_CHECK_ARG(dz && y && mean_invstd && gamma && sums && dy && M > 0 && bn_shape_ok(C) && (!relu || z || zmask || beta),
                  "dfu_bn_bwd_apply: bad args");
    const int rpi = 256 / (C / 8);
    dfu_launch(bn_bwd_apply_kernel, grid_cap((M + rpi - 1) / rpi, 8), 256, 0, (cudaStream_t)stream, (const bf16*)dz, (const bf16*)z, (const bf16*)y, mean_invstd, gamma, beta, sums, (bf16*)dy, (bf16*)dres, dgamma, dbeta,
        M, C, relu, train, zmask);
    DFU_CHECK_LAUNCH("bn_bwd_apply");
    return DFU_OK;
}

extern "C" int dfu_layernorm(const void* x, const float* gamma, const float* beta, void* y, float* stats, int64_t rows,
                             int32_t D, float eps, void* stream) {
    DFU_CHECK_ARG(x && gamma && beta && y && rows > 0 && D % 8 == 0 && D <= LN_MAXV * 256, "dfu_layernorm: unsupported D=%d", D);
    const int nv = (D / 8 + 31) / 32;
    const int grid = grid_cap((rows + 7) / 8, 3);
#define DFU_LN_FWD(NV) dfu_launch(layernorm_kernel<NV>, grid, 256, 0, (cudaStream_t)stream, (const bf16*)x, gamma, beta, (bf16*)y, stats, rows, D, eps)
    if (nv == 1) DFU_LN_FWD(1);
    else if (nv == 2) DFU_LN_FWD(2);
    else if (nv == 3) DFU_LN_FWD(3);
    else DFU_LN_FWD(4);
#undef DFU_LN_FWD
    DFU_CHECK_LAUNCH("layernorm");
    return DFU_OK;
}
extern "C" int dfu_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* stats, const void* dres,
                                 void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t D, void* stream) {
    DFU_CHECK_ARG(dy && x && gamma && stats && dx && dgamma && dbeta && rows > 0 && D % 8 == 0 && D <= LN_MAXV * 256,
                  "dfu_layernorm_bwd: unsupported D=%d", D);
    const int grid = grid_cap((rows + 7) / 8, 2);
    const size_t smem = 3 * (size_t)D * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    const int nv = (D / 8 + 31) / 32;
#define DFU_LN_BWD(NV)                                                                                              \
    dfu_launch(layernorm_bwd_kernel<NV>, grid, 256, smem, st, (const bf16*)dy, (const bf16*)x, gamma, stats, (const bf16*)dres, \
                                                      (bf16*)dx, dgamma, dbeta, rows, D)
    if (nv == 1) DFU_LN_BWD(1);
    else if (nv == 2) DFU_LN_BWD(2);
    else if (nv == 3) DFU_LN_BWD(3);
    else DFU_LN_BWD(4);
#undef DFU_LN_BWD
    DFU_CHECK_LAUNCH("layernorm_bwd");
    return DFU_OK;
}

extern "C" int dfu_softmax_rows(const void* s, void* p, int64_t rows, int32_t n, int64_t ld, void* stream) {
    DFU_CHECK_ARG(s && p && rows > 0 && n > 0 && n <= 256 && ld >= n && ld <= 256, "dfu_softmax_rows: unsupported n=%d ld=%lld", n,
                  (long long)ld);
    dfu_launch(softmax_rows_kernel, grid_cap((rows + 7) / 8, 8), 256, 0, (cudaStream_t)stream, (const bf16*)s, (bf16*)p, rows, n, ld);
    DFU_CHECK_LAUNCH("softmax_rows");
    return DFU_OK;
}
extern "C" int dfu_softmax_rows_bwd(const void* p, const void* dp, void* ds, int64_t rows, int32_t n, int64_t ld, float scale,
                                    void* stream) {
    DFU_CHECK_ARG(p && dp && ds && rows > 0 && n > 0 && n <= 256 && ld >= n && ld <= 256, "dfu_softmax_rows_bwd: unsupported n=%d", n);
    dfu_launch(softmax_rows_bwd_kernel, grid_cap((rows + 7) / 8, 8), 256, 0, (cudaStream_t)stream, (const bf16*)p, (const bf16*)dp,
                                                                                          (bf16*)ds, rows, n, ld, scale);
    DFU_CHECK_LAUNCH("softmax_rows_bwd");
    return DFU_OK;
}

extern "C" int dfu_fold_bn_weights(const float* w, int32_t Cout, int32_t K, const float* gamma, const float* beta, const float* mean,
                                   const float* var, float eps, void* w_out, int64_t ldo, float* bias_out, void* stream) {
    DFU_CHECK_ARG(w && gamma && beta && mean && var && w_out && bias_out && Cout > 0 && K > 0 && ldo >= K, "dfu_fold_bn_weights: bad args");
    dfu_launch(fold_bn_weights_kernel, Cout, 128, 0, (cudaStream_t)stream, w, K, gamma, beta, mean, var, eps, (bf16*)w_out, ldo, bias_out);
    DFU_CHECK_LAUNCH("dfu_fold_bn_weights");
    return DFU_OK;
}
