// Elementwise / small reduction kernels: ReLU, add, GELU backward, bias-gradient column sums, strided copies,
// dtype casts, dropout, ViT token assembly, pooling.  128-bit accesses, grid-stride loops.
#include "common.cuh"
#include "../../include/dfu_b200.h"

namespace {

inline int grid_for(long long items, int threads, int per_sm = 8) {
    long long blocks = (items + threads - 1) / threads;
    long long cap = (long long)dfu_num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

#define GRID_STRIDE(i, n) \
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

constexpr int PW_UNROLL = 4;
// OP: 0 relu, 1 relu_bwd(a=dy, b=y), 2 add, 3 gelu_bwd(a=dy, b=h), 4 gelu
template <int OP>
__global__ void __launch_bounds__(256) binary_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                     bf16* __restrict__ out, long long n) {
    pdl_sync();
    const long long nv = n >> 3;
    // PW_UNROLL vectors per thread per trip, all loads issued before the first use
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < nv; i0 += stride * PW_UNROLL) {
        bf16x8 va[PW_UNROLL], vb[PW_UNROLL];
#pragma unroll
        for (int u = 0; u < PW_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < nv) {
                va[u] = ld8(a + i * 8);
                if (OP != 0 && OP != 4) vb[u] = ld8(b + i * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < PW_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i >= nv) break;
            float x[8], y[8], o[8];
            unpack8(va[u], x);
            if (OP != 0 && OP != 4) unpack8(vb[u], y);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (OP == 0) o[j] = fmaxf(x[j], 0.f);
                else if (OP == 1) o[j] = (y[j] > 0.f) ? x[j] : 0.f;
                else if (OP == 2) o[j] = x[j] + y[j];
                else if (OP == 4) o[j] = gelu_fast(x[j]);
                else o[j] = x[j] * gelu_grad_fast(y[j]);
            }
            st8(out + i * 8, pack8(o));
        }
    }
    // tail
    if (blockIdx.x == 0) {
        for (long long i = (nv << 3) + threadIdx.x; i < n; i += blockDim.x) {
            const float x = __bfloat162float(a[i]);
            const float y = (OP != 0 && OP != 4) ? __bfloat162float(b[i]) : 0.f;
            float o;
            if (OP == 0) o = fmaxf(x, 0.f);
            else if (OP == 1) o = (y > 0.f) ? x : 0.f;
            else if (OP == 2) o = x + y;
            else if (OP == 4) o = gelu_erf(x);
            else o = x * gelu_erf_grad(y);
            out[i] = __float2bfloat16(o);
        }
    }
}

// Zero fill: 16-byte stores over the aligned body, byte stores for the (at most 15-byte) head and tail.
__global__ void __launch_bounds__(256) zero_kernel(unsigned char* __restrict__ p, long long bytes) {
    pdl_sync();
    const long long head = ((16 - ((uintptr_t)p & 15)) & 15) < (unsigned long long)bytes ? ((16 - ((uintptr_t)p & 15)) & 15) : bytes;
    const long long nv = (bytes - head) >> 4;
    uint4* body = reinterpret_cast<uint4*>(p + head);
    GRID_STRIDE(i, nv) body[i] = make_uint4(0u, 0u, 0u, 0u);
    if (blockIdx.x == 0) {
        for (long long i = threadIdx.x; i < head; i += blockDim.x) p[i] = 0;
        for (long long i = head + (nv << 4) + threadIdx.x; i < bytes; i += blockDim.x) p[i] = 0;
    }
}

// dx = dy * gelu'(h) and, in the same pass, colsum[n] += sum_m dx[m, n] (the bias gradient of the Linear that produced h).
// Block = 128 column vectors (1024 columns) x 2 row lanes; GB_UNROLL rows per thread in flight.
constexpr int GB_UNROLL = 4;
__global__ void __launch_bounds__(256) gelu_bwd_colsum_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ h,
                                                              bf16* __restrict__ dx, float* __restrict__ colsum, long long M,
                                                              int N) {
    pdl_sync();
    __shared__ float red[128][9];
    const int cvl = threadIdx.x & 127, rl = threadIdx.x >> 7;
    const int cv = blockIdx.y * 128 + cvl;                 // column vector (8 columns)
    const bool col_ok = cv * 8 < N;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const long long rstride = (long long)gridDim.x * 2;
    if (col_ok) {
        for (long long r0 = (long long)blockIdx.x * 2 + rl; r0 < M; r0 += rstride * GB_UNROLL) {
            bf16x8 va[GB_UNROLL], vb[GB_UNROLL];
#pragma unroll
            for (int u = 0; u < GB_UNROLL; ++u) {
                const long long r = r0 + u * rstride;
                if (r < M) {
                    va[u] = ld8(dy + r * N + cv * 8);
                    vb[u] = ld8(h + r * N + cv * 8);
                }
            }
#pragma unroll
            for (int u = 0; u < GB_UNROLL; ++u) {
                const long long r = r0 + u * rstride;
                if (r >= M) break;
                float x[8], y[8], o[8];
                unpack8(va[u], x);
                unpack8(vb[u], y);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = x[j] * gelu_grad_fast(y[j]);
                const bf16x8 pk = pack8(o);
                st8(dx + r * N + cv * 8, pk);
                float rb[8];
                unpack8(pk, rb);                               // sum what was stored (the bf16-rounded values)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += rb[j];
            }
        }
    }
    if (rl == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[cvl][i] = acc[i];
    }
    __syncthreads();
    if (rl == 0 && col_ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += red[cvl][i];
        // one lane per 8 columns: transposed through shared memory so that consecutive lanes hit consecutive addresses
#pragma unroll
        for (int i = 0; i < 8; ++i) red[cvl][i] = acc[i];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 1024; t += 256) {
        const int col = blockIdx.y * 1024 + t;
        if (col < N) atomicAdd(colsum + col, red[t >> 3][t & 7]);
    }
}

// Column sums of x[M, N] (row stride ld) added into out[N].  Block = 32 column-vectors x 8 row lanes.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long M, int N,
                                                     long long ld, int rows_per_block) {
    pdl_sync();
    __shared__ float red[8][32][9];
    const int cvl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int cv = blockIdx.x * 32 + cvl;              // column vector (8 columns)
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, M);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (cv * 8 < N) {
        long long r = r0 + rl;
        for (; r + 24 < r1; r += 32) {                 // four independent 16-byte loads in flight per thread
            bf16x8 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld8(x + (r + 8 * u) * ld + cv * 8);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += f[i];
            }
        }
        for (; r < r1; r += 8) {
            float f[8];
            unpack8(ld8(x + r * ld + cv * 8), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rl][cvl][i] = acc[i];
    __syncthreads();
    const int c = threadIdx.x;                         // 256 columns per block
    const int col = blockIdx.x * 256 + c;
    if (col < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][c >> 3][c & 7];
        atomicAdd(out + col, t);
    }
}
__global__ void __launch_bounds__(256) colsum_scalar_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long M,
                                                            int N, long long ld) {
    pdl_sync();
    // small / unaligned N: one thread per column, rows split over blockIdx.y
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    float t = 0.f;
    for (long long r = blockIdx.y; r < M; r += gridDim.y) t += __bfloat162float(x[r * ld + col]);
    atomicAdd(out + col, t);
}

__global__ void __launch_bounds__(256) copy2d_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst,
                                                     long long ldd, long long rows, int cols, int vec) {
    pdl_sync();
    if (vec) {
        const int cv = cols >> 3;
        GRID_STRIDE(t, rows * cv) {
            const long long r = t / cv;
            const int v = (int)(t % cv);
            st8(dst + r * ldd + v * 8, ld8(src + r * lds + v * 8));
        }
    } else {
        GRID_STRIDE(t, rows * cols) {
            const long long r = t / cols;
            const int c = (int)(t % cols);
            dst[r * ldd + c] = src[r * lds + c];
        }
    }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                            long long rows, int cols, long long ldo) {
    pdl_sync();
    GRID_STRIDE(t, rows * ldo) {
        const long long r = t / ldo;
        const int c = (int)(t % ldo);
        dst[t] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.f);
    }
}
__global__ void __launch_bounds__(256) cast_f32_bf16_flat_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                                 long long n) {
    pdl_sync();
    const long long nv = n >> 3;
    GRID_STRIDE(i, nv) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src) + i * 2);
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + i * 2 + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        st8(dst + i * 8, pack8(f));
    }
    if (blockIdx.x == 0)
        for (long long i = (nv << 3) + threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16(src[i]);
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
    pdl_sync();
    GRID_STRIDE(i, n) dst[i] = __bfloat162float(src[i]);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                      uint8_t* __restrict__ mask, long long n, float p,
                                                      const unsigned long long* __restrict__ seed_dev,
                                                      unsigned long long stream_id) {
    pdl_sync();
    const uint64_t seed = seed_dev ? *seed_dev : 0ull;
    const float inv_keep = 1.f / (1.f - p);
    GRID_STRIDE(i, n) {
        const uint64_t h = splitmix64(seed * 0x100000001B3ull + stream_id * 0x9E3779B1ull + (uint64_t)i);
        const float u = (float)(h >> 40) * (1.f / 16777216.f);
        const uint8_t keep = u >= p;
        mask[i] = keep;
        y[i] = __float2bfloat16(keep ? __bfloat162float(x[i]) * inv_keep : 0.f);
    }
}
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ mask,
                                                          bf16* __restrict__ dx, long long n, float p) {
    pdl_sync();
    const float inv_keep = 1.f / (1.f - p);
    GRID_STRIDE(i, n) dx[i] = __float2bfloat16(mask[i] ? __bfloat162float(dy[i]) * inv_keep : 0.f);
}
__global__ void tick_u64_kernel(unsigned long long* c) {
    pdl_sync(); *c += 1ull; }

// ViT token assembly (timm VisionTransformer._pos_embed): cls + pos, patches + pos.
__global__ void __launch_bounds__(256) vit_assemble_kernel(const bf16* __restrict__ patch_out, const float* __restrict__ cls,
                                                           const float* __restrict__ pos, bf16* __restrict__ tokens, int B,
                                                           int T, int D) {
    pdl_sync();
    const int dv = D >> 3;
    GRID_STRIDE(t, (long long)B * T * dv) {
        const int v = (int)(t % dv);
        const long long row = t / dv;
        const int tok = (int)(row % T);
        const int b = (int)(row / T);
        float f[8];
        if (tok == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = cls[v * 8 + i];
        } else {
            unpack8(ld8(patch_out + ((long long)b * (T - 1) + tok - 1) * D + v * 8), f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += pos[(long long)tok * D + v * 8 + i];
        st8(tokens + row * D + v * 8, pack8(f));
    }
}
// adjoint: thread per (token, 8-channel vector); loops over the batch for dpos/dcls (no atomics), writes dpatch.
__global__ void __launch_bounds__(256) vit_assemble_bwd_kernel(const bf16* __restrict__ dtok, bf16* __restrict__ dpatch,
                                                               float* __restrict__ dcls, float* __restrict__ dpos, int B,
                                                               int T, int D) {
    pdl_sync();
    const int dv = D >> 3;
    GRID_STRIDE(t, (long long)T * dv) {
        const int v = (int)(t % dv);
        const int tok = (int)(t / dv);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int b = 0; b < B; ++b) {
            const bf16x8 raw = ld8(dtok + ((long long)b * T + tok) * D + v * 8);
            float f[8];
            unpack8(raw, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += f[i];
            if (tok > 0 && dpatch) st8(dpatch + ((long long)b * (T - 1) + tok - 1) * D + v * 8, raw);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dpos[(long long)tok * D + v * 8 + i] += acc[i];
            if (tok == 0) dcls[v * 8 + i] += acc[i];
        }
    }
}

// ------------------------------------------------------------------ pooling
__global__ void __launch_bounds__(256) maxpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                      uint8_t* __restrict__ idx, int B, int H, int W, int C, int Ho, int Wo) {
    pdl_sync();
    // one thread = 8 channels of one output pixel; the nine window loads are issued before any is consumed
    const int cv = C >> 3;
    const unsigned total = (unsigned)B * Ho * Wo * cv;            // host checks < 2^31
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned v = t % cv, pix = t / cv;
        const int wo = (int)(pix % Wo);
        const unsigned row = pix / Wo;
        const int ho = (int)(row % Ho), b = (int)(row / Ho);
        bf16x8 val[9];
        bool ok[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int h = ho * 2 - 1 + r, w = wo * 2 - 1 + s;
                ok[r * 3 + s] = h >= 0 && h < H && w >= 0 && w < W;
                if (ok[r * 3 + s]) val[r * 3 + s] = ld8(x + (((long long)b * H + h) * W + w) * C + v * 8);
            }
        float best[8];
        uint32_t bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; bi[i] = 255; }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (!ok[k]) continue;
            float f[8];
            unpack8(val[k], f);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (f[i] > best[i] || bi[i] == 255) { best[i] = f[i]; bi[i] = (uint32_t)k; }
        }
        st8(y + (long long)pix * C + v * 8, pack8(best));
        if (idx) {
            uint2 packed;
            packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
            packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
            *reinterpret_cast<uint2*>(idx + (long long)pix * C + v * 8) = packed;
        }
    }
}
// gather form: input pixel (h,w) is tap (r,s) of window (ho,wo) iff h = 2ho-1+r, w = 2wo-1+s -- at most 2 x 2 windows
// (taps of the parity of h+1 / w+1); their winner bytes and gradients are loaded before any is consumed.
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint8_t* __restrict__ idx, const bf16* __restrict__ dy,
                                                          bf16* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo) {
    pdl_sync();
    const int cv = C >> 3;
    const unsigned total = (unsigned)B * H * W * cv;              // host checks < 2^31
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned v = t % cv, pix = t / cv;
        const int w = (int)(pix % W);
        const unsigned row = pix / W;
        const int h = (int)(row % H), b = (int)(row / H);
        // rows: h odd -> taps r = 0 (ho = (h+1)/2) and r = 2 (ho = (h-1)/2); h even -> r = 1 (ho = h/2); same for columns
        int hos[2], rs[2], wos[2], ss[2];
        if (h & 1) { hos[0] = (h + 1) >> 1; rs[0] = 0; hos[1] = (h - 1) >> 1; rs[1] = 2; }
        else       { hos[0] = h >> 1; rs[0] = 1; hos[1] = -1; rs[1] = 0; }
        if (w & 1) { wos[0] = (w + 1) >> 1; ss[0] = 0; wos[1] = (w - 1) >> 1; ss[1] = 2; }
        else       { wos[0] = w >> 1; ss[0] = 1; wos[1] = -1; ss[1] = 0; }
        uint2 win[4];
        bf16x8 gv[4];
        bool ok[4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int k = a * 2 + c;
                ok[k] = hos[a] >= 0 && hos[a] < Ho && wos[c] >= 0 && wos[c] < Wo;
                if (ok[k]) {
                    const long long off = ((((long long)b * Ho + hos[a]) * Wo + wos[c]) * C) + v * 8;
                    win[k] = *reinterpret_cast<const uint2*>(idx + off);
                    gv[k] = ld8(dy + off);
                }
            }
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int k = a * 2 + c;
                if (!ok[k]) continue;
                float g[8];
                unpack8(gv[k], g);
                const uint32_t tap = (uint32_t)(rs[a] * 3 + ss[c]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t word = (i < 4) ? win[k].x : win[k].y;
                    if (((word >> ((i & 3) * 8)) & 0xFFu) == tap) acc[i] += g[i];
                }
            }
        st8(dx + (long long)pix * C + v * 8, pack8(acc));
    }
}

__global__ void __launch_bounds__(256) avgpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int HW, int C,
                                                      long long ldo) {
    pdl_sync();
    const int cv = C >> 3;
    GRID_STRIDE(t, (long long)B * cv) {
        const int v = (int)(t % cv);
        const int b = (int)(t / cv);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int p = 0; p < HW; ++p) {
            float f[8];
            unpack8(ld8(x + ((long long)b * HW + p) * C + v * 8), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
        const float inv = 1.f / (float)HW;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= inv;
        st8(out + (long long)b * ldo + v * 8, pack8(acc));
    }
}
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const bf16* __restrict__ dout, long long ldo, bf16* __restrict__ dx,
                                                          int B, int HW, int C) {
    pdl_sync();
    const int cv = C >> 3;
    const float inv = 1.f / (float)HW;
    GRID_STRIDE(t, (long long)B * HW * cv) {
        const int v = (int)(t % cv);
        const long long q = t / cv;
        const int b = (int)(q / HW);
        float f[8];
        unpack8(ld8(dout + (long long)b * ldo + v * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] *= inv;
        st8(dx + q * C + v * 8, pack8(f));
    }
}

bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <int OP>
int launch_binary(const void* a, const void* b, void* out, long long n, void* stream, const char* name) {
    DFU_CHECK_ARG(a && out && n > 0 && (OP == 0 || OP == 4 || b), "%s: bad args", name);
    DFU_CHECK_ARG(al16(a) && al16(out) && (OP == 0 || OP == 4 || al16(b)), "%s: buffers must be 16-byte aligned", name);
    dfu_launch(binary_kernel<OP>, grid_for((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream, (const bf16*)a, (const bf16*)b, (bf16*)out, n);
    DFU_CHECK_LAUNCH(name);
    return DFU_OK;
}

}  // namespace

extern "C" int dfu_zero(void* p, int64_t bytes, void* stream) {
    DFU_CHECK_ARG(p && bytes >= 0, "dfu_zero: bad args");
    if (bytes == 0) return DFU_OK;
    dfu_launch(zero_kernel, grid_for((bytes + 15) / 16, 256), 256, 0, (cudaStream_t)stream, (unsigned char*)p, bytes);
    DFU_CHECK_LAUNCH("dfu_zero");
    return DFU_OK;
}
extern "C" int dfu_relu(const void* x, void* y, int64_t n, void* stream) { return launch_binary<0>(x, nullptr, y, n, stream, "dfu_relu"); }
extern "C" int dfu_gelu(const void* x, void* y, int64_t n, void* stream) { return launch_binary<4>(x, nullptr, y, n, stream, "dfu_gelu"); }
extern "C" int dfu_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream) {
    return launch_binary<1>(dy, y, dx, n, stream, "dfu_relu_bwd");
}
extern "C" int dfu_add(const void* a, const void* b, void* out, int64_t n, void* stream) {
    return launch_binary<2>(a, b, out, n, stream, "dfu_add");
}
extern "C" int dfu_gelu_bwd(const void* dy, const void* h, void* dx, int64_t n, void* stream) {
    return launch_binary<3>(dy, h, dx, n, stream, "dfu_gelu_bwd");
}

extern "C" int dfu_gelu_bwd_colsum(const void* dy, const void* h, void* dx, float* colsum, int64_t M, int32_t N, void* stream) {
    DFU_CHECK_ARG(dy && h && dx && colsum && M > 0 && N > 0 && N % 8 == 0, "dfu_gelu_bwd_colsum: bad args (N=%d must be a multiple of 8)", N);
    DFU_CHECK_ARG(((uintptr_t)dy & 15) == 0 && ((uintptr_t)h & 15) == 0 && ((uintptr_t)dx & 15) == 0, "dfu_gelu_bwd_colsum: unaligned");
    const int ycnt = dfu_cdiv(N, 1024);
    long long xb = ((long long)dfu_num_sms() * 6 + ycnt - 1) / ycnt;
    const long long max_xb = (M + 1) / 2;
    if (xb > max_xb) xb = max_xb;
    if (xb < 1) xb = 1;
    dfu_launch(gelu_bwd_colsum_kernel, dim3((unsigned)xb, (unsigned)ycnt), 256, 0, (cudaStream_t)stream, (const bf16*)dy, (const bf16*)h, (bf16*)dx, colsum, M, N);
    DFU_CHECK_LAUNCH("dfu_gelu_bwd_colsum");
    return DFU_OK;
}
extern "C" int dfu_colsum(const void* x, float* out, int64_t M, int32_t N, int64_t ld, void* stream) {
    DFU_CHECK_ARG(x && out && M > 0 && N > 0 && ld >= N, "dfu_colsum: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (N % 8 == 0 && ld % 8 == 0 && al16(x)) {
        const int col_blocks = dfu_cdiv(N, 256);
        long long row_blocks = dfu_num_sms() * 4 / col_blocks + 1;
        if (row_blocks > (M + 63) / 64) row_blocks = (M + 63) / 64;
        if (row_blocks < 1) row_blocks = 1;
        const int rows_per_block = (int)((M + row_blocks - 1) / row_blocks);
        dim3 grid(col_blocks, dfu_cdiv(M, rows_per_block));
        dfu_launch(colsum_kernel, grid, 256, 0, st, (const bf16*)x, out, M, N, ld, rows_per_block);
    } else {
        dim3 grid(dfu_cdiv(N, 64), (unsigned)(M < 256 ? M : 256));
        dfu_launch(colsum_scalar_kernel, grid, 64, 0, st, (const bf16*)x, out, M, N, ld);
    }
    DFU_CHECK_LAUNCH("dfu_colsum");
    return DFU_OK;
}

extern "C" int dfu_copy2d(const void* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t cols, void* stream) {
    DFU_CHECK_ARG(src && dst && rows > 0 && cols > 0, "dfu_copy2d: bad args");
    const int vec = (cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && al16(src) && al16(dst)) ? 1 : 0;
    const long long items = vec ? rows * (cols / 8) : rows * cols;
    dfu_launch(copy2d_kernel, grid_for(items, 256), 256, 0, (cudaStream_t)stream, (const bf16*)src, lds, (bf16*)dst, ldd, rows, cols, vec);
    DFU_CHECK_LAUNCH("dfu_copy2d");
    return DFU_OK;
}

extern "C" int dfu_cast_f32_bf16(const float* src, void* dst, int64_t rows, int32_t cols, int64_t ldo, void* stream) {
    DFU_CHECK_ARG(src && dst && rows > 0 && cols > 0 && ldo >= cols, "dfu_cast_f32_bf16: bad args");
    if (ldo == cols && al16(src) && al16(dst)) {
        const long long n = rows * cols;
        dfu_launch(cast_f32_bf16_flat_kernel, grid_for((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, n);
    } else {
        dfu_launch(cast_f32_bf16_kernel, grid_for(rows * ldo, 256), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, rows, cols, ldo);
    }
    DFU_CHECK_LAUNCH("dfu_cast_f32_bf16");
    return DFU_OK;
}
extern "C" int dfu_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream) {
    DFU_CHECK_ARG(src && dst && n > 0, "dfu_cast_bf16_f32: bad args");
    dfu_launch(cast_bf16_f32_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, (const bf16*)src, dst, n);
    DFU_CHECK_LAUNCH("dfu_cast_bf16_f32");
    return DFU_OK;
}

extern "C" int dfu_dropout(const void* x, void* y, uint8_t* mask, int64_t n, float p, const uint64_t* seed_dev,
                           uint64_t stream_id, void* stream) {
    DFU_CHECK_ARG(x && y && mask && n > 0 && p >= 0.f && p < 1.f, "dfu_dropout: bad args (p=%f)", p);
    dfu_launch(dropout_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, mask, n, p,
                                                                      (const unsigned long long*)seed_dev, stream_id);
    DFU_CHECK_LAUNCH("dfu_dropout");
    return DFU_OK;
}
extern "C" int dfu_dropout_bwd(const void* dy, const uint8_t* mask, void* dx, int64_t n, float p, void* stream) {
    DFU_CHECK_ARG(dy && mask && dx && n > 0 && p >= 0.f && p < 1.f, "dfu_dropout_bwd: bad args");
    dfu_launch(dropout_bwd_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, (const bf16*)dy, mask, (bf16*)dx, n, p);
    DFU_CHECK_LAUNCH("dfu_dropout_bwd");
    return DFU_OK;
}
extern "C" int dfu_tick_u64(uint64_t* counter, void* stream) {
    DFU_CHECK_ARG(counter, "dfu_tick_u64: null");
    dfu_launch(tick_u64_kernel, 1, 1, 0, (cudaStream_t)stream, (unsigned long long*)counter);
    DFU_CHECK_LAUNCH("dfu_tick_u64");
    return DFU_OK;
}

extern "C" int dfu_vit_assemble(const void* patch_out, const float* cls, const float* pos, void* tokens, int32_t B, int32_t T,
                                int32_t D, void* stream) {
    DFU_CHECK_ARG(patch_out && cls && pos && tokens && B > 0 && T > 1 && D % 8 == 0, "dfu_vit_assemble: bad args");
    dfu_launch(vit_assemble_kernel, grid_for((long long)B * T * (D / 8), 256), 256, 0, (cudaStream_t)stream, (const bf16*)patch_out, cls, pos, (bf16*)tokens, B, T, D);
    DFU_CHECK_LAUNCH("dfu_vit_assemble");
    return DFU_OK;
}
extern "C" int dfu_vit_assemble_bwd(const void* dtokens, void* dpatch, float* dcls, float* dpos, int32_t B, int32_t T,
                                    int32_t D, void* stream) {
    DFU_CHECK_ARG(dtokens && dcls && dpos && B > 0 && T > 1 && D % 8 == 0, "dfu_vit_assemble_bwd: bad args");
    dfu_launch(vit_assemble_bwd_kernel, grid_for((long long)T * (D / 8), 64, 32), 64, 0, (cudaStream_t)stream, (const bf16*)dtokens, (bf16*)dpatch, dcls, dpos, B, T, D);
    DFU_CHECK_LAUNCH("dfu_vit_assemble_bwd");
    return DFU_OK;
}

extern "C" int dfu_maxpool3x3s2(const void* x, void* y, uint8_t* idx, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
    DFU_CHECK_ARG(x && y && B > 0 && C % 8 == 0 && (long long)B * H * W * (C / 8) < (1ll << 31), "dfu_maxpool3x3s2: bad args");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    dfu_launch(maxpool_kernel, grid_for((long long)B * Ho * Wo * (C / 8), 256), 256, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, idx,
                                                                                                     B, H, W, C, Ho, Wo);
    DFU_CHECK_LAUNCH("dfu_maxpool3x3s2");
    return DFU_OK;
}
extern "C" int dfu_maxpool3x3s2_bwd(const uint8_t* idx, const void* dy, void* dx, int32_t B, int32_t H, int32_t W, int32_t C,
                                    void* stream) {
    DFU_CHECK_ARG(idx && dy && dx && B > 0 && C % 8 == 0 && (long long)B * H * W * (C / 8) < (1ll << 31),
                  "dfu_maxpool3x3s2_bwd: bad args");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    dfu_launch(maxpool_bwd_kernel, grid_for((long long)B * H * W * (C / 8), 256), 256, 0, (cudaStream_t)stream, idx, (const bf16*)dy,
                                                                                                       (bf16*)dx, B, H, W, C, Ho, Wo);
    DFU_CHECK_LAUNCH("dfu_maxpool3x3s2_bwd");
    return DFU_OK;
}
extern "C" int dfu_avgpool(const void* x, void* out, int32_t B, int32_t HW, int32_t C, int64_t ldo, void* stream) {
    DFU_CHECK_ARG(x && out && B > 0 && HW > 0 && C % 8 == 0 && ldo % 8 == 0 && al16(out), "dfu_avgpool: bad args");
    dfu_launch(avgpool_kernel, grid_for((long long)B * (C / 8), 128), 128, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)out, B, HW, C, ldo);
    DFU_CHECK_LAUNCH("dfu_avgpool");
    return DFU_OK;
}
extern "C" int dfu_avgpool_bwd(const void* dout, int64_t ldo, void* dx, int32_t B, int32_t HW, int32_t C, void* stream) {
    DFU_CHECK_ARG(dout && dx && B > 0 && HW > 0 && C % 8 == 0 && ldo % 8 == 0 && al16(dout), "dfu_avgpool_bwd: bad args");
    dfu_launch(avgpool_bwd_kernel, grid_for((long long)B * HW * (C / 8), 256), 256, 0, (cudaStream_t)stream, (const bf16*)dout, ldo, (bf16*)dx,
                                                                                                    B, HW, C);
    DFU_CHECK_LAUNCH("dfu_avgpool_bwd");
    return DFU_OK;
}
