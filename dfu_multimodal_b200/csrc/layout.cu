// Bandwidth-bound layout kernels: uint8 preprocessing, NCHW<->NHWC, ViT patch extraction, im2col / col2im.
// All use 128-bit accesses on the contiguous (channel) dimension and grid-stride loops sized from the SM count.
#include "common.cuh"
#include "../../include/dfu_b200.h"

namespace {

inline int grid_for(long long work_items, int threads, int per_sm = 8) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)dfu_num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

struct Norm3 {
    float scale[3];   // 1/(255*std)
    float shift[3];   // -mean/std
};

// ------------------------------------------------------------------ a9: u8 HWC -> normalised bf16 NHWC
// Thread t converts the 16 bytes [16 t, 16 t + 16): one 128-bit load, two 128-bit stores, consecutive threads on consecutive
// addresses in both directions.  16 = 1 (mod 3), so byte j of the thread is channel (t + j) mod 3: the three possible
// (scale, shift) sequences of a thread are rotations of one another and are picked by t mod 3 once per thread.
// (The first version gave each thread 48 bytes -- three loads 48 bytes apart, six stores 96 bytes apart -- and ran at 16 %
// of the HBM copy rate: profiles/r02b_ncu_full_bandwidth_kernels.txt.)
__global__ void __launch_bounds__(256) preprocess_u8_nhwc_kernel(const uint8_t* __restrict__ in, bf16* __restrict__ out,
                                                                long long n_vec, long long n_bytes, Norm3 nm) {
    pdl_sync();
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_vec; t += (long long)gridDim.x * blockDim.x) {
        const long long base = t * 16;
        const int c0 = (int)(t % 3);
        // channel of byte j is (c0 + j) % 3
        const float s0 = nm.scale[c0], s1 = nm.scale[(c0 + 1) % 3], s2 = nm.scale[(c0 + 2) % 3];
        const float h0 = nm.shift[c0], h1 = nm.shift[(c0 + 1) % 3], h2 = nm.shift[(c0 + 2) % 3];
        if (base + 16 <= n_bytes) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(in + base));
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float v = (float)((w[j >> 2] >> ((j & 3) * 8)) & 0xffu);
                const int r = j % 3;
                f[j] = r == 0 ? fmaf(v, s0, h0) : (r == 1 ? fmaf(v, s1, h1) : fmaf(v, s2, h2));
            }
            st8(out + base, pack8(f));
            st8(out + base + 8, pack8(f + 8));
        } else {
            for (long long i = base; i < n_bytes; ++i) {
                const int c = (int)(i % 3);
                out[i] = __float2bfloat16(fmaf((float)in[i], nm.scale[c], nm.shift[c]));
            }
        }
    }
}

// ------------------------------------------------------------------ a9 fused with ViT patch extraction
// out[(b*196 + ph*14 + pw), c*256 + i*16 + j] = norm(in[b, ph*16+i, pw*16+j, c])
// One thread produces 8 consecutive j (one 16-byte store) from 24 input bytes.
__global__ void __launch_bounds__(256) preprocess_u8_patches_kernel(const uint8_t* __restrict__ in,
                                                                   bf16* __restrict__ out, int B, Norm3 nm) {
    pdl_sync();
    const long long total = (long long)B * 196 * 96;   // 768/8 = 96 vectors per patch row
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int vec = (int)(t % 96);
        const long long row = t / 96;
        const int p = (int)(row % 196);
        const int b = (int)(row / 196);
        const int c = vec / 32;
        const int i = (vec % 32) / 2;
        const int j0 = (vec & 1) * 8;
        const int ph = p / 14, pw = p % 14;
        const uint8_t* src = in + (((long long)b * 224 + ph * 16 + i) * 224 + pw * 16 + j0) * 3 + c;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (float)__ldg(src + j * 3) * nm.scale[c] + nm.shift[c];
        st8(out + row * 768 + vec * 8, pack8(f));
    }
}

// ------------------------------------------------------------------ fp32 NCHW <-> bf16 NHWC
// Tiled transpose through shared memory: [C, HW] <-> [HW, C] per image.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, bf16* __restrict__ out, int C,
                                                           int HW) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float* src = in + (long long)b * C * HW;
    bf16* dst = out + (long long)b * C * HW;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, hw = hw0 + tx;
        tile[r][tx] = (c < C && hw < HW) ? src[(long long)c * HW + hw] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int hw = hw0 + r, c = c0 + tx;
        if (c < C && hw < HW) dst[(long long)hw * C + c] = __float2bfloat16(tile[tx][r]);
    }
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const bf16* __restrict__ in, float* __restrict__ out, int C,
                                                           int HW) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bf16* src = in + (long long)b * C * HW;
    float* dst = out + (long long)b * C * HW;
    for (int r = ty; r < 32; r += 8) {
        const int hw = hw0 + r, c = c0 + tx;
        tile[r][tx] = (c < C && hw < HW) ? __bfloat162float(src[(long long)hw * C + c]) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, hw = hw0 + tx;
        if (c < C && hw < HW) dst[(long long)c * HW + hw] = tile[tx][r];
    }
}

// ------------------------------------------------------------------ fp32 NCHW -> ViT patches and adjoint
__global__ void __launch_bounds__(256) patchify_f32_kernel(const float* __restrict__ in, bf16* __restrict__ out, int B) {
    pdl_sync();
    const long long total = (long long)B * 196 * 96;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int vec = (int)(t % 96);
        const long long row = t / 96;
        const int p = (int)(row % 196);
        const int b = (int)(row / 196);
        const int c = vec / 32, i = (vec % 32) / 2, j0 = (vec & 1) * 8;
        const int ph = p / 14, pw = p % 14;
        const float* src = in + (((long long)b * 3 + c) * 224 + ph * 16 + i) * 224 + pw * 16 + j0;
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
        float f[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
        st8(out + row * 768 + vec * 8, pack8(f));
    }
}
__global__ void __launch_bounds__(256) unpatchify_f32_kernel(const bf16* __restrict__ dp, float* __restrict__ dimg,
                                                             int B) {
    pdl_sync();
    const long long total = (long long)B * 196 * 96;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int vec = (int)(t % 96);
        const long long row = t / 96;
        const int p = (int)(row % 196);
        const int b = (int)(row / 196);
        const int c = vec / 32, i = (vec % 32) / 2, j0 = (vec & 1) * 8;
        const int ph = p / 14, pw = p % 14;
        float f[8];
        unpack8(ld8(dp + row * 768 + vec * 8), f);
        float* dst = dimg + (((long long)b * 3 + c) * 224 + ph * 16 + i) * 224 + pw * 16 + j0;
        *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
}

// ------------------------------------------------------------------ im2col / col2im (NHWC bf16)
// Vector path (C % 8 == 0): one thread moves 8 channels of one (output pixel, tap).
__global__ void __launch_bounds__(256) im2col_vec_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int B, int H,
                                                         int W, int C, int R, int S, int stride, int pad, int Ho, int Wo,
                                                         long long ldc) {
    pdl_sync();
    const int cv = C / 8;
    const long long per_row = (long long)R * S * cv;
    const long long total = (long long)B * Ho * Wo * per_row;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t % cv);
        long long q = t / cv;
        const int tap = (int)(q % (R * S));
        const long long m = q / (R * S);
        const int wo = (int)(m % Wo);
        const int ho = (int)((m / Wo) % Ho);
        const int b = (int)(m / ((long long)Wo * Ho));
        const int r = tap / S, s = tap % S;
        const int h = ho * stride - pad + r, w = wo * stride - pad + s;
        bf16x8 val;
        if (h >= 0 && h < H && w >= 0 && w < W) {
            val = ld8(x + (((long long)b * H + h) * W + w) * C + v * 8);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) val.v[i] = __floats2bfloat162_rn(0.f, 0.f);
        }
        st8(col + m * ldc + (long long)tap * C + v * 8, val);
    }
}
// Scalar path (stem, C = 3): one thread per (output pixel, column) including the zero pad columns.
__global__ void __launch_bounds__(256) im2col_scalar_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int B,
                                                            int H, int W, int C, int R, int S, int stride, int pad,
                                                            int Ho, int Wo, long long ldc) {
    pdl_sync();
    const int Kc = R * S * C;
    const long long total = (long long)B * Ho * Wo * ldc;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % ldc);
        const long long m = t / ldc;
        float val = 0.f;
        if (k < Kc) {
            const int c = k % C;
            const int tap = k / C;
            const int r = tap / S, s = tap % S;
            const int wo = (int)(m % Wo);
            const int ho = (int)((m / Wo) % Ho);
            const int b = (int)(m / ((long long)Wo * Ho));
            const int h = ho * stride - pad + r, w = wo * stride - pad + s;
            if (h >= 0 && h < H && w >= 0 && w < W)
                val = __bfloat162float(x[(((long long)b * H + h) * W + w) * C + c]);
        }
        col[t] = __float2bfloat16(val);
    }
}

// Stem path (C = 3, e.g. the 7x7/s2 first convolution): one block per (image, output row).  The R input rows the
// output row touches are staged in shared memory with 128-bit loads, then every thread assembles 8 consecutive
// columns of one col row (k = (r*S + s)*3 + c, i.e. S*3 contiguous input elements per filter row) and writes them
// with one 128-bit store; consecutive threads write consecutive 16-byte chunks, so stores are fully coalesced.
constexpr int C3_ROWS_PER_BLOCK = 4;      // output rows per block (their input rows overlap: staged once)
__global__ void __launch_bounds__(256) im2col_c3_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int H, int W,
                                                        int R, int S, int stride, int pad, int Ho, int Wo,
                                                        int ldc) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    bf16* rows = reinterpret_cast<bf16*>(sm_raw);            // [(C3_ROWS_PER_BLOCK-1)*stride + R][W*3]
    const int hblocks = (Ho + C3_ROWS_PER_BLOCK - 1) / C3_ROWS_PER_BLOCK;
    const int ho0 = (blockIdx.x % hblocks) * C3_ROWS_PER_BLOCK;
    const int b = blockIdx.x / hblocks;
    const int nho = min(C3_ROWS_PER_BLOCK, Ho - ho0);
    const int nrows = (nho - 1) * stride + R;
    const int rowlen = W * 3;
    const int vec_per_row = rowlen >> 3;
    const int h_first = ho0 * stride - pad;
    for (int t = threadIdx.x; t < nrows * vec_per_row; t += blockDim.x) {
        const int r = t / vec_per_row, v = t - r * vec_per_row;
        const int h = h_first + r;
        bf16x8 val;
        if (h >= 0 && h < H) {
            val = ld8(x + ((long long)b * H + h) * rowlen + v * 8);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) val.v[i] = __floats2bfloat162_rn(0.f, 0.f);
        }
        *reinterpret_cast<bf16x8*>(rows + r * rowlen + v * 8) = val;
    }
    __syncthreads();
    const int chunks = ldc >> 3;
    const int Kc = R * S * 3;
    const int seg = S * 3;
    const bf16 zero = __float2bfloat16(0.f);
    bf16* out = col + ((long long)b * Ho + ho0) * Wo * ldc;
    const int per_row = Wo * chunks;
    for (int t = threadIdx.x; t < nho * per_row; t += blockDim.x) {
        const int hl = t / per_row;
        const int tt = t - hl * per_row;
        const int wo = tt / chunks, ch = tt - wo * chunks;
        const int w0 = (wo * stride - pad) * 3;
        const bf16* base = rows + hl * stride * rowlen;
        alignas(16) bf16 e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = ch * 8 + i;
            bf16 v = zero;
            if (k < Kc) {
                const int r = k / seg, j = k - r * seg;
                const int we = w0 + j;
                if (we >= 0 && we < rowlen) v = base[r * rowlen + we];
            }
            e[i] = v;
        }
        *reinterpret_cast<bf16x8*>(out + (long long)t * 8) = *reinterpret_cast<const bf16x8*>(e);
    }
}
// The same with the filter geometry fixed at compile time (the ResNet stem: 7x7, stride 2, pad 3, ldc 152): one thread
// assembles a whole col row, so every source offset (filter row r, column j of the S*3 contiguous input elements) is
// an immediate -- no division, no per-element bounds test for pixels whose window lies inside the image row.
// (Measured on B200, B=128: 449 us generic -> 270 us; assembling whole output rows in shared memory and draining them
// with cp.async.bulk was slower, 590 us.)
template <int R, int S, int STRIDE, int PAD, int LDC>
__global__ void __launch_bounds__(256) im2col_c3_fixed_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int H, int W,
                                                              int Ho, int Wo) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    bf16* rows = reinterpret_cast<bf16*>(sm_raw);            // [(C3_ROWS_PER_BLOCK-1)*STRIDE + R][W*3]
    const int hblocks = (Ho + C3_ROWS_PER_BLOCK - 1) / C3_ROWS_PER_BLOCK;
    const int ho0 = (blockIdx.x % hblocks) * C3_ROWS_PER_BLOCK;
    const int b = blockIdx.x / hblocks;
    const int nho = min(C3_ROWS_PER_BLOCK, Ho - ho0);
    const int nrows = (nho - 1) * STRIDE + R;
    const int rowlen = W * 3;
    const int vec_per_row = rowlen >> 3;
    const int h_first = ho0 * STRIDE - PAD;
    for (int t = threadIdx.x; t < nrows * vec_per_row; t += blockDim.x) {
        const int r = t / vec_per_row, v = t - r * vec_per_row;
        const int h = h_first + r;
        bf16x8 val;
        if (h >= 0 && h < H) {
            val = ld8(x + ((long long)b * H + h) * rowlen + v * 8);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) val.v[i] = __floats2bfloat162_rn(0.f, 0.f);
        }
        *reinterpret_cast<bf16x8*>(rows + r * rowlen + v * 8) = val;
    }
    __syncthreads();
    constexpr int SEG = S * 3, KC = R * S * 3, CHUNKS = LDC / 8;
    const unsigned short* rows16 = reinterpret_cast<const unsigned short*>(rows);
    bf16* out = col + ((long long)b * Ho + ho0) * Wo * LDC;
    for (int t = threadIdx.x; t < nho * Wo; t += blockDim.x) {
        const int hl = t / Wo, wo = t - hl * Wo;
        const int w0 = (wo * STRIDE - PAD) * 3;
        const unsigned short* base = rows16 + hl * STRIDE * rowlen + w0;
        uint4* orow = reinterpret_cast<uint4*>(out + (long long)t * LDC);
        if (w0 >= 0 && w0 + SEG <= rowlen) {                 // window inside the image row: no per-element tests
#pragma unroll
            for (int ch = 0; ch < CHUNKS; ++ch) {
                uint32_t wd[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k0 = ch * 8 + 2 * i, k1 = k0 + 1;
                    uint32_t lo = 0, hi = 0;
                    if (k0 < KC) lo = base[(k0 / SEG) * rowlen + (k0 % SEG)];
                    if (k1 < KC) hi = base[(k1 / SEG) * rowlen + (k1 % SEG)];
                    wd[i] = lo | (hi << 16);
                }
                orow[ch] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            }
        } else {
#pragma unroll
            for (int ch = 0; ch < CHUNKS; ++ch) {
                uint32_t wd[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k0 = ch * 8 + 2 * i, k1 = k0 + 1;
                    uint32_t lo = 0, hi = 0;
                    if (k0 < KC && w0 + (k0 % SEG) >= 0 && w0 + (k0 % SEG) < rowlen) lo = base[(k0 / SEG) * rowlen + (k0 % SEG)];
                    if (k1 < KC && w0 + (k1 % SEG) >= 0 && w0 + (k1 % SEG) < rowlen) hi = base[(k1 / SEG) * rowlen + (k1 % SEG)];
                    wd[i] = lo | (hi << 16);
                }
                orow[ch] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            }
        }
    }
}
// Zero the pad columns [Kc, ldc) of the vector path when ldc > Kc.
__global__ void __launch_bounds__(256) zero_pad_cols_kernel(bf16* __restrict__ col, long long rows, int Kc, long long ldc) {
    pdl_sync();
    const int padw = (int)(ldc - Kc);
    const long long total = rows * padw;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x)
        col[(t / padw) * ldc + Kc + (t % padw)] = __float2bfloat16(0.f);
}

// col2im, gather form: each thread owns 8 channels of one input pixel and sums the taps that touched it.
__global__ void __launch_bounds__(256) col2im_vec_kernel(const bf16* __restrict__ col, bf16* __restrict__ dx, int B, int H,
                                                         int W, int C, int R, int S, int stride, int pad, int Ho, int Wo,
                                                         long long ldc) {
    pdl_sync();
    const int cv = C / 8;
    const long long total = (long long)B * H * W * cv;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t % cv);
        const long long pix = t / cv;
        const int w = (int)(pix % W);
        const int h = (int)((pix / W) % H);
        const int b = (int)(pix / ((long long)W * H));
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int r = 0; r < R; ++r) {
            const int hn = h + pad - r;
            if (hn < 0 || hn % stride != 0) continue;
            const int ho = hn / stride;
            if (ho >= Ho) continue;
            for (int s = 0; s < S; ++s) {
                const int wn = w + pad - s;
                if (wn < 0 || wn % stride != 0) continue;
                const int wo = wn / stride;
                if (wo >= Wo) continue;
                const long long m = ((long long)b * Ho + ho) * Wo + wo;
                float f[8];
                unpack8(ld8(col + m * ldc + (long long)(r * S + s) * C + v * 8), f);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += f[i];
            }
        }
        st8(dx + pix * C + v * 8, pack8(acc));
    }
}
__global__ void __launch_bounds__(256) col2im_scalar_kernel(const bf16* __restrict__ col, bf16* __restrict__ dx, int B,
                                                            int H, int W, int C, int R, int S, int stride, int pad,
                                                            int Ho, int Wo, long long ldc) {
    pdl_sync();
    const long long total = (long long)B * H * W * C;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        const long long pix = t / C;
        const int w = (int)(pix % W);
        const int h = (int)((pix / W) % H);
        const int b = (int)(pix / ((long long)W * H));
        float acc = 0.f;
        for (int r = 0; r < R; ++r) {
            const int hn = h + pad - r;
            if (hn < 0 || hn % stride != 0) continue;
            const int ho = hn / stride;
            if (ho >= Ho) continue;
            for (int s = 0; s < S; ++s) {
                const int wn = w + pad - s;
                if (wn < 0 || wn % stride != 0) continue;
                const int wo = wn / stride;
                if (wo >= Wo) continue;
                const long long m = ((long long)b * Ho + ho) * Wo + wo;
                acc += __bfloat162float(col[m * ldc + (long long)(r * S + s) * C + c]);
            }
        }
        dx[t] = __float2bfloat16(acc);
    }
}

Norm3 make_norm(const float* mean, const float* std) {
    Norm3 n;
    for (int c = 0; c < 3; ++c) {
        n.scale[c] = 1.f / (255.f * std[c]);
        n.shift[c] = -mean[c] / std[c];
    }
    return n;
}

}  // namespace

extern "C" int dfu_preprocess_u8_nhwc(const uint8_t* in, void* out, int64_t n_pixels, const float* mean3_host,
                                      const float* std3_host, void* stream) {
    DFU_CHECK_ARG(in && out && n_pixels > 0 && mean3_host && std3_host, "dfu_preprocess_u8_nhwc: bad args");
    DFU_CHECK_ARG(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "dfu_preprocess_u8_nhwc: unaligned buffers");
    const long long n_bytes = n_pixels * 3;
    const long long n_vec = (n_bytes + 15) / 16;
    dfu_launch(preprocess_u8_nhwc_kernel, grid_for(n_vec, 256, 16), 256, 0, (cudaStream_t)stream, in, (bf16*)out, n_vec, n_bytes, make_norm(mean3_host, std3_host));
    DFU_CHECK_LAUNCH("preprocess_u8_nhwc");
    return DFU_OK;
}

extern "C" int dfu_preprocess_u8_patches(const uint8_t* in, void* out, int32_t B, const float* mean3_host,
                                         const float* std3_host, void* stream) {
    DFU_CHECK_ARG(in && out && B > 0 && mean3_host && std3_host, "dfu_preprocess_u8_patches: bad args");
    dfu_launch(preprocess_u8_patches_kernel, grid_for((long long)B * 196 * 96, 256), 256, 0, (cudaStream_t)stream, in, (bf16*)out, B, make_norm(mean3_host, std3_host));
    DFU_CHECK_LAUNCH("preprocess_u8_patches");
    return DFU_OK;
}

extern "C" int dfu_nchw_f32_to_nhwc_bf16(const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W,
                                         void* stream) {
    DFU_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "dfu_nchw_f32_to_nhwc_bf16: bad args");
    dim3 grid(dfu_cdiv(H * W, 32), dfu_cdiv(C, 32), B);
    dfu_launch(nchw_to_nhwc_kernel, grid, 256, 0, (cudaStream_t)stream, in, (bf16*)out, C, H * W);
    DFU_CHECK_LAUNCH("nchw_to_nhwc");
    return DFU_OK;
}
extern "C" int dfu_nhwc_bf16_to_nchw_f32(const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                                         void* stream) {
    DFU_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "dfu_nhwc_bf16_to_nchw_f32: bad args");
    dim3 grid(dfu_cdiv(H * W, 32), dfu_cdiv(C, 32), B);
    dfu_launch(nhwc_to_nchw_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)in, out, C, H * W);
    DFU_CHECK_LAUNCH("nhwc_to_nchw");
    return DFU_OK;
}

extern "C" int dfu_patchify_f32(const float* in, void* out, int32_t B, void* stream) {
    DFU_CHECK_ARG(in && out && B > 0, "dfu_patchify_f32: bad args");
    DFU_CHECK_ARG(((uintptr_t)in & 15) == 0, "dfu_patchify_f32: input must be 16-byte aligned");
    dfu_launch(patchify_f32_kernel, grid_for((long long)B * 196 * 96, 256), 256, 0, (cudaStream_t)stream, in, (bf16*)out, B);
    DFU_CHECK_LAUNCH("patchify_f32");
    return DFU_OK;
}
extern "C" int dfu_unpatchify_f32(const void* dpatches, float* dimg, int32_t B, void* stream) {
    DFU_CHECK_ARG(dpatches && dimg && B > 0, "dfu_unpatchify_f32: bad args");
    dfu_launch(unpatchify_f32_kernel, grid_for((long long)B * 196 * 96, 256), 256, 0, (cudaStream_t)stream, (const bf16*)dpatches,
                                                                                                  dimg, B);
    DFU_CHECK_LAUNCH("unpatchify_f32");
    return DFU_OK;
}

extern "C" int dfu_im2col_nhwc(const void* x, void* col, int32_t B, int32_t H, int32_t W, int32_t C, int32_t R, int32_t S,
                               int32_t stride, int32_t pad, int32_t Ho, int32_t Wo, int64_t ldc, void* stream) {
    DFU_CHECK_ARG(x && col && B > 0 && ldc >= (int64_t)R * S * C, "dfu_im2col_nhwc: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * Ho * Wo;
    if (C % 8 == 0 && ldc % 8 == 0) {
        const long long total = rows * R * S * (C / 8);
        dfu_launch(im2col_vec_kernel, grid_for(total, 256, 16), 256, 0, st, (const bf16*)x, (bf16*)col, B, H, W, C, R, S, stride,
                                                                    pad, Ho, Wo, ldc);
        DFU_CHECK_LAUNCH("im2col_vec");
        if (ldc > (int64_t)R * S * C) {
            dfu_launch(zero_pad_cols_kernel, grid_for(rows * (ldc - R * S * C), 256), 256, 0, st, (bf16*)col, rows, R * S * C, ldc);
            DFU_CHECK_LAUNCH("zero_pad_cols");
        }
    } else if (C == 3 && (W * 3) % 8 == 0 && ldc % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)col & 15) == 0 &&
               (size_t)((C3_ROWS_PER_BLOCK - 1) * stride + R) * W * 3 * sizeof(bf16) <= 48 * 1024) {
        const int hblocks = (Ho + C3_ROWS_PER_BLOCK - 1) / C3_ROWS_PER_BLOCK;
        if (R == 7 && S == 7 && stride == 2 && pad == 3 && ldc == 152) {
            dfu_launch(im2col_c3_fixed_kernel<7, 7, 2, 3, 152>, B * hblocks, 256, (size_t)((C3_ROWS_PER_BLOCK - 1) * 2 + 7) * W * 3 * sizeof(bf16), st, (const bf16*)x, (bf16*)col, H, W, Ho, Wo);
            DFU_CHECK_LAUNCH("im2col_c3_fixed");
            return DFU_OK;
        }
        dfu_launch(im2col_c3_kernel, B * hblocks, 256, (size_t)((C3_ROWS_PER_BLOCK - 1) * stride + R) * W * 3 * sizeof(bf16), st, (const bf16*)x, (bf16*)col, H, W, R, S, stride, pad, Ho, Wo, (int)ldc);
        DFU_CHECK_LAUNCH("im2col_c3");
    } else {
        dfu_launch(im2col_scalar_kernel, grid_for(rows * ldc, 256, 16), 256, 0, st, (const bf16*)x, (bf16*)col, B, H, W, C, R, S,
                                                                            stride, pad, Ho, Wo, ldc);
        DFU_CHECK_LAUNCH("im2col_scalar");
    }
    return DFU_OK;
}

extern "C" int dfu_col2im_nhwc(const void* col, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t R, int32_t S,
                               int32_t stride, int32_t pad, int32_t Ho, int32_t Wo, int64_t ldc, void* stream) {
    DFU_CHECK_ARG(col && dx && B > 0 && ldc >= (int64_t)R * S * C, "dfu_col2im_nhwc: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 == 0 && ldc % 8 == 0) {
        dfu_launch(col2im_vec_kernel, grid_for((long long)B * H * W * (C / 8), 256, 16), 256, 0, st, (const bf16*)col, (bf16*)dx, B, H, W, C, R, S, stride, pad, Ho, Wo, ldc);
        DFU_CHECK_LAUNCH("col2im_vec");
    } else {
        dfu_launch(col2im_scalar_kernel, grid_for((long long)B * H * W * C, 256, 16), 256, 0, st, (const bf16*)col, (bf16*)dx, B, H, W, C, R, S, stride, pad, Ho, Wo, ldc);
        DFU_CHECK_LAUNCH("col2im_scalar");
    }
    return DFU_OK;
}
