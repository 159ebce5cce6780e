// fp32 VALIDATION MODE (north_star: "logits within 1e-4 (fp32-accum validation mode)").
// Every op of the inference path once more, with fp32 activations, the fp32 MASTER weights and fp32 (statistics: fp64)
// accumulation, as plain SIMT kernels: slow by design, used by dfu_multimodal_b200/validation.py to separate "the bf16
// kernels round differently" from "the algorithm differs from the reference".  Activations are NHWC rows [pixels, C]
// like on the fast path, tokens [B*T, D].  Nothing here is on the benchmarked path.
#include "common.cuh"
#include "../../include/dfu_b200.h"

namespace {

#define V32_LOOP(i, n) for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

inline int v32_grid(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    const long long cap = (long long)dfu_num_sms() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------------------------------------- layout
__global__ void __launch_bounds__(256) v32_preprocess_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, long long npix,
                                                                float m0, float m1, float m2, float s0, float s1, float s2) {
    V32_LOOP(p, npix) {
        // ToTensor: x / 255, then Normalize: (x - mean) / std -- the same operation order as torchvision
        out[p * 3 + 0] = ((float)img[p * 3 + 0] / 255.f - m0) / s0;
        out[p * 3 + 1] = ((float)img[p * 3 + 1] / 255.f - m1) / s1;
        out[p * 3 + 2] = ((float)img[p * 3 + 2] / 255.f - m2) / s2;
    }
}
__global__ void __launch_bounds__(256) v32_nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W) {
    const long long n = (long long)B * C * H * W;
    V32_LOOP(i, n) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        y[i] = x[(((long long)b * C + c) * H + h) * W + w];
    }
}
// im2col over NHWC: col[(b,ho,wo), (r,s,c)]
__global__ void __launch_bounds__(256) v32_im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H, int W, int C,
                                                         int R, int S, int stride, int pad, int Ho, int Wo) {
    const long long K = (long long)R * S * C, n = (long long)B * Ho * Wo * K;
    V32_LOOP(i, n) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int s = (int)(t % S); t /= S;
        const int r = (int)(t % R); t /= R;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const int h = ho * stride - pad + r, w = wo * stride - pad + s;
        col[i] = (h >= 0 && h < H && w >= 0 && w < W) ? x[(((long long)b * H + h) * W + w) * C + c] : 0.f;
    }
}
// timm PatchEmbed unfold of an NHWC image: out[(b, ph, pw), (c, i, j)]
__global__ void __launch_bounds__(256) v32_patchify_kernel(const float* __restrict__ x, float* __restrict__ out, int B) {
    const long long n = (long long)B * 196 * 768;
    V32_LOOP(t, n) {
        const int col = (int)(t % 768);
        const long long row = t / 768;
        const int c = col / 256, i = (col % 256) / 16, j = col % 16;
        const int b = (int)(row / 196), p = (int)(row % 196), ph = p / 14, pw = p % 14;
        out[t] = x[(((long long)b * 224 + ph * 16 + i) * 224 + pw * 16 + j) * 3 + c];
    }
}
// x <- float(bf16(x)): what storing an activation (or a weight shadow) as bf16 does to it
__global__ void __launch_bounds__(256) v32_round_bf16_kernel(float* __restrict__ x, long long n) {
    V32_LOOP(i, n) x[i] = __bfloat162float(__float2bfloat16(x[i]));
}
__global__ void __launch_bounds__(256) v32_copy2d_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
                                                         long long rows, int cols) {
    V32_LOOP(i, rows * cols) dst[(i / cols) * ldd + i % cols] = src[(i / cols) * lds + i % cols];
}
__global__ void __launch_bounds__(256) v32_vit_assemble_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls,
                                                               const float* __restrict__ pos, float* __restrict__ tok, int B, int T, int D) {
    V32_LOOP(i, (long long)B * T * D) {
        const int d = (int)(i % D);
        const int t = (int)((i / D) % T);
        const int b = (int)(i / ((long long)D * T));
        const float v = t == 0 ? cls[d] : patch_out[((long long)b * (T - 1) + t - 1) * D + d];
        tok[i] = v + pos[(long long)t * D + d];
    }
}

// ---------------------------------------------------------------------------------------------- GEMM
// D[z][m][n] = act(alpha * sum_k A[z][m][k] * B[z][n][k] + bias[n]) + res[z][m][n]; every operand addressed by element strides,
// so K-major / MN-major operands, head slices of the packed qkv matrix and batched launches are all the same kernel.
struct V32Gemm {
    int M, N, K, nb0, nb1;
    const float *A, *B, *bias, *res;
    float* D;
    long long a_rs, a_cs, a_s0, a_s1, b_rs, b_cs, b_s0, b_s1, ldd, d_s0, d_s1, ldr, r_s0, r_s1;
    float alpha;
    int act;
};
constexpr int VT = 64, VK = 16;
__global__ void __launch_bounds__(256) v32_gemm_kernel(V32Gemm p) {
    __shared__ float As[VK][VT + 1], Bs[VK][VT + 1];
    const int z0 = blockIdx.z % p.nb0, z1 = blockIdx.z / p.nb0;
    const float* A = p.A + z0 * p.a_s0 + z1 * p.a_s1;
    const float* B = p.B + z0 * p.b_s0 + z1 * p.b_s1;
    const int m0 = blockIdx.y * VT, n0 = blockIdx.x * VT;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += VK) {
        for (int e = threadIdx.x; e < VT * VK; e += 256) {
            const int k = e % VK, r = e / VK;
            As[k][r] = (m0 + r < p.M && k0 + k < p.K) ? A[(long long)(m0 + r) * p.a_rs + (long long)(k0 + k) * p.a_cs] : 0.f;
            Bs[k][r] = (n0 + r < p.N && k0 + k < p.K) ? B[(long long)(n0 + r) * p.b_rs + (long long)(k0 + k) * p.b_cs] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < VK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* D = p.D + z0 * p.d_s0 + z1 * p.d_s1;
    const float* R = p.res ? p.res + z0 * p.r_s0 + z1 * p.r_s1 : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m >= p.M || n >= p.N) continue;
            float v = acc[i][j] * p.alpha;
            if (p.bias) v += p.bias[n];
            if (p.act == 1) v = fmaxf(v, 0.f);
            else if (p.act == 2) v = gelu_erf(v);
            if (R) v += R[(long long)m * p.ldr + n];
            D[(long long)m * p.ldd + n] = v;
        }
}

// ---------------------------------------------------------------------------------------------- normalisation / pooling
// per-channel batch mean and biased variance over the rows of y [M, C], fp64 accumulation (one block per 32 channels)
__global__ void __launch_bounds__(256) v32_bn_batch_stats_kernel(const float* __restrict__ y, float* __restrict__ mean, float* __restrict__ var,
                                                                 long long M, int C) {
    __shared__ double s1[8][32], s2[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
    double a = 0.0, b = 0.0;
    if (c < C)
        for (long long m = g; m < M; m += 8) {
            const double v = (double)y[m * C + c];
            a += v;
            b += v * v;
        }
    s1[g][threadIdx.x & 31] = a;
    s2[g][threadIdx.x & 31] = b;
    __syncthreads();
    if (g == 0 && c < C) {
        for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
        const double mu = a / (double)M;
        mean[c] = (float)mu;
        var[c] = (float)fmax(b / (double)M - mu * mu, 0.0);
    }
}
__global__ void __launch_bounds__(256) v32_bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ var,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ res, float* __restrict__ out, long long M, int C, int relu,
                                                           float eps) {
    V32_LOOP(i, M * C) {
        const int c = (int)(i % C);
        float v = (y[i] - mean[c]) / sqrtf(var[c] + eps) * gamma[c] + beta[c];
        if (res) v += res[i];
        out[i] = relu ? fmaxf(v, 0.f) : v;
    }
}
__global__ void __launch_bounds__(256) v32_maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
    V32_LOOP(i, (long long)B * Ho * Wo * C) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int b = (int)(t / Ho);
        float m = -INFINITY;
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
                const int h = ho * 2 - 1 + r, w = wo * 2 - 1 + s;
                if (h >= 0 && h < H && w >= 0 && w < W) m = fmaxf(m, x[(((long long)b * H + h) * W + w) * C + c]);
            }
        y[i] = m;
    }
}
__global__ void __launch_bounds__(256) v32_avgpool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int HW, int C) {
    V32_LOOP(i, (long long)B * C) {
        const int c = (int)(i % C), b = (int)(i / C);
        double s = 0.0;
        for (int p = 0; p < HW; ++p) s += (double)x[((long long)b * HW + p) * C + c];
        y[i] = (float)(s / (double)HW);
    }
}
// one warp per row
__global__ void __launch_bounds__(256) v32_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ y, long long rows, int D, float eps) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + row * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += xr[d];
    const float mu = warp_sum(s) / (float)D;
    float q = 0.f;
    for (int d = lane; d < D; d += 32) { const float t = xr[d] - mu; q += t * t; }
    const float rstd = 1.f / sqrtf(warp_sum(q) / (float)D + eps);
    for (int d = lane; d < D; d += 32) y[row * D + d] = (xr[d] - mu) * rstd * gamma[d] + beta[d];
}
__global__ void __launch_bounds__(256) v32_softmax_rows_kernel(float* __restrict__ x, long long rows, int T, long long ld) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* xr = x + row * ld;
    float m = -INFINITY;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, xr[t]);
    m = warp_max(m);
    float s = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(xr[t] - m); xr[t] = e; s += e; }
    s = warp_sum(s);
    for (int t = lane; t < T; t += 32) xr[t] = xr[t] / s;
}

}  // namespace

#define V32_ST ((cudaStream_t)stream)

extern "C" int dfu_v32_preprocess_u8(const uint8_t* img_nhwc, float* out_nhwc, int64_t npix, const float* mean3, const float* std3, void* stream) {
    DFU_CHECK_ARG(img_nhwc && out_nhwc && npix > 0 && mean3 && std3, "dfu_v32_preprocess_u8: bad args");
    v32_preprocess_u8_kernel<<<v32_grid(npix), 256, 0, V32_ST>>>(img_nhwc, out_nhwc, npix, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
    DFU_CHECK_LAUNCH("dfu_v32_preprocess_u8");
    return DFU_OK;
}
extern "C" int dfu_v32_nchw_to_nhwc(const float* x, float* y, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
    DFU_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0, "dfu_v32_nchw_to_nhwc: bad args");
    v32_nchw_to_nhwc_kernel<<<v32_grid((long long)B * C * H * W), 256, 0, V32_ST>>>(x, y, B, C, H, W);
    DFU_CHECK_LAUNCH("dfu_v32_nchw_to_nhwc");
    return DFU_OK;
}
extern "C" int dfu_v32_im2col(const float* x, float* col, int32_t B, int32_t H, int32_t W, int32_t C, int32_t R, int32_t S, int32_t stride,
                              int32_t pad, int32_t Ho, int32_t Wo, void* stream) {
    DFU_CHECK_ARG(x && col && B > 0 && H > 0 && W > 0 && C > 0 && R > 0 && S > 0 && stride > 0 && Ho > 0 && Wo > 0, "dfu_v32_im2col: bad args");
    v32_im2col_kernel<<<v32_grid((long long)B * Ho * Wo * R * S * C), 256, 0, V32_ST>>>(x, col, B, H, W, C, R, S, stride, pad, Ho, Wo);
    DFU_CHECK_LAUNCH("dfu_v32_im2col");
    return DFU_OK;
}
extern "C" int dfu_v32_patchify(const float* x_nhwc, float* out, int32_t B, void* stream) {
    DFU_CHECK_ARG(x_nhwc && out && B > 0, "dfu_v32_patchify: bad args");
    v32_patchify_kernel<<<v32_grid((long long)B * 196 * 768), 256, 0, V32_ST>>>(x_nhwc, out, B);
    DFU_CHECK_LAUNCH("dfu_v32_patchify");
    return DFU_OK;
}
extern "C" int dfu_v32_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int32_t cols, void* stream) {
    DFU_CHECK_ARG(src && dst && rows > 0 && cols > 0, "dfu_v32_copy2d: bad args");
    v32_copy2d_kernel<<<v32_grid(rows * cols), 256, 0, V32_ST>>>(src, lds, dst, ldd, rows, cols);
    DFU_CHECK_LAUNCH("dfu_v32_copy2d");
    return DFU_OK;
}
extern "C" int dfu_v32_round_bf16(float* x, int64_t n, void* stream) {
    DFU_CHECK_ARG(x && n > 0, "dfu_v32_round_bf16: bad args");
    v32_round_bf16_kernel<<<v32_grid(n), 256, 0, V32_ST>>>(x, n);
    DFU_CHECK_LAUNCH("dfu_v32_round_bf16");
    return DFU_OK;
}
extern "C" int dfu_v32_vit_assemble(const float* patch_out, const float* cls, const float* pos, float* tokens, int32_t B, int32_t T, int32_t D,
                                    void* stream) {
    DFU_CHECK_ARG(patch_out && cls && pos && tokens && B > 0 && T > 1 && D > 0, "dfu_v32_vit_assemble: bad args");
    v32_vit_assemble_kernel<<<v32_grid((long long)B * T * D), 256, 0, V32_ST>>>(patch_out, cls, pos, tokens, B, T, D);
    DFU_CHECK_LAUNCH("dfu_v32_vit_assemble");
    return DFU_OK;
}
extern "C" int dfu_v32_gemm(const dfu_v32_gemm_args* a, void* stream) {
    DFU_CHECK_ARG(a && a->A && a->B && a->D && a->M > 0 && a->N > 0 && a->K > 0 && a->nb0 > 0 && a->nb1 > 0, "dfu_v32_gemm: bad args");
    DFU_CHECK_ARG((long long)a->nb0 * a->nb1 <= 65535, "dfu_v32_gemm: too many batches (%d x %d)", a->nb0, a->nb1);
    V32Gemm p;
    p.M = a->M; p.N = a->N; p.K = a->K; p.nb0 = a->nb0; p.nb1 = a->nb1;
    p.A = a->A; p.B = a->B; p.bias = a->bias; p.res = a->res; p.D = a->D;
    p.a_rs = a->a_row_stride; p.a_cs = a->a_col_stride; p.a_s0 = a->sa0; p.a_s1 = a->sa1;
    p.b_rs = a->b_row_stride; p.b_cs = a->b_col_stride; p.b_s0 = a->sb0; p.b_s1 = a->sb1;
    p.ldd = a->ldd; p.d_s0 = a->sd0; p.d_s1 = a->sd1; p.ldr = a->ldr; p.r_s0 = a->sr0; p.r_s1 = a->sr1;
    p.alpha = a->alpha; p.act = a->act;
    dim3 grid(dfu_cdiv(a->N, VT), dfu_cdiv(a->M, VT), a->nb0 * a->nb1);
    DFU_CHECK_ARG(grid.y <= 65535, "dfu_v32_gemm: M too large for one launch (%d)", a->M);
    v32_gemm_kernel<<<grid, 256, 0, V32_ST>>>(p);
    DFU_CHECK_LAUNCH("dfu_v32_gemm");
    return DFU_OK;
}
extern "C" int dfu_v32_bn_batch_stats(const float* y, float* mean, float* var, int64_t M, int32_t C, void* stream) {
    DFU_CHECK_ARG(y && mean && var && M > 0 && C > 0, "dfu_v32_bn_batch_stats: bad args");
    v32_bn_batch_stats_kernel<<<dfu_cdiv(C, 32), 256, 0, V32_ST>>>(y, mean, var, M, C);
    DFU_CHECK_LAUNCH("dfu_v32_bn_batch_stats");
    return DFU_OK;
}
extern "C" int dfu_v32_bn_apply(const float* y, const float* mean, const float* var, const float* gamma, const float* beta, const float* res,
                                float* out, int64_t M, int32_t C, int32_t relu, float eps, void* stream) {
    DFU_CHECK_ARG(y && mean && var && gamma && beta && out && M > 0 && C > 0, "dfu_v32_bn_apply: bad args");
    v32_bn_apply_kernel<<<v32_grid(M * C), 256, 0, V32_ST>>>(y, mean, var, gamma, beta, res, out, M, C, relu, eps);
    DFU_CHECK_LAUNCH("dfu_v32_bn_apply");
    return DFU_OK;
}
extern "C" int dfu_v32_maxpool3x3s2(const float* x, float* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
    DFU_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0, "dfu_v32_maxpool3x3s2: bad args");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    v32_maxpool_kernel<<<v32_grid((long long)B * Ho * Wo * C), 256, 0, V32_ST>>>(x, y, B, H, W, C, Ho, Wo);
    DFU_CHECK_LAUNCH("dfu_v32_maxpool3x3s2");
    return DFU_OK;
}
extern "C" int dfu_v32_avgpool(const float* x, float* y, int32_t B, int32_t HW, int32_t C, void* stream) {
    DFU_CHECK_ARG(x && y && B > 0 && HW > 0 && C > 0, "dfu_v32_avgpool: bad args");
    v32_avgpool_kernel<<<v32_grid((long long)B * C), 256, 0, V32_ST>>>(x, y, B, HW, C);
    DFU_CHECK_LAUNCH("dfu_v32_avgpool");
    return DFU_OK;
}
extern "C" int dfu_v32_layernorm(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int32_t D, float eps, void* stream) {
    DFU_CHECK_ARG(x && gamma && beta && y && rows > 0 && D > 0, "dfu_v32_layernorm: bad args");
    v32_layernorm_kernel<<<dfu_cdiv(rows, 8), 256, 0, V32_ST>>>(x, gamma, beta, y, rows, D, eps);
    DFU_CHECK_LAUNCH("dfu_v32_layernorm");
    return DFU_OK;
}
extern "C" int dfu_v32_softmax_rows(float* x, int64_t rows, int32_t T, int64_t ld, void* stream) {
    DFU_CHECK_ARG(x && rows > 0 && T > 0 && ld >= T, "dfu_v32_softmax_rows: bad args");
    v32_softmax_rows_kernel<<<dfu_cdiv(rows, 8), 256, 0, V32_ST>>>(x, rows, T, ld);
    DFU_CHECK_LAUNCH("dfu_v32_softmax_rows");
    return DFU_OK;
}
