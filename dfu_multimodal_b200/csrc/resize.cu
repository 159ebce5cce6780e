// a9, first half: `transforms.Resize((224, 224))` of the reference pipelines (train_multimodal_fusion.py:174,186,195,204,
// extended_metrics.py:290-300) on the device, bit-identical to what the reference computes on the CPU, fused with
// ToTensor + Normalize (and, for the thermal branch, the ViT patch unfold).
//
// torchvision resizes a PIL image through Pillow's ImagingResample (BILINEAR): a separable triangle filter whose support
// grows with the down-scaling factor (antialiasing); coefficients are computed in float64 and converted to 22-bit fixed
// point; a horizontal pass and then a vertical pass EACH round to uint8.  Both passes are reproduced exactly:
//   resize_coeffs_kernel  -- the float64 coefficient table of one axis, with round-to-nearest intrinsics so that no FMA
//                            contraction changes a bit;
//   resize_u8_kernel      -- one CTA per (band of output rows, image): horizontal pass of the input rows the band needs
//                            into shared memory (uint8, as Pillow's temporary image), vertical pass from shared memory,
//                            then x/255 -> (x - mean)/std -> bf16 NHWC, the ViT patch matrix, or the uint8 image itself.
// HBM traffic: the source image is read once (plus 2*support rows of overlap per band), the output written once.
#include "common.cuh"
#include "../../include/dfu_b200.h"

namespace {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

struct Norm3 {
    float scale[3];   // 1/(255*std)
    float shift[3];   // -mean/std
};

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc, bilinear filter, full-image box.  One thread per output index.
__global__ void resize_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ coeffs) {
    pdl_sync();
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                              // 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int* k = coeffs + (long long)xx * ksize;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        ww = __dadd_rn(ww, w);
    }
    for (int x = 0; x < ksize; ++x) {
        int v = 0;
        if (x < xmax) {
            double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
            if (a < 0.0) a = -a;
            double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double f = __dmul_rn(w, (double)(1 << RS_PRECISION_BITS));
            v = w < 0.0 ? (int)__dadd_rn(-0.5, f) : (int)__dadd_rn(0.5, f);
        }
        k[x] = v;
    }
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
}

__device__ __forceinline__ int clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// OUT: 0 = uint8 NHWC (the resized image), 1 = normalised bf16 NHWC, 2 = normalised bf16 ViT patch matrix [B*196, 768]
template <int OUT>
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, long long img_stride, int Hs, int Ws,
                                                        const int* __restrict__ hbounds, const int* __restrict__ hcoef, int hks,
                                                        const int* __restrict__ vbounds, const int* __restrict__ vcoef, int vks,
                                                        int OH, int OW, int band, int max_rows, void* __restrict__ out, Norm3 nm) {
    pdl_sync();
    extern __shared__ uint8_t tmp[];                     // [rows of the band][OW*3]
    const int b = blockIdx.y;
    const int y0 = blockIdx.x * band, y1 = min(OH, y0 + band);
    const int rmin = vbounds[y0 * 2];
    const int rend = vbounds[(y1 - 1) * 2] + vbounds[(y1 - 1) * 2 + 1];
    const int nrows = rend - rmin;
    if (nrows > max_rows) __trap();                      // host sized the band from the same bounds formula
    const int rowlen = OW * 3;
    const uint8_t* img = src + (long long)b * img_stride;
    // horizontal pass: input rows rmin .. rend-1 -> tmp
    for (int idx = threadIdx.x; idx < nrows * rowlen; idx += blockDim.x) {
        const int r = idx / rowlen, e = idx - r * rowlen;
        const int xx = e / 3, c = e - xx * 3;
        const int xmin = __ldg(hbounds + xx * 2), n = __ldg(hbounds + xx * 2 + 1);
        const int* k = hcoef + (long long)xx * hks;
        const uint8_t* p = img + ((long long)(rmin + r) * Ws + xmin) * 3 + c;
        int ss = 1 << (RS_PRECISION_BITS - 1);
        for (int x = 0; x < n; ++x) ss += (int)__ldg(p + x * 3) * __ldg(k + x);
        tmp[idx] = (uint8_t)clip8(ss);
    }
    __syncthreads();
    // vertical pass + output
    for (int idx = threadIdx.x; idx < (y1 - y0) * rowlen; idx += blockDim.x) {
        const int ry = idx / rowlen, e = idx - ry * rowlen;
        const int yy = y0 + ry;
        const int ymin = __ldg(vbounds + yy * 2), n = __ldg(vbounds + yy * 2 + 1);
        const int* k = vcoef + (long long)yy * vks;
        const uint8_t* p = tmp + (ymin - rmin) * rowlen + e;
        int ss = 1 << (RS_PRECISION_BITS - 1);
        for (int y = 0; y < n; ++y) ss += (int)p[y * rowlen] * __ldg(k + y);
        const int v = clip8(ss);
        const int xx = e / 3, c = e - xx * 3;
        if (OUT == 0) {
            reinterpret_cast<uint8_t*>(out)[((long long)b * OH + yy) * rowlen + e] = (uint8_t)v;
        } else if (OUT == 1) {
            reinterpret_cast<bf16*>(out)[((long long)b * OH + yy) * rowlen + e] = __float2bfloat16((float)v * nm.scale[c] + nm.shift[c]);
        } else {
            const int ph = yy >> 4, i = yy & 15, pw = xx >> 4, j = xx & 15;
            reinterpret_cast<bf16*>(out)[((long long)b * 196 + ph * 14 + pw) * 768 + c * 256 + i * 16 + j] =
                __float2bfloat16((float)v * nm.scale[c] + nm.shift[c]);
        }
    }
}

inline int resize_ksize(int in_size, int out_size) {
    double scale = (double)in_size / (double)out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)ceil(scale) * 2 + 1;
}

}  // namespace

extern "C" int dfu_resize_workspace_ints(int32_t Hs, int32_t Ws, int32_t OH, int32_t OW) {
    if (Hs <= 0 || Ws <= 0 || OH <= 0 || OW <= 0) return -1;
    return OW * 2 + OW * resize_ksize(Ws, OW) + OH * 2 + OH * resize_ksize(Hs, OH);
}

extern "C" int dfu_resize_u8(const uint8_t* src, int64_t img_stride_bytes, int32_t B, int32_t Hs, int32_t Ws, int32_t OH, int32_t OW,
                             int32_t out_mode, void* out, int32_t* workspace, const float* mean3_host, const float* std3_host,
                             void* stream) {
    DFU_CHECK_ARG(src && out && workspace && B > 0 && Hs > 0 && Ws > 0 && OH > 0 && OW > 0, "dfu_resize_u8: bad args");
    DFU_CHECK_ARG(out_mode >= 0 && out_mode <= 2, "dfu_resize_u8: out_mode %d", out_mode);
    DFU_CHECK_ARG(out_mode == 0 || (mean3_host && std3_host), "dfu_resize_u8: normalisation constants missing");
    DFU_CHECK_ARG(out_mode != 2 || (OH == 224 && OW == 224), "dfu_resize_u8: the patch matrix needs a 224x224 output");
    DFU_CHECK_ARG(B <= 65535, "dfu_resize_u8: batch too large for one launch (%d)", B);
    cudaStream_t st = (cudaStream_t)stream;
    const int hks = resize_ksize(Ws, OW), vks = resize_ksize(Hs, OH);
    int* hbounds = workspace;
    int* hcoef = hbounds + OW * 2;
    int* vbounds = hcoef + OW * hks;
    int* vcoef = vbounds + OH * 2;
    dfu_launch(resize_coeffs_kernel, dfu_cdiv(OW, 128), 128, 0, st, Ws, OW, hks, hbounds, hcoef);
    dfu_launch(resize_coeffs_kernel, dfu_cdiv(OH, 128), 128, 0, st, Hs, OH, vks, vbounds, vcoef);
    DFU_CHECK_LAUNCH("dfu_resize_u8 coefficients");
    // band of output rows per CTA: the largest whose input rows (band * scale + 2 * support + 2, the bound of Pillow's
    // xmin/xmax formulas) fit in shared memory
    const double vscale = (double)Hs / (double)OH;
    const double vsupport = vscale < 1.0 ? 1.0 : vscale;
    const int rowlen = OW * 3;
    int band = 16, max_rows = 0;
    for (; band >= 1; band >>= 1) {
        max_rows = (int)ceil(band * vscale + 2.0 * vsupport) + 3;
        if (max_rows > Hs) max_rows = Hs;
        if ((size_t)max_rows * rowlen <= 200 * 1024) break;
    }
    DFU_CHECK_ARG(band >= 1, "dfu_resize_u8: a single output row needs %d input rows of %d bytes: exceeds shared memory", max_rows, rowlen);
    const size_t smem = (size_t)max_rows * rowlen;
    Norm3 nm;
    for (int c = 0; c < 3; ++c) {
        nm.scale[c] = out_mode ? 1.f / (255.f * std3_host[c]) : 0.f;
        nm.shift[c] = out_mode ? -mean3_host[c] / std3_host[c] : 0.f;
    }
    dim3 grid(dfu_cdiv(OH, band), B);
#define RS_LAUNCH(MODE)                                                                                                            \
    do {                                                                                                                           \
        if (smem > 48 * 1024) cudaFuncSetAttribute(resize_u8_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        dfu_launch(resize_u8_kernel<MODE>, grid, 256, smem, st, src, img_stride_bytes, Hs, Ws, hbounds, hcoef, hks, vbounds, vcoef, vks,   \
                                                        OH, OW, band, max_rows, out, nm);                                          \
    } while (0)
    if (out_mode == 0) RS_LAUNCH(0);
    else if (out_mode == 1) RS_LAUNCH(1);
    else RS_LAUNCH(2);
#undef RS_LAUNCH
    DFU_CHECK_LAUNCH("dfu_resize_u8");
    return DFU_OK;
}
