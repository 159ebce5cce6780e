// Train-time augmentation on the device, fused with ToTensor + Normalize (+ ViT patch unfold):
//   RandomHorizontalFlip -> RandomVerticalFlip -> RandomRotation(30) -> [ColorJitter] -> [RandomAffine] -> ToTensor -> Normalize
// (notebooks/train_multimodal_fusion.py:173-199, train_thermal_only.py:103-112), which the reference runs on PIL images in
// four DataLoader worker processes.  GIVEN the drawn parameters the result is bit-identical to torchvision / Pillow:
//   * the geometric steps are NEAREST lookups through Pillow's 16.16 fixed-point inverse maps (Geometry.c affine_fixed; the
//     scale-only and right-angle fast paths are reproduced too), out-of-image pixels are black;
//   * the colour steps are Pillow's ImagingBlend in float32 without FMA contraction, rgb2l in integer arithmetic, and the
//     contrast mean is the rounded mean luma of the image AT THAT STAGE (pass 1, one integer sum per image).
// Every output pixel is one chain  affine^-1 -> rotation^-1 -> flips -> source pixel -> colour ops  evaluated in registers:
// the source image is read once (through L2 for the gathers), the output written once, no intermediate image exists.
#include "common.cuh"
#include "../../include/dfu_b200.h"

namespace {

struct Norm3 {
    float scale[3];
    float shift[3];
};

__device__ __forceinline__ int rgb2l(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Pillow Blend.c: out = d + f * (x - d) in float32 (separate multiply and add), truncated; clipped when f is outside [0, 1]
__device__ __forceinline__ int blend1(int d, int x, float f, bool inside) {
    const float t = __fadd_rn((float)d, __fmul_rn(f, (float)(x - d)));
    if (inside) return (int)t & 255;
    return t <= 0.f ? 0 : (t >= 255.f ? 255 : (int)t);
}

// inverse map of one geometric step: (y, x) of the step's output -> source coordinates, false when outside the image
__device__ __forceinline__ bool geo_map(int kind, const int* c, const short* xtab, const short* ytab, int H, int W, int& y, int& x) {
    int ys, xs;
    switch (kind) {
        case 0: return true;                                    // identity
        case 1:                                                 // Geometry.c affine_fixed
            xs = (c[2] + c[1] * y + c[0] * x) >> 16;
            ys = (c[5] + c[4] * y + c[3] * x) >> 16;
            break;
        case 2: xs = W - 1 - x; ys = H - 1 - y; break;          // rotate 180
        case 3: xs = W - 1 - y; ys = x; break;                  // rotate 90 (square image): out[y][x] = in[x][W-1-y]
        case 4: xs = y; ys = H - 1 - x; break;                  // rotate 270:               out[y][x] = in[H-1-x][y]
        default: xs = xtab[x]; ys = ytab[y]; break;             // 5: ImagingScaleAffine lookup tables (host-accumulated)
    }
    if (xs < 0 || xs >= W || ys < 0 || ys >= H) return false;
    x = xs;
    y = ys;
    return true;
}

// pixel of the image after flips + rotation, i.e. what the colour ops see at (y, x)
__device__ __forceinline__ void rotated_pixel(const uint8_t* __restrict__ img, const dfu_augment_params& p, int H, int W, int y, int x, int* rgb) {
    rgb[0] = rgb[1] = rgb[2] = 0;
    if (!geo_map(p.rot_kind, p.rot, nullptr, nullptr, H, W, y, x)) return;
    if (p.flags & 1) x = W - 1 - x;
    if (p.flags & 2) y = H - 1 - y;
    const uint8_t* s = img + ((long long)y * W + x) * 3;
    rgb[0] = s[0]; rgb[1] = s[1]; rgb[2] = s[2];
}

__device__ __forceinline__ void color_ops(const dfu_augment_params& p, int n_ops, int mean, int* rgb) {
    for (int k = 0; k < n_ops; ++k) {
        const float f = p.color_f[k];
        const bool inside = f >= 0.f && f <= 1.f;
        const int op = p.color_op[k];
        if (op == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[c] = blend1(0, rgb[c], f, inside);
        } else if (op == 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[c] = blend1(mean, rgb[c], f, inside);
        } else {
            const int l = rgb2l(rgb[0], rgb[1], rgb[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[c] = blend1(l, rgb[c], f, inside);
        }
    }
}

// pass 1: per image with a contrast op, the integer sum of the luma of the image as the contrast op will see it
__global__ void __launch_bounds__(256) augment_luma_sum_kernel(const uint8_t* __restrict__ src, const dfu_augment_params* __restrict__ params,
                                                               int H, int W, unsigned long long* __restrict__ sums) {
    pdl_sync();
    const int b = blockIdx.y;
    const dfu_augment_params& p = params[b];
    int before = -1;
    for (int k = 0; k < p.n_color; ++k)
        if (p.color_op[k] == 1) before = k;
    if (before < 0) return;
    const uint8_t* img = src + (long long)b * H * W * 3;
    unsigned long long local = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        int rgb[3];
        rotated_pixel(img, p, H, W, i / W, i % W, rgb);
        color_ops(p, before, 0, rgb);
        local += (unsigned)rgb2l(rgb[0], rgb[1], rgb[2]);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(sums + b, local);
}

// OUT: 0 = uint8 NHWC, 1 = normalised bf16 NHWC, 2 = normalised bf16 ViT patch matrix (224 x 224 only)
template <int OUT>
__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ src, const dfu_augment_params* __restrict__ params, int H,
                                                      int W, const unsigned long long* __restrict__ sums, void* __restrict__ out, Norm3 nm) {
    pdl_sync();
    const int b = blockIdx.y;
    const dfu_augment_params& p = params[b];
    const uint8_t* img = src + (long long)b * H * W * 3;
    // ImageStat mean of the luma image, rounded as ImageEnhance.Contrast does: int(sum / count + 0.5)
    const int mean = (int)((double)sums[b] / (double)(H * W) + 0.5);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int yo = i / W, xo = i % W;
        int y = yo, x = xo;
        int rgb[3] = {0, 0, 0};
        if (geo_map(p.aff_kind, p.aff, p.aff_xtab, p.aff_ytab, H, W, y, x)) {     // outside the affine map: fill 0, no colour ops
            rotated_pixel(img, p, H, W, y, x, rgb);
            color_ops(p, p.n_color, mean, rgb);
        }
        if (OUT == 0) {
            uint8_t* o = reinterpret_cast<uint8_t*>(out) + ((long long)b * H * W + i) * 3;
            o[0] = (uint8_t)rgb[0]; o[1] = (uint8_t)rgb[1]; o[2] = (uint8_t)rgb[2];
        } else if (OUT == 1) {
            bf16* o = reinterpret_cast<bf16*>(out) + ((long long)b * H * W + i) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = __float2bfloat16((float)rgb[c] * nm.scale[c] + nm.shift[c]);
        } else {
            const int ph = yo >> 4, pi = yo & 15, pw = xo >> 4, pj = xo & 15;
            bf16* o = reinterpret_cast<bf16*>(out) + ((long long)b * 196 + ph * 14 + pw) * 768 + pi * 16 + pj;
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c * 256] = __float2bfloat16((float)rgb[c] * nm.scale[c] + nm.shift[c]);
        }
    }
}

}  // namespace

extern "C" int dfu_augment_u8(const uint8_t* src, int32_t B, int32_t H, int32_t W, const dfu_augment_params* params, uint64_t* luma_sums,
                              int32_t out_mode, void* out, const float* mean3_host, const float* std3_host, void* stream) {
    DFU_CHECK_ARG(src && params && luma_sums && out && B > 0 && H > 0 && W > 0 && H <= 224 && W <= 224, "dfu_augment_u8: bad args (H, W <= 224)");
    DFU_CHECK_ARG(out_mode >= 0 && out_mode <= 2 && (out_mode == 0 || (mean3_host && std3_host)), "dfu_augment_u8: out_mode %d", out_mode);
    DFU_CHECK_ARG(out_mode != 2 || (H == 224 && W == 224), "dfu_augment_u8: the patch matrix needs 224x224 images");
    DFU_CHECK_ARG(B <= 65535, "dfu_augment_u8: batch too large for one launch (%d)", B);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(luma_sums, 0, (size_t)B * sizeof(uint64_t), st);
    if (e != cudaSuccess) { dfu_set_error("dfu_augment_u8: memset failed: %s", cudaGetErrorString(e)); return DFU_ERR_CUDA; }
    Norm3 nm;
    for (int c = 0; c < 3; ++c) {
        nm.scale[c] = out_mode ? 1.f / (255.f * std3_host[c]) : 0.f;
        nm.shift[c] = out_mode ? -mean3_host[c] / std3_host[c] : 0.f;
    }
    const dim3 grid(dfu_cdiv((long long)H * W, 256 * 4), B);
    dfu_launch(augment_luma_sum_kernel, grid, 256, 0, st, src, params, H, W, (unsigned long long*)luma_sums);
    if (out_mode == 0) dfu_launch(augment_kernel<0>, grid, 256, 0, st, src, params, H, W, (const unsigned long long*)luma_sums, out, nm);
    else if (out_mode == 1) dfu_launch(augment_kernel<1>, grid, 256, 0, st, src, params, H, W, (const unsigned long long*)luma_sums, out, nm);
    else dfu_launch(augment_kernel<2>, grid, 256, 0, st, src, params, H, W, (const unsigned long long*)luma_sums, out, nm);
    DFU_CHECK_LAUNCH("dfu_augment_u8");
    return DFU_OK;
}
