// Loss, prediction, AdamW, Grad-CAM reduction, TTA vote.
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

// ------------------------------------------------------------------ weighted cross entropy (single block)
__global__ void __launch_bounds__(256) cross_entropy_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                            const float* __restrict__ weight, float* __restrict__ loss,
                                                            float* __restrict__ dlogits, int B, int C) {
    pdl_sync();
    __shared__ float red[32];
    float wsum = 0.f, lsum = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float* row = logits + (long long)b * C;
        const int y = (int)labels[b];
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, row[c]);
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
        const float lse = mx + logf(se);
        const float w = weight ? weight[y] : 1.f;
        wsum += w;
        lsum += w * (lse - row[y]);
    }
    wsum = block_sum(wsum, red);
    lsum = block_sum(lsum, red);
    if (threadIdx.x == 0 && loss) *loss = lsum / wsum;
    if (dlogits) {
        const float inv = 1.f / wsum;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const float* row = logits + (long long)b * C;
            const int y = (int)labels[b];
            float mx = -INFINITY;
            for (int c = 0; c < C; ++c) mx = fmaxf(mx, row[c]);
            float se = 0.f;
            for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
            const float w = (weight ? weight[y] : 1.f) * inv;
            for (int c = 0; c < C; ++c) {
                const float p = expf(row[c] - mx) / se;
                dlogits[(long long)b * C + c] = w * (p - (c == y ? 1.f : 0.f));
            }
        }
    }
}

__global__ void __launch_bounds__(256) predict_kernel(const float* __restrict__ logits, float* __restrict__ prob1,
                                                      long long* __restrict__ pred, int B, int C) {
    pdl_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* row = logits + (long long)b * C;
    float mx = row[0];
    int arg = 0;
    for (int c = 1; c < C; ++c)
        if (row[c] > mx) { mx = row[c]; arg = c; }
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
    if (prob1) prob1[b] = (C > 1) ? expf(row[1] - mx) / se : 1.f / (1.f + expf(-row[0]));
    if (pred) pred[b] = arg;
}

// out[i] = a[i] * (*s)   (fp32; upstream scalar gradient of the loss)
__global__ void __launch_bounds__(256) scale_by_scalar_kernel(const float* __restrict__ a, const float* __restrict__ s,
                                                              float* __restrict__ out, long long n) {
    pdl_sync();
    const float k = *s;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] * k;
}

// ------------------------------------------------------------------ AdamW
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps,
                                          float wd, float bc1, float bc2_sqrt) {
    // torch.optim.AdamW (single-tensor path): p *= 1 - lr*wd; m,v update; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
    p *= 1.f - lr * wd;
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= (lr / bc1) * (m / denom);
}

template <bool G_BF16>
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const void* __restrict__ g_, float* __restrict__ m,
                                                         float* __restrict__ v, bf16* __restrict__ shadow, long long n,
                                                         const float* __restrict__ step, float lr, float b1, float b2,
                                                         float eps, float wd, float gscale) {
    pdl_sync();
    const float t = *step;
    const float bc1 = 1.f - powf(b1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
    const long long nv = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float gg[4];
        if (G_BF16) {
            const uint2 raw = reinterpret_cast<const uint2*>(g_)[i];
            const float2 a = __bfloat1622float2(*reinterpret_cast<const bf162*>(&raw.x));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const bf162*>(&raw.y));
            gg[0] = a.x; gg[1] = a.y; gg[2] = b.x; gg[3] = b.y;
        } else {
            const float4 g4 = reinterpret_cast<const float4*>(g_)[i];
            gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
        }
        adamw_one(pp.x, gg[0] * gscale, mm.x, vv.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        adamw_one(pp.y, gg[1] * gscale, mm.y, vv.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        adamw_one(pp.z, gg[2] * gscale, mm.z, vv.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        adamw_one(pp.w, gg[3] * gscale, mm.w, vv.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (shadow) {
            bf162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo);
            o.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(shadow)[i] = o;
        }
    }
    if (blockIdx.x == 0) {
        for (long long i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const float g = (G_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(g_)[i]) : reinterpret_cast<const float*>(g_)[i]) * gscale;
            adamw_one(p[i], g, m[i], v[i], lr, b1, b2, eps, wd, bc1, bc2_sqrt);
            if (shadow) shadow[i] = __float2bfloat16(p[i]);
        }
    }
}

// Multi-tensor fp32 -> bf16 cast: one block per chunk of `chunk_size` elements of one tensor (the gradients of a whole
// data-parallel bucket go onto the wire with ONE launch instead of one per parameter).
__global__ void __launch_bounds__(256) cast_multi_kernel(const dfu_cast_tensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                                                         const int* __restrict__ chunk_offset, int chunk_size) {
    pdl_sync();
    const dfu_cast_tensor t = table[chunk_tensor[blockIdx.x]];
    const long long start = (long long)chunk_offset[blockIdx.x] * chunk_size;
    const long long end = min(start + chunk_size, (long long)t.numel);
    const float* src = t.src;
    bf16* dst = reinterpret_cast<bf16*>(t.dst);
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const long long v0 = start >> 3, v1 = end >> 3;     // chunk_size is a multiple of 8
        for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src) + i * 2);
            const float4 b = __ldg(reinterpret_cast<const float4*>(src) + i * 2 + 1);
            const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            st8(dst + i * 8, pack8(f));
        }
        for (long long i = (v1 << 3) + threadIdx.x; i < end; i += blockDim.x) dst[i] = __float2bfloat16(src[i]);
    } else {
        for (long long i = start + threadIdx.x; i < end; i += blockDim.x) dst[i] = __float2bfloat16(src[i]);
    }
}

// Multi-tensor: one block per chunk of `chunk_size` elements of one tensor.
template <bool G_BF16>
__global__ void __launch_bounds__(256) adamw_multi_kernel(const dfu_adamw_tensor* __restrict__ table,
                                                          const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_offset,
                                                          int chunk_size, const float* __restrict__ step, float lr, float b1,
                                                          float b2, float eps, float wd, float gscale) {
    pdl_sync();
    const dfu_adamw_tensor t = table[chunk_tensor[blockIdx.x]];
    const long long start = (long long)chunk_offset[blockIdx.x] * chunk_size;
    const long long end = min(start + chunk_size, (long long)t.numel);
    const float ts = *step;
    const float bc1 = 1.f - powf(b1, ts);
    const float bc2_sqrt = sqrtf(1.f - powf(b2, ts));
    bf16* shadow = reinterpret_cast<bf16*>(t.shadow);
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(t.g) & (G_BF16 ? 7 : 15)) == 0 &&
                     (!shadow || (reinterpret_cast<uintptr_t>(shadow) & 7) == 0);
    if (vec) {
        const long long v0 = start >> 2, v1 = end >> 2;     // chunk_size is a multiple of 4
        for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) {
            float4 pp = reinterpret_cast<float4*>(t.p)[i];
            float4 mm = reinterpret_cast<float4*>(t.m)[i];
            float4 vv = reinterpret_cast<float4*>(t.v)[i];
            float gg[4];
            if (G_BF16) {
                const uint2 raw = reinterpret_cast<const uint2*>(t.g)[i];
                const float2 a = __bfloat1622float2(*reinterpret_cast<const bf162*>(&raw.x));
                const float2 b = __bfloat1622float2(*reinterpret_cast<const bf162*>(&raw.y));
                gg[0] = a.x; gg[1] = a.y; gg[2] = b.x; gg[3] = b.y;
            } else {
                const float4 g4 = reinterpret_cast<const float4*>(t.g)[i];
                gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
            }
            adamw_one(pp.x, gg[0] * gscale, mm.x, vv.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
            adamw_one(pp.y, gg[1] * gscale, mm.y, vv.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
            adamw_one(pp.z, gg[2] * gscale, mm.z, vv.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
            adamw_one(pp.w, gg[3] * gscale, mm.w, vv.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
            reinterpret_cast<float4*>(t.p)[i] = pp;
            reinterpret_cast<float4*>(t.m)[i] = mm;
            reinterpret_cast<float4*>(t.v)[i] = vv;
            if (shadow) {
                bf162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
                uint2 o;
                o.x = *reinterpret_cast<uint32_t*>(&lo);
                o.y = *reinterpret_cast<uint32_t*>(&hi);
                reinterpret_cast<uint2*>(shadow)[i] = o;
            }
        }
        for (long long i = (v1 << 2) + threadIdx.x; i < end; i += blockDim.x) {
            const float g = (G_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(t.g)[i]) : reinterpret_cast<const float*>(t.g)[i]) * gscale;
            adamw_one(t.p[i], g, t.m[i], t.v[i], lr, b1, b2, eps, wd, bc1, bc2_sqrt);
            if (shadow) shadow[i] = __float2bfloat16(t.p[i]);
        }
    } else {
        for (long long i = start + threadIdx.x; i < end; i += blockDim.x) {
            const float g = (G_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(t.g)[i]) : reinterpret_cast<const float*>(t.g)[i]) * gscale;
            adamw_one(t.p[i], g, t.m[i], t.v[i], lr, b1, b2, eps, wd, bc1, bc2_sqrt);
            if (shadow) shadow[i] = __float2bfloat16(t.p[i]);
        }
    }
}
__global__ void tick_f32_kernel(float* s) {
    pdl_sync(); *s += 1.f; }

// ------------------------------------------------------------------ Grad-CAM reduction (single block)
// w_c = mean_hw G[hw, c]; cam[hw] = relu(sum_{c<nc} w_c * A[hw, c]); cam /= max if max > 0.
__global__ void __launch_bounds__(256) gradcam_kernel(const bf16* __restrict__ A, int CA, const bf16* __restrict__ G, int CG, int HW,
                                                      int nc, float* __restrict__ cam) {
    pdl_sync();
    extern __shared__ float sh[];           // w[nc] then cam[HW]
    __shared__ float red[32];
    float* w = sh;
    float* cm = sh + nc;
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += __bfloat162float(G[(long long)p * CG + c]);
        w[c] = s / (float)HW;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int p = wid; p < HW; p += nw) {
        float s = 0.f;
        for (int c = lane; c < nc; c += 32) s += w[c] * __bfloat162float(A[(long long)p * CA + c]);
        s = warp_sum(s);
        if (lane == 0) cm[p] = fmaxf(s, 0.f);
    }
    __syncthreads();
    float mx = 0.f;
    for (int p = threadIdx.x; p < HW; p += blockDim.x) mx = fmaxf(mx, cm[p]);
    mx = block_max(mx, red);
    const float inv = mx > 0.f ? 1.f / mx : 1.f;
    for (int p = threadIdx.x; p < HW; p += blockDim.x) cam[p] = cm[p] * inv;
}

// saliency: out[hw] = mean_c |d[c, hw]|, then / max (two kernels: the max is grid wide)
__global__ void __launch_bounds__(256) saliency_kernel(const float* __restrict__ d, float* __restrict__ out, float* __restrict__ mx,
                                                       int C, int HW) {
    pdl_sync();
    __shared__ float red[32];
    float local = 0.f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += fabsf(d[(long long)c * HW + p]);
        s /= (float)C;
        out[p] = s;
        local = fmaxf(local, s);
    }
    local = block_max(local, red);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(mx), __float_as_int(local));   // values >= 0
}
__global__ void __launch_bounds__(256) saliency_norm_kernel(float* __restrict__ out, const float* __restrict__ mx, int HW) {
    pdl_sync();
    const float m = *mx;
    if (!(m > 0.f)) return;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) out[p] /= m;
}

__global__ void __launch_bounds__(256) tta_vote_kernel(const float* __restrict__ probs, float* __restrict__ mean_prob,
                                                       long long* __restrict__ pred, int n, int V) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f, votes = 0.f;
    for (int v = 0; v < V; ++v) {
        const float p = probs[(long long)i * V + v];
        s += p;
        votes += (p > 0.5f) ? 1.f : 0.f;
    }
    mean_prob[i] = s / (float)V;
    pred[i] = (votes / (float)V > 0.5f) ? 1 : 0;
}

}  // namespace

extern "C" int dfu_cross_entropy(const float* logits, const int64_t* labels, const float* weight, float* loss, float* dlogits,
                                 int32_t B, int32_t C, void* stream) {
    DFU_CHECK_ARG(logits && labels && B > 0 && C > 0 && (loss || dlogits), "dfu_cross_entropy: bad args");
    dfu_launch(cross_entropy_kernel, 1, 256, 0, (cudaStream_t)stream, logits, (const long long*)labels, weight, loss, dlogits, B, C);
    DFU_CHECK_LAUNCH("dfu_cross_entropy");
    return DFU_OK;
}
extern "C" int dfu_predict(const float* logits, float* prob1, int64_t* pred, int32_t B, int32_t C, void* stream) {
    DFU_CHECK_ARG(logits && B > 0 && C > 0, "dfu_predict: bad args");
    dfu_launch(predict_kernel, dfu_cdiv(B, 256), 256, 0, (cudaStream_t)stream, logits, prob1, (long long*)pred, B, C);
    DFU_CHECK_LAUNCH("dfu_predict");
    return DFU_OK;
}
extern "C" int dfu_scale_by_scalar(const float* a, const float* scalar_dev, float* out, int64_t n, void* stream) {
    DFU_CHECK_ARG(a && scalar_dev && out && n > 0, "dfu_scale_by_scalar: bad args");
    dfu_launch(scale_by_scalar_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, a, scalar_dev, out, n);
    DFU_CHECK_LAUNCH("dfu_scale_by_scalar");
    return DFU_OK;
}

extern "C" int dfu_adamw_tick(float* step, void* stream) {
    DFU_CHECK_ARG(step, "dfu_adamw_tick: null");
    dfu_launch(tick_f32_kernel, 1, 1, 0, (cudaStream_t)stream, step);
    DFU_CHECK_LAUNCH("dfu_adamw_tick");
    return DFU_OK;
}
extern "C" int dfu_adamw(const dfu_adamw_tensor* table, int32_t n_tensors, const int32_t* chunk_tensor, const int32_t* chunk_offset,
                         int32_t n_chunks, int32_t chunk_size, const float* step, float lr, float beta1, float beta2, float eps,
                         float weight_decay, float grad_scale, int32_t g_bf16, void* stream) {
    DFU_CHECK_ARG(table && n_tensors > 0 && chunk_tensor && chunk_offset && n_chunks > 0 && chunk_size > 0 && chunk_size % 4 == 0 && step,
                  "dfu_adamw: bad args");
    if (g_bf16)
        dfu_launch(adamw_multi_kernel<true>, n_chunks, 256, 0, (cudaStream_t)stream, table, chunk_tensor, chunk_offset, chunk_size, step, lr,
                                                                            beta1, beta2, eps, weight_decay, grad_scale);
    else
        dfu_launch(adamw_multi_kernel<false>, n_chunks, 256, 0, (cudaStream_t)stream, table, chunk_tensor, chunk_offset, chunk_size, step, lr,
                                                                             beta1, beta2, eps, weight_decay, grad_scale);
    DFU_CHECK_LAUNCH("dfu_adamw");
    return DFU_OK;
}
extern "C" int dfu_cast_multi_f32_bf16(const dfu_cast_tensor* table, int32_t n_tensors, const int32_t* chunk_tensor,
                                       const int32_t* chunk_offset, int32_t n_chunks, int32_t chunk_size, void* stream) {
    DFU_CHECK_ARG(table && n_tensors > 0 && chunk_tensor && chunk_offset && n_chunks > 0 && chunk_size > 0 && chunk_size % 8 == 0,
                  "dfu_cast_multi_f32_bf16: bad args");
    dfu_launch(cast_multi_kernel, n_chunks, 256, 0, (cudaStream_t)stream, table, chunk_tensor, chunk_offset, chunk_size);
    DFU_CHECK_LAUNCH("dfu_cast_multi_f32_bf16");
    return DFU_OK;
}
extern "C" int dfu_adamw_flat(float* p, const void* g, float* m, float* v, void* shadow, int64_t n, const float* step, float lr,
                              float beta1, float beta2, float eps, float weight_decay, float grad_scale, int32_t g_bf16, void* stream) {
    DFU_CHECK_ARG(p && g && m && v && n > 0 && step, "dfu_adamw_flat: bad args");
    DFU_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)shadow) & 15) == 0,
                  "dfu_adamw_flat: buffers must be 16-byte aligned");
    const int grid = grid_for((n + 3) / 4, 256, 8);
    if (g_bf16)
        dfu_launch(adamw_flat_kernel<true>, grid, 256, 0, (cudaStream_t)stream, p, g, m, v, (bf16*)shadow, n, step, lr, beta1, beta2, eps,
                                                                       weight_decay, grad_scale);
    else
        dfu_launch(adamw_flat_kernel<false>, grid, 256, 0, (cudaStream_t)stream, p, g, m, v, (bf16*)shadow, n, step, lr, beta1, beta2, eps,
                                                                        weight_decay, grad_scale);
    DFU_CHECK_LAUNCH("dfu_adamw_flat");
    return DFU_OK;
}

extern "C" int dfu_gradcam(const void* A, int32_t CA, const void* G, int32_t CG, int32_t HW, int32_t nc, float* cam, void* stream) {
    DFU_CHECK_ARG(A && G && cam && HW > 0 && nc > 0 && nc <= CA && nc <= CG, "dfu_gradcam: bad args");
    const size_t smem = (size_t)(nc + HW) * sizeof(float);
    DFU_CHECK_ARG(smem <= 48 * 1024, "dfu_gradcam: nc+HW too large (%d + %d)", nc, HW);
    dfu_launch(gradcam_kernel, 1, 256, smem, (cudaStream_t)stream, (const bf16*)A, CA, (const bf16*)G, CG, HW, nc, cam);
    DFU_CHECK_LAUNCH("dfu_gradcam");
    return DFU_OK;
}
extern "C" int dfu_saliency(const float* dimg, float* out, float* max_scratch, int32_t C, int32_t HW, void* stream) {
    DFU_CHECK_ARG(dimg && out && max_scratch && C > 0 && HW > 0, "dfu_saliency: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(max_scratch, 0, sizeof(float), st);
    const int grid = grid_for(HW, 256, 2);
    dfu_launch(saliency_kernel, grid, 256, 0, st, dimg, out, max_scratch, C, HW);
    DFU_CHECK_LAUNCH("dfu_saliency");
    dfu_launch(saliency_norm_kernel, grid, 256, 0, st, out, max_scratch, HW);
    DFU_CHECK_LAUNCH("dfu_saliency_norm");
    return DFU_OK;
}
extern "C" int dfu_tta_vote(const float* probs, float* mean_prob, int64_t* pred, int32_t n, int32_t V, void* stream) {
    DFU_CHECK_ARG(probs && mean_prob && pred && n > 0 && V > 0, "dfu_tta_vote: bad args");
    dfu_launch(tta_vote_kernel, dfu_cdiv(n, 256), 256, 0, (cudaStream_t)stream, probs, mean_prob, (long long*)pred, n, V);
    DFU_CHECK_LAUNCH("dfu_tta_vote");
    return DFU_OK;
}
