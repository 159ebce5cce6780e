// Shared device/host helpers for the dfu_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DFU_OK 0
#define DFU_ERR_ARG 1
#define DFU_ERR_CUDA 2
#define DFU_ERR_UNSUPPORTED 3

// Last error text (thread-local), readable through dfu_last_error().
void dfu_set_error(const char* fmt, ...);

#define DFU_CHECK_ARG(cond, ...)                \
    do {                                        \
        if (!(cond)) {                          \
            dfu_set_error(__VA_ARGS__);         \
            return DFU_ERR_ARG;                 \
        }                                       \
    } while (0)

#define DFU_CHECK_LAUNCH(name)                                                   \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            dfu_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return DFU_ERR_CUDA;                                                 \
        }                                                                        \
    } while (0)

static inline int dfu_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Number of SMs of the current device (cached).
int dfu_num_sms();

#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `red` is a shared array of >= 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// 8 x bf16 <-> 8 x float through one 128-bit access.
struct alignas(16) bf16x8 {
    bf162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(p.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
    bf16x8 p;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return p;
}
__device__ __forceinline__ bf16x8 ld8(const bf16* p) { return *reinterpret_cast<const bf16x8*>(p); }
__device__ __forceinline__ void st8(bf16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// exact-erf GELU through Abramowitz-Stegun 7.1.28 (|erf error| <= 3e-7): one MUFU.RCP and ~14 FP32 ops per element
// instead of erff(); gelu(x) = relu(x) - 0.5 |x| / (1 + a1 z + ... + a6 z^6)^16 with z = |x| / sqrt(2).
__device__ __forceinline__ float gelu_fast(float x) {
    const float ax = fabsf(x);
    const float z = ax * 0.70710678118654752440f;
    float q = 0.0000430638f;
    q = fmaf(q, z, 0.0002765672f);
    q = fmaf(q, z, 0.0001520143f);
    q = fmaf(q, z, 0.0092705272f);
    q = fmaf(q, z, 0.0422820123f);
    q = fmaf(q, z, 0.0705230784f);
    q = fmaf(q, z, 1.0f);
    q = q * q; q = q * q; q = q * q; q = q * q;
    float r;                               // q = inf for large |x| -> r = 0
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
    return fmaf(-0.5f * ax, r, fmaxf(x, 0.f));
}
// d/dx gelu(x) = Phi(x) + x phi(x), same erf approximation (Phi = 1 - r/2 for x >= 0, r/2 otherwise).
__device__ __forceinline__ float gelu_grad_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float q = 0.0000430638f;
    q = fmaf(q, z, 0.0002765672f);
    q = fmaf(q, z, 0.0001520143f);
    q = fmaf(q, z, 0.0092705272f);
    q = fmaf(q, z, 0.0422820123f);
    q = fmaf(q, z, 0.0705230784f);
    q = fmaf(q, z, 1.0f);
    q = q * q; q = q * q; q = q * q; q = q * q;
    float r, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));
    const float cdf = x >= 0.f ? fmaf(-0.5f, r, 1.f) : 0.5f * r;
    return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// gelu(x) and d/dx gelu(x) together (shared polynomial and reciprocal): g = gelu_fast(x), d = gelu_grad_fast(x).
__device__ __forceinline__ void gelu_both_fast(float x, float& g, float& d) {
    const float ax = fabsf(x);
    const float z = ax * 0.70710678118654752440f;
    float q = 0.0000430638f;
    q = fmaf(q, z, 0.0002765672f);
    q = fmaf(q, z, 0.0001520143f);
    q = fmaf(q, z, 0.0092705272f);
    q = fmaf(q, z, 0.0422820123f);
    q = fmaf(q, z, 0.0705230784f);
    q = fmaf(q, z, 1.0f);
    q = q * q; q = q * q; q = q * q; q = q * q;
    float r, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));
    g = fmaf(-0.5f * ax, r, fmaxf(x, 0.f));
    const float cdf = x >= 0.f ? fmaf(-0.5f, r, 1.f) : 0.5f * r;
    d = fmaf(x * 0.39894228040143267794f, e, cdf);
}

#endif  // __CUDACC__
