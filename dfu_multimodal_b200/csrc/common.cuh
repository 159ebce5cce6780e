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

// Programmatic dependent launch (PDL).  Every kernel of the library is launched through dfu_launch(), i.e. with
// cudaLaunchAttributeProgrammaticStreamSerialization, and begins with pdl_sync(): `griddepcontrol.launch_dependents` lets
// the NEXT kernel of the stream be scheduled as soon as all CTAs of this one are resident (its CTAs take over SMs as this
// kernel's tail drains and run their prologue), `griddepcontrol.wait` blocks until the PREVIOUS kernel has completed and
// its writes are visible.  Every kernel waits before its first global-memory access, so the chain of waits keeps the
// stream's order for data (a kernel that passed its wait implies all earlier kernels have finished).  Meant to save the
// launch latency and prologue at each of the ~600 kernel boundaries of a train step.  MEASURED on B200 (profiles/r02c_*): no
// gain under CUDA-graph replay (31.76 ms with, 31.51 ms without), so the attribute is only set with DFU_PDL=1; without it
// the two instructions are no-ops.
bool dfu_pdl_enabled();

#ifdef __CUDACC__
#include <utility>

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
    pdl_trigger();
    pdl_wait();
}

template <typename... KArgs, typename... Args>
inline cudaError_t dfu_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = dfu_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue like one FFMA and do two lanes' worth of work).
// The GEMM epilogues are bound by the FMA pipe of the 8 epilogue warps, so the GELU polynomial runs two elements per op.
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_bcast(float c) { return f2_pack(c, c); }

// gelu_both_fast on two elements at once: (x0, x1) -> gelu in (g0, g1), derivative in (d0, d1).  Same arithmetic as the
// scalar version, except the normal CDF is formed as 0.5 + copysign(0.5 - r/2, x).
__device__ __forceinline__ void gelu_both_fast2(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const uint64_t x = f2_pack(x0, x1), ax = f2_pack(a0, a1);
    const uint64_t z = f2_mul(ax, f2_bcast(0.70710678118654752440f));
    uint64_t q = f2_fma(f2_bcast(0.0000430638f), z, f2_bcast(0.0002765672f));
    q = f2_fma(q, z, f2_bcast(0.0001520143f));
    q = f2_fma(q, z, f2_bcast(0.0092705272f));
    q = f2_fma(q, z, f2_bcast(0.0422820123f));
    q = f2_fma(q, z, f2_bcast(0.0705230784f));
    q = f2_fma(q, z, f2_bcast(1.0f));
    q = f2_mul(q, q); q = f2_mul(q, q); q = f2_mul(q, q); q = f2_mul(q, q);
    float q0, q1, r0, r1, e0, e1, t0, t1;
    f2_unpack(q, q0, q1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(q0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(q1));
    f2_unpack(f2_mul(f2_mul(x, x), f2_bcast(-0.72134752044448170368f)), t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    const uint64_t r = f2_pack(r0, r1), e = f2_pack(e0, e1);
    // gelu = max(x, 0) - 0.5 |x| r
    f2_unpack(f2_fma(f2_mul(ax, f2_bcast(-0.5f)), r, f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), g0, g1);
    // Phi = 0.5 + copysign(0.5 - 0.5 r, x);  gelu' = Phi + x phi(x)
    float h0, h1;
    f2_unpack(f2_fma(r, f2_bcast(-0.5f), f2_bcast(0.5f)), h0, h1);
    const uint64_t cdf = f2_add(f2_bcast(0.5f), f2_pack(copysignf(h0, x0), copysignf(h1, x1)));
    f2_unpack(f2_fma(f2_mul(x, f2_bcast(0.39894228040143267794f)), e, cdf), d0, d1);
}
// gelu_fast on two elements at once.
__device__ __forceinline__ void gelu_fast2(float x0, float x1, float& g0, float& g1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const uint64_t ax = f2_pack(a0, a1);
    const uint64_t z = f2_mul(ax, f2_bcast(0.70710678118654752440f));
    uint64_t q = f2_fma(f2_bcast(0.0000430638f), z, f2_bcast(0.0002765672f));
    q = f2_fma(q, z, f2_bcast(0.0001520143f));
    q = f2_fma(q, z, f2_bcast(0.0092705272f));
    q = f2_fma(q, z, f2_bcast(0.0422820123f));
    q = f2_fma(q, z, f2_bcast(0.0705230784f));
    q = f2_fma(q, z, f2_bcast(1.0f));
    q = f2_mul(q, q); q = f2_mul(q, q); q = f2_mul(q, q); q = f2_mul(q, q);
    float q0, q1, r0, r1;
    f2_unpack(q, q0, q1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(q0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(q1));
    f2_unpack(f2_fma(f2_mul(ax, f2_bcast(-0.5f)), f2_pack(r0, r1), f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), g0, g1);
}

#endif  // __CUDACC__
