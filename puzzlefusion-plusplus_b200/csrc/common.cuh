// Shared helpers for the pfpp sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define PFPP_OK 0
#define PFPP_EINVAL (-1)
#define PFPP_EWORKSPACE (-2)
#define PFPP_EUNSUPPORTED (-3)

// All entry points are asynchronous on the caller's stream; launch-configuration errors are
// returned as positive cudaError_t values.
#define PFPP_RETURN_LAST()                         \
  do {                                             \
    cudaError_t e__ = cudaPeekAtLastError();       \
    return e__ == cudaSuccess ? PFPP_OK : (int)e__; \
  } while (0)

#define PFPP_CHECK_ARG(cond) \
  do {                       \
    if (!(cond)) return PFPP_EINVAL; \
  } while (0)

// Raise a kernel's dynamic shared-memory limit once (and again only if a larger size is requested): keeps
// cudaFuncSetAttribute out of the steady-state launch path and out of CUDA-graph capture.
#define PFPP_ENSURE_SMEM(kernel, bytes)                                                              \
  do {                                                                                               \
    static int cur__ = 48 * 1024;                                                                    \
    if ((int)(bytes) > cur__) {                                                                      \
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));       \
      cur__ = (int)(bytes);                                                                          \
    }                                                                                                \
  } while (0)

static inline int pfpp_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Individually-rounded fp32 arithmetic.  The discrete stages (FPS argmax, ball-query radius
// test, DDPM update) mirror the oracle op for op; nvcc must not contract these into FMAs.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// tanh-form GELU on the MUFU pipe, used ONLY where the result is immediately rounded to bf16 (GEGLU hidden of
// the tensor-core path): |gelu_tanh - gelu_erf| <= 5e-4 absolute, below half a bf16 ulp of the product it
// feeds; the fp32 (parity) kernels keep the exact erf form.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  float t;
  const float u = 0.7978845608028654f * fmaf(0.044715f * x, x * x, x);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

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
