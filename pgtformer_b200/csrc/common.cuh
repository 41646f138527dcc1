// Shared host/device helpers for libpgt_b200.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pgt_b200.h"

namespace pgt {

// ---- host-side status plumbing (api.cu)
void set_cuda_error(cudaError_t e, const char* where);
void count_launch(int n = 1);
int num_sms();
bool prof_enabled();
void prof_before(int cls, double work, cudaStream_t st, const char* desc);
void prof_after(cudaStream_t st);
struct ProfScope {      // RAII: events around one launch when the profiler is on
  cudaStream_t st; bool on;
  ProfScope(int cls, double work, cudaStream_t s, const char* desc = nullptr) : st(s), on(prof_enabled()) { if (on) prof_before(cls, work, s, desc); }
  ~ProfScope() { if (on) prof_after(st); }
};

#define PGT_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return PGT_ERR_INVALID; \
  } while (0)

#define PGT_CUDA_OK(expr)                         \
  do {                                            \
    cudaError_t _e = (expr);                      \
    if (_e != cudaSuccess) {                      \
      ::pgt::set_cuda_error(_e, #expr);           \
      return PGT_ERR_CUDA;                        \
    }                                             \
  } while (0)

#define PGT_LAUNCH_OK()                           \
  do {                                            \
    cudaError_t _e = cudaGetLastError();          \
    if (_e != cudaSuccess) {                      \
      ::pgt::set_cuda_error(_e, "kernel launch"); \
      return PGT_ERR_CUDA;                        \
    }                                             \
    ::pgt::count_launch();                        \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device math
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// exact-erf GELU, v * Phi(v), with erfc from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7):
//   w = erfc(|v| / sqrt 2) = t * poly(t) * exp(-v^2 / 2),  t = 1 / (1 + p |v| / sqrt 2)
//   gelu(v) = relu(v) - 0.5 |v| w          (v >= 0: v (1 - w/2);  v < 0: v w / 2)
// 2 MUFU + 12 FP32 ops, no branches — the multi-branch libdevice erff dominated the GEMM epilogue
__device__ __forceinline__ float gelu_erf(float v) {
  const float az = fabsf(v);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, az, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_approx(v * v * (-0.5f * 1.4426950408889634f));
  const float w = poly * t * e;
  return fmaf(az * -0.5f, w, fmaxf(v, 0.f));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case PGT_ACT_GELU: return gelu_erf(v);
    case PGT_ACT_SILU: return v * rcp_approx(1.0f + ex2_approx(v * -1.4426950408889634f));
    case PGT_ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
    case PGT_ACT_RELU: return fmaxf(v, 0.f);
    case PGT_ACT_SIGMOID: return rcp_approx(1.0f + ex2_approx(v * -1.4426950408889634f));
    default: return v;
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ float bf16_to_f(const __nv_bfloat16* p) { return __bfloat162float(*p); }

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

}  // namespace pgt
