// Shared helpers for the neko_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/neko_b200.h"

namespace neko {

// thread-local error string behind neko_last_error()
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define NEKO_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::neko::set_error(__VA_ARGS__);      \
      return NEKO_EINVAL;                  \
    }                                      \
  } while (0)

#define NEKO_LAUNCH_CHECK(name)                                   \
  do {                                                            \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return ::neko::check_cuda(_e, name);   \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

typedef __nv_bfloat16 bf16;

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

// erf with |error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26): one rcp, one ex2, a degree-5 Horner chain -- a
// fraction of libdevice erff(); well below the rounding of the 16-bit outputs it feeds.
__device__ __forceinline__ float erf_fast(float x) {
  const float a = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, a, 1.0f));  // MUFU.RCP, ~1 ulp
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float y = 1.0f - p * t * __expf(-a * a);
  return copysignf(y, x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erf_fast(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// fp16 conversions saturate at the largest finite half instead of overflowing to inf
__device__ __forceinline__ float sat_f16(float v) { return fminf(fmaxf(v, -65504.0f), 65504.0f); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(sat_f16(lo), sat_f16(hi));
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-bit pair in the requested format (f16 = IEEE half, else bfloat16)
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi, bool f16) { return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
__device__ __forceinline__ uint16_t cvt_16(float v, bool f16) {
  if (f16) { __half h = __float2half_rn(sat_f16(v)); return *reinterpret_cast<uint16_t*>(&h); }
  bf16 b = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
  __half2 v = *reinterpret_cast<__half2*>(&u);
  return __half22float2(v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

}  // namespace neko
