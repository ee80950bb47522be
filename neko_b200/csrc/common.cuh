// Shared helpers for the neko_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

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

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------------
// Kernels launched through launch_pdl() may be scheduled while the previous kernel of the stream is still draining: they
// run their prologue (barrier init, TMEM allocation, descriptor prefetch, index math), then pdl_wait() blocks until the
// previous grid has completed and its writes are visible.  Every kernel calls pdl_launch_dependents() first thing, so its
// own successor can start the same way.  Both instructions are no-ops for a normally launched kernel.  NEKO_PDL=0 turns
// the launch attribute off (plain stream order).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
// ---- GELU (exact-erf form, nn.GELU default) on the same A&S 7.1.26 kernel, arranged for instruction count ----
// h(x) = Phi(-|x|) = 0.5 erfc(|x| / sqrt2) = t (a1' + t (a2' + ...)) e,  t = 1 / (1 + p |x| / sqrt2),  e = exp(-x^2 / 2)
// (coefficients pre-multiplied by 0.5).  13 instructions per GELU incl. one MUFU.RCP and one MUFU.EX2; the .ftz
// forms skip libdevice's denormal fix-ups (arguments are >= 1 resp. results below 2^-126 flush to the correct 0).
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_tail(float x, float& e) {
  const float t = rcp_ftz(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  const float k = x * 0.84932180028801904272f;   // sqrt(log2(e) / 2):  exp(-x^2/2) = 2^(-k^2)
  e = ex2_ftz(-k * k);
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  return p * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float h = gelu_tail(x, e);
  return fmaf(-fabsf(x), h, fmaxf(x, 0.0f));      // x >= 0: x (1 - h);  x < 0: x h
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float h = gelu_tail(x, e);
  const float cdf = x >= 0.0f ? 1.0f - h : h;
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
// value and derivative from one evaluation of the tail
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  float e;
  const float h = gelu_tail(x, e);
  y = fmaf(-fabsf(x), h, fmaxf(x, 0.0f));
  dy = fmaf(x * 0.39894228040143267794f, e, x >= 0.0f ? 1.0f - h : h);
}

// ---- GELU, tanh form (HF "gelu_new", what pretrained GPT-2 checkpoints use: gato_policy.py:79-95 with --pretrained_lm) ----
// 0.5 x (1 + tanh(u)) = x sigmoid(2u),  u = sqrt(2/pi) (x + 0.044715 x^3);  d/dx = s + x s (1 - s) 2 u'
__device__ __forceinline__ float gelu_tanh_sig(float x, float& x2) {
  x2 = x * x;
  const float u2 = x * fmaf(0.044715f * 1.5957691216057308f, x2, 1.5957691216057308f);   // 2u
  return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * u2));
}
__device__ __forceinline__ float gelu_tanh(float x) {
  float x2;
  return x * gelu_tanh_sig(x, x2);
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  float x2;
  const float sg = gelu_tanh_sig(x, x2);
  const float du2 = fmaf(3.0f * 0.044715f * 1.5957691216057308f, x2, 1.5957691216057308f);      // d(2u)/dx
  return fmaf(x * sg * (1.0f - sg), du2, sg);
}

template <bool TANH> __device__ __forceinline__ float gelu_grad_sel(float x) { return TANH ? gelu_tanh_grad(x) : gelu_erf_grad(x); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// fp16 conversions saturate at the largest finite half instead of overflowing to inf (one F2FP.SATFINITE)
__device__ __forceinline__ float sat_f16(float v) { return fminf(fmaxf(v, -65504.0f), 65504.0f); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
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
