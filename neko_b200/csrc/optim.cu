// Optimiser step over one flat fp32 arena: global-norm clip (trainer.py:181-182, max_norm 1.0) and AdamW
// (train.py:127-133: betas 0.9/0.95, eps 1e-8, weight decay 0.1 on every parameter), fused in one pass.
// HBM-bound: 16 bytes read + 12 written per parameter.
#include "common.cuh"

namespace neko {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nv = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 v = __ldg(x4 + i);
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  for (long long i = (nv << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += x[i] * x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, const float* __restrict__ sumsq,
                                                    float max_norm, float grad_div, uint16_t* __restrict__ w_f16,
                                                    uint16_t* __restrict__ w_bf16, long long n_cast) {
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
  float coef = 1.0f / grad_div;
  if (sumsq != nullptr && max_norm > 0.f) {
    const float norm = sqrtf(__ldg(sumsq)) / grad_div;
    coef *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    float pi = p[i];
    pi *= (1.0f - lr * wd);                      // decoupled weight decay (torch.optim.AdamW)
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (i < n_cast) {   // refresh the 16-bit operand copies of the GEMM weights in the same pass
      if (w_f16) w_f16[i] = cvt_16(pi, true);
      if (w_bf16) w_bf16[i] = cvt_16(pi, false);
    }
  }
}

}  // namespace neko

extern "C" {

int neko_sumsq_f32(const float* x, int64_t n, float* out_accum, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(x && out_accum && n >= 0, "sumsq: bad arguments");
  if (n == 0) return NEKO_OK;
  NEKO_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sumsq: misaligned");
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n, out_accum);
  NEKO_LAUNCH_CHECK("sumsq_kernel");
  return NEKO_OK;
}

int neko_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, const float* grad_sumsq, float max_norm,
                    float grad_div, uint16_t* w_f16, uint16_t* w_bf16, int64_t n_cast, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1 && grad_div > 0.f, "adamw: bad arguments");
  if (n == 0) return NEKO_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                               bc1, bc2_sqrt, grad_sumsq, max_norm, grad_div, w_f16, w_bf16,
                                                               (w_f16 || w_bf16) ? (long long)n_cast : 0LL);
  NEKO_LAUNCH_CHECK("adamw_kernel");
  return NEKO_OK;
}

}  // extern "C"
