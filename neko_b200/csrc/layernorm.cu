// LayerNorm forward / backward for the decoder (ln_1, ln_2, ln_f; trajectory_gpt2.py:323,353,779).
//
// Forward reads the fp32 residual stream and emits the bf16 GEMM operand plus per-row mean / rstd.
// Backward fuses the residual-gradient add: dx_resid += LN'(dy), optionally emitting the bf16 copy
// the next dgrad/wgrad GEMM consumes.  HBM-bound; one warp per row, 128-bit accesses, row kept in
// registers (d <= 2048).
#include "common.cuh"

namespace neko {

constexpr int LN_MAX_D = 2048;  // NV (float4 per lane) is a template parameter: 1,2,4,6,8,16

template <int LN_MAX_VEC>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, uint16_t* __restrict__ y,
                                                            uint16_t* __restrict__ y2, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int N, int d, float eps, int out_f16) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int nv = d >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x + (size_t)row * d);
  float4 v[LN_MAX_VEC];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    const int i = lane + k * 32;
    if (i < nv) {
      v[k] = x4[i];
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  const float mean = warp_sum(sum) / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    const int i = lane + k * 32;
    if (i < nv) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
      sq += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)d + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yv = reinterpret_cast<uint2*>(y + (size_t)row * d);
  uint2* yv2 = y2 ? reinterpret_cast<uint2*>(y2 + (size_t)row * d) : nullptr;
#pragma unroll
  for (int k = 0; k < LN_MAX_VEC; ++k) {
    const int i = lane + k * 32;
    if (i < nv) {
      const float4 g = __ldg(g4 + i), bb = __ldg(b4 + i);
      const float o0 = (v[k].x - mean) * rstd * g.x + bb.x, o1 = (v[k].y - mean) * rstd * g.y + bb.y;
      const float o2 = (v[k].z - mean) * rstd * g.z + bb.z, o3 = (v[k].w - mean) * rstd * g.w + bb.w;
      yv[i] = make_uint2(pack_16x2(o0, o1, out_f16 != 0), pack_16x2(o2, o3, out_f16 != 0));
      if (yv2) yv2[i] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
  }
}

// Each CTA owns `rows_per_cta` consecutive rows; warps stride over them.  The row lives in registers (x and dy
// only: xhat and gamma*dy are recomputed in the second pass to keep the register count, hence the occupancy of
// this pure streaming kernel, reasonable).  dgamma / dbeta partials go to shared-memory accumulators (RED.shared),
// then one global atomicAdd per column per CTA.
template <int LN_MAX_VEC>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ dx_resid,
                                                            bf16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int N, int d, int rows_per_cta) {
  extern __shared__ float smem[];  // [2][d] accumulators
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int nv = d >> 2;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = min(N, row0 + rows_per_cta);
  for (int row = row0 + wid; row < row1; row += nwarps) {
    const float4* x4 = reinterpret_cast<const float4*>(x + (size_t)row * d);
    const uint2* dy2 = reinterpret_cast<const uint2*>(dy + (size_t)row * d);
    const float mu = mean[row], rs = rstd[row];
    float4 xv[LN_MAX_VEC];
    uint2 dv[LN_MAX_VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int i = lane + k * 32;
      if (i < nv) {
        xv[k] = x4[i];
        dv[k] = dy2[i];
      }
    }
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int i = lane + k * 32;
      if (i < nv) {
        const float2 d01 = unpack_bf16x2(dv[k].x), d23 = unpack_bf16x2(dv[k].y);
        const float4 g = __ldg(g4 + i);
        const float g0 = d01.x * g.x, g1 = d01.y * g.y, g2 = d23.x * g.z, g3 = d23.y * g.w;
        s1 += (g0 + g1) + (g2 + g3);
        s2 += (g0 * (xv[k].x - mu) + g1 * (xv[k].y - mu)) + (g2 * (xv[k].z - mu) + g3 * (xv[k].w - mu));
      }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) * rs / (float)d;
    float4* r4 = reinterpret_cast<float4*>(dx_resid + (size_t)row * d);
    uint2* o2 = dx_bf16 ? reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * d) : nullptr;
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int i = lane + k * 32;
      if (i < nv) {
        const float2 d01 = unpack_bf16x2(dv[k].x), d23 = unpack_bf16x2(dv[k].y);
        const float4 g = __ldg(g4 + i);
        const float h0 = (xv[k].x - mu) * rs, h1 = (xv[k].y - mu) * rs, h2 = (xv[k].z - mu) * rs, h3 = (xv[k].w - mu) * rs;
        float4 r = r4[i];
        r.x += rs * (d01.x * g.x - s1 - h0 * s2);
        r.y += rs * (d01.y * g.y - s1 - h1 * s2);
        r.z += rs * (d23.x * g.z - s1 - h2 * s2);
        r.w += rs * (d23.y * g.w - s1 - h3 * s2);
        r4[i] = r;
        if (o2) o2[i] = make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w));
        float* sg = smem + 4 * i;
        float* sb = smem + d + 4 * i;
        atomicAdd(sg + 0, d01.x * h0); atomicAdd(sg + 1, d01.y * h1); atomicAdd(sg + 2, d23.x * h2); atomicAdd(sg + 3, d23.y * h3);
        atomicAdd(sb + 0, d01.x); atomicAdd(sb + 1, d01.y); atomicAdd(sb + 2, d23.x); atomicAdd(sb + 3, d23.y);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(dgamma + i, smem[i]);
    atomicAdd(dbeta + i, smem[d + i]);
  }
}

}  // namespace neko

extern "C" {

int neko_layernorm_fwd(const float* x, const float* gamma, const float* beta, uint16_t* y_bf16, uint16_t* y2_bf16, float* mean,
                       float* rstd, int N, int d, float eps, int out_f16, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(x && gamma && beta && y_bf16 && mean && rstd, "layernorm_fwd: null pointer");
  NEKO_REQUIRE(N > 0 && d > 0 && d % 4 == 0 && d <= LN_MAX_D, "layernorm_fwd: need d %% 4 == 0 and d <= %d (got %d)", LN_MAX_D, d);
  const int threads = 256;
  const long long blocks = ((long long)N * 32 + threads - 1) / threads;
#define NEKO_LN_FWD(NV) layernorm_fwd_kernel<NV><<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(x, gamma, beta, y_bf16, y2_bf16, mean, rstd, N, d, eps, out_f16)
  const int need = (d / 4 + 31) / 32;
  if (need <= 1) NEKO_LN_FWD(1); else if (need <= 2) NEKO_LN_FWD(2); else if (need <= 4) NEKO_LN_FWD(4);
  else if (need <= 6) NEKO_LN_FWD(6); else if (need <= 8) NEKO_LN_FWD(8); else NEKO_LN_FWD(16);
#undef NEKO_LN_FWD
  NEKO_LAUNCH_CHECK("layernorm_fwd_kernel");
  return NEKO_OK;
}

int neko_layernorm_bwd(const uint16_t* dy_bf16, const float* x, const float* gamma, const float* mean, const float* rstd,
                       float* dx_resid, uint16_t* dx_bf16, float* dgamma, float* dbeta, int N, int d, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(dy_bf16 && x && gamma && mean && rstd && dx_resid && dgamma && dbeta, "layernorm_bwd: null pointer");
  NEKO_REQUIRE(N > 0 && d > 0 && d % 4 == 0 && d <= LN_MAX_D, "layernorm_bwd: need d %% 4 == 0 and d <= %d (got %d)", LN_MAX_D, d);
  const int threads = 256;
  // one resident wave (2 CTAs per SM at ~100 registers), each CTA a contiguous row range: few CTAs keep the
  // final global atomics (2*d per CTA, all on the same 2*d addresses) cheap
  int ctas = sm_count() * 2;
  int rows_per_cta = (N + ctas - 1) / ctas;
  if (rows_per_cta < 8) rows_per_cta = 8;
  ctas = (N + rows_per_cta - 1) / rows_per_cta;
  const size_t smem = 2 * (size_t)d * sizeof(float);
#define NEKO_LN_BWD(NV) layernorm_bwd_kernel<NV><<<ctas, threads, smem, as_stream(stream)>>>(reinterpret_cast<const bf16*>(dy_bf16), x, gamma, mean, rstd, dx_resid, reinterpret_cast<bf16*>(dx_bf16), dgamma, dbeta, N, d, rows_per_cta)
  const int need = (d / 4 + 31) / 32;
  if (need <= 1) NEKO_LN_BWD(1); else if (need <= 2) NEKO_LN_BWD(2); else if (need <= 4) NEKO_LN_BWD(4);
  else if (need <= 6) NEKO_LN_BWD(6); else if (need <= 8) NEKO_LN_BWD(8); else NEKO_LN_BWD(16);
#undef NEKO_LN_BWD
  NEKO_LAUNCH_CHECK("layernorm_bwd_kernel");
  return NEKO_OK;
}

}  // extern "C"
