// Masked cross entropy over the joint text + continuous + discrete vocabulary (gato_policy.py:174-186).
//
// The reference shifts logits/targets by one position, boolean-selects the rows where
// token_masks[:, :-1] * token_target_masks[:, 1:] > 0 (a dynamic-shape gather that syncs the host) and
// takes F.cross_entropy(mean).  Here the host already knows the row list from the tokenisation plan
// (no sync); one CTA streams one selected logits row: online max / sum-exp in fp32, loss_i = lse - z[tgt].
// Backward writes (softmax - onehot) * g / n_rows as bf16 straight into the dlogits operand of the head
// dgrad / wgrad GEMMs.  HBM-bound: V*4 bytes read per selected row (+ V*2 written in backward).
#include <cuda_fp16.h>

#include "common.cuh"

namespace neko {

constexpr int CE_THREADS = 256;

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { m = mn; s = 0.f; return; }
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const float* __restrict__ logits, long long ld, int V,
                                                            const int32_t* __restrict__ rows, const int64_t* __restrict__ tokens,
                                                            float* __restrict__ row_lse, float* __restrict__ row_loss, int flags) {
  __shared__ float sm[CE_THREADS / 32], ss[CE_THREADS / 32];
  const int r = blockIdx.x;
  const long long pos = rows[r];
  const float* z = logits + ((flags & NEKO_CE_LOGITS_COMPACT) ? (long long)r : pos) * ld;
  float m = -INFINITY, s = 0.f;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  int start = 0;
  if (vec) {
    const int nv = V >> 2;
    const float4* z4 = reinterpret_cast<const float4*>(z);
    for (int i = threadIdx.x; i < nv; i += CE_THREADS) {
      const float4 v = __ldg(z4 + i);
      const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      const float mn = fmaxf(m, mx);
      s = s * __expf(m - mn) + __expf(v.x - mn) + __expf(v.y - mn) + __expf(v.z - mn) + __expf(v.w - mn);
      m = mn;
    }
    start = nv << 2;
  }
  for (int i = start + threadIdx.x; i < V; i += CE_THREADS) {
    const float v = z[i];
    const float mn = fmaxf(m, v);
    s = s * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CE_THREADS / 32; ++w) online_merge(m, s, sm[w], ss[w]);
    const float lse = m + logf(s);
    const long long tgt = tokens[pos + 1];
    row_lse[r] = lse;
    row_loss[r] = (tgt >= 0 && tgt < V) ? lse - z[tgt] : 0.f;
  }
}

// deterministic mean of row_loss
__global__ void __launch_bounds__(1024) ce_mean_kernel(const float* __restrict__ row_loss, int n, float* __restrict__ loss) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)row_loss[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    *loss = (float)(t / (double)n);
  }
}

__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const float* __restrict__ logits, long long ld, int V,
                                                            const int32_t* __restrict__ rows, int n_rows,
                                                            const int64_t* __restrict__ tokens, const float* __restrict__ row_lse,
                                                            const float* __restrict__ gscale, bf16* __restrict__ dlogits, long long ldd,
                                                            int flags) {
  const int r = blockIdx.x;
  const long long pos = rows[r];
  const float* z = logits + ((flags & NEKO_CE_LOGITS_COMPACT) ? (long long)r : pos) * ld;
  bf16* dz = dlogits + ((flags & NEKO_CE_DLOGITS_COMPACT) ? (long long)r : pos) * ldd;
  const float lse = row_lse[r];
  const float g = __ldg(gscale) / (float)n_rows;
  const int tgt = (int)tokens[pos + 1];
  const bool vec = ((ld & 3) == 0) && ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dlogits) & 7) == 0);
  int start = 0;
  if (vec) {
    const int nv = V >> 2;
    const float4* z4 = reinterpret_cast<const float4*>(z);
    uint2* d2 = reinterpret_cast<uint2*>(dz);
    for (int i = threadIdx.x; i < nv; i += CE_THREADS) {
      const float4 v = __ldg(z4 + i);
      const int c = i << 2;
      const float p0 = (__expf(v.x - lse) - (c == tgt ? 1.f : 0.f)) * g;
      const float p1 = (__expf(v.y - lse) - (c + 1 == tgt ? 1.f : 0.f)) * g;
      const float p2 = (__expf(v.z - lse) - (c + 2 == tgt ? 1.f : 0.f)) * g;
      const float p3 = (__expf(v.w - lse) - (c + 3 == tgt ? 1.f : 0.f)) * g;
      d2[i] = make_uint2(pack_bf16x2(p0, p1), pack_bf16x2(p2, p3));
    }
    start = nv << 2;
  }
  for (int i = start + threadIdx.x; i < V; i += CE_THREADS)
    dz[i] = __float2bfloat16_rn((__expf(z[i] - lse) - (i == tgt ? 1.f : 0.f)) * g);
  if (flags & NEKO_CE_ZERO_PAD)  // pad columns V..ld feed the K tail of the head dgrad GEMM
    for (long long i = V + threadIdx.x; i < ldd; i += CE_THREADS) dz[i] = __float2bfloat16_rn(0.f);
}


// ---------------------------------------------------------------------------------------------
// Fused forward + gradient (training): one CTA keeps a whole logits row (V fp32 = 209 KB at V = 52 305) in shared
// memory, so the row is read from HBM ONCE: max -> exp / sum -> loss, then (softmax - onehot) / n_rows is written as the
// bf16 dlogits operand of the head backward GEMMs.  The separate kernels above read every selected row twice (forward,
// backward).  The upstream gradient of the loss is not known yet in forward: backward multiplies by it only when it is
// not exactly 1 (ce_scale_kernel: gradient accumulation divides the loss, plain loss.backward() does not).
// ---------------------------------------------------------------------------------------------
constexpr int CEF_THREADS = 1024;

__device__ __forceinline__ void cef_cp_async16(void* smem_dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cef_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cef_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float cef_block_reduce(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();   // protects `red` against the previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < CEF_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

__global__ void __launch_bounds__(CEF_THREADS, 1) ce_fused_kernel(const float* __restrict__ logits, long long ld, int V,
                                                                  const int32_t* __restrict__ rows, int n_rows,
                                                                  const int64_t* __restrict__ tokens, float* __restrict__ row_lse,
                                                                  float* __restrict__ row_loss, bf16* __restrict__ dlogits, long long ldd,
                                                                  int flags) {
  extern __shared__ __align__(16) float cef_row[];   // [V rounded up to 4]
  __shared__ float red[CEF_THREADS / 32];
  const int nv = V >> 2;
  const float inv_n = 1.0f / (float)n_rows;
  float4* s4 = reinterpret_cast<float4*>(cef_row);
  // Every thread owns the float4 slots tid, tid + 1024, ... of the row buffer through all passes, so the buffer needs no
  // block-wide hand-over: a slot is refilled with the NEXT row (cp.async, 16 B, L2 only) right after its gradient has been
  // written, and the next row's HBM read overlaps this row's gradient write instead of following it.
  auto row_src = [&](int r) { return logits + ((flags & NEKO_CE_LOGITS_COMPACT) ? (long long)r : (long long)rows[r]) * ld; };
  if (blockIdx.x < n_rows) {
    const float4* z4 = reinterpret_cast<const float4*>(row_src(blockIdx.x));
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) cef_cp_async16(s4 + i, z4 + i);
    cef_cp_async_commit();
  }
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const long long pos = rows[r];
    const float* z = row_src(r);
    bf16* dz = dlogits + ((flags & NEKO_CE_DLOGITS_COMPACT) ? (long long)r : pos) * ldd;
    const bool has_next = r + (int)gridDim.x < n_rows;
    const float4* zn4 = reinterpret_cast<const float4*>(has_next ? row_src(r + (int)gridDim.x) : z);
    for (int i = (nv << 2) + threadIdx.x; i < V; i += CEF_THREADS) cef_row[i] = z[i];
    cef_cp_async_wait_all();
    // pass 1 (shared memory): maximum
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) {
      const float4 v = s4[i];
      mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    for (int i = (nv << 2) + threadIdx.x; i < V; i += CEF_THREADS) mx = fmaxf(mx, cef_row[i]);
    mx = cef_block_reduce(mx, true, red);
    // pass 2 (shared memory): e = exp(z - max), sum
    float sum = 0.f;
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) {
      float4 v = s4[i];
      v.x = __expf(v.x - mx); v.y = __expf(v.y - mx); v.z = __expf(v.z - mx); v.w = __expf(v.w - mx);
      s4[i] = v;
      sum += (v.x + v.y) + (v.z + v.w);
    }
    for (int i = (nv << 2) + threadIdx.x; i < V; i += CEF_THREADS) { const float e = __expf(cef_row[i] - mx); cef_row[i] = e; sum += e; }
    sum = cef_block_reduce(sum, false, red);
    const int tgt = (int)tokens[pos + 1];
    if (threadIdx.x == 0) {
      const float lse = mx + logf(sum);
      row_lse[r] = lse;
      row_loss[r] = (tgt >= 0 && tgt < V) ? lse - z[tgt] : 0.f;
    }
    // pass 3 (shared memory -> HBM): (softmax - onehot) / n_rows as bf16; each consumed slot starts loading the next row
    const float sc = inv_n / sum;
    uint2* d2 = reinterpret_cast<uint2*>(dz);
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) {
      const float4 v = s4[i];
      const int c = i << 2;
      d2[i] = make_uint2(pack_bf16x2(v.x * sc - (c == tgt ? inv_n : 0.f), v.y * sc - (c + 1 == tgt ? inv_n : 0.f)),
                         pack_bf16x2(v.z * sc - (c + 2 == tgt ? inv_n : 0.f), v.w * sc - (c + 3 == tgt ? inv_n : 0.f)));
      if (has_next) cef_cp_async16(s4 + i, zn4 + i);   // after the value above has been consumed (same thread, same slot)
    }
    cef_cp_async_commit();
    for (int i = (nv << 2) + threadIdx.x; i < V; i += CEF_THREADS) dz[i] = __float2bfloat16_rn(cef_row[i] * sc - (i == tgt ? inv_n : 0.f));
    if (flags & NEKO_CE_ZERO_PAD)
      for (long long i = V + threadIdx.x; i < ldd; i += CEF_THREADS) dz[i] = __float2bfloat16_rn(0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// Same fused pass for 16-bit (fp16) logits: the training path that does not hand logits back to the caller
// (materialize_logits = False, what the trainer runs: trainer.py:178 discards them) lets the head GEMM write fp16, so the
// row costs V*2 bytes to write, V*2 to read here and V*2 to write as the gradient operand -- 6V instead of the 10V of the fp32
// route (4V written by the GEMM, 4V read, 2V written).  Statistics are fp32; the fp16 rounding of a logit (<= 2^-11 relative)
// moves the loss by ~1e-5 relative and the softmax by less than the bf16 rounding of the gradient itself.  The row buffer holds
// the logits only (105 KB): pass 3 recomputes exp(z - lse) instead of keeping fp32 exponentials.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void h8_to_f(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

__global__ void __launch_bounds__(CEF_THREADS, 1) ce_fused_f16_kernel(const __half* __restrict__ logits, long long ld, int V,
                                                                      const int32_t* __restrict__ rows, int n_rows,
                                                                      const int64_t* __restrict__ tokens, float* __restrict__ row_lse,
                                                                      float* __restrict__ row_loss, bf16* __restrict__ dlogits, long long ldd,
                                                                      int flags) {
  extern __shared__ __align__(16) uint4 cef_row_h[];   // [V rounded up to 8] halves
  __shared__ float red[CEF_THREADS / 32];
  const int nv = V >> 3;                                // full 8-element (16-byte) slots
  const float inv_n = 1.0f / (float)n_rows;
  __half* row_h = reinterpret_cast<__half*>(cef_row_h);
  auto row_src = [&](int r) { return logits + ((flags & NEKO_CE_LOGITS_COMPACT) ? (long long)r : (long long)rows[r]) * ld; };
  if (blockIdx.x < n_rows) {
    const uint4* z8 = reinterpret_cast<const uint4*>(row_src(blockIdx.x));
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) cef_cp_async16(cef_row_h + i, z8 + i);
    cef_cp_async_commit();
  }
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const long long pos = rows[r];
    const __half* z = row_src(r);
    bf16* dz = dlogits + ((flags & NEKO_CE_DLOGITS_COMPACT) ? (long long)r : pos) * ldd;
    const bool has_next = r + (int)gridDim.x < n_rows;
    const uint4* zn8 = reinterpret_cast<const uint4*>(has_next ? row_src(r + (int)gridDim.x) : z);
    for (int i = (nv << 3) + threadIdx.x; i < V; i += CEF_THREADS) row_h[i] = z[i];
    cef_cp_async_wait_all();
    float f[8];
    // pass 1: maximum
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) {
      h8_to_f(cef_row_h[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
    }
    for (int i = (nv << 3) + threadIdx.x; i < V; i += CEF_THREADS) mx = fmaxf(mx, __half2float(row_h[i]));
    mx = cef_block_reduce(mx, true, red);
    // pass 2: sum of exp(z - max)
    float sum = 0.f;
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) {
      h8_to_f(cef_row_h[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += __expf(f[j] - mx);
    }
    for (int i = (nv << 3) + threadIdx.x; i < V; i += CEF_THREADS) sum += __expf(__half2float(row_h[i]) - mx);
    sum = cef_block_reduce(sum, false, red);
    const int tgt = (int)tokens[pos + 1];
    const float lse = mx + logf(sum);
    if (threadIdx.x == 0) {
      row_lse[r] = lse;
      row_loss[r] = (tgt >= 0 && tgt < V) ? lse - __half2float(z[tgt]) : 0.f;
    }
    // pass 3: (softmax - onehot) / n_rows as bf16; each consumed slot starts loading the next row
    uint4* d8 = reinterpret_cast<uint4*>(dz);
    for (int i = threadIdx.x; i < nv; i += CEF_THREADS) {
      h8_to_f(cef_row_h[i], f);
      const int c = i << 3;
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = __expf(f[j] - lse) * inv_n - (c + j == tgt ? inv_n : 0.f);
      d8[i] = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
      if (has_next) cef_cp_async16(cef_row_h + i, zn8 + i);   // after the slot has been consumed (same thread, same slot)
    }
    cef_cp_async_commit();
    __syncthreads();   // the scalar tail below reads row_h entries other threads wrote at the top of this iteration
    for (int i = (nv << 3) + threadIdx.x; i < V; i += CEF_THREADS)
      dz[i] = __float2bfloat16_rn(__expf(__half2float(row_h[i]) - lse) * inv_n - (i == tgt ? inv_n : 0.f));
    if (flags & NEKO_CE_ZERO_PAD)
      for (long long i = V + threadIdx.x; i < ldd; i += CEF_THREADS) dz[i] = __float2bfloat16_rn(0.f);
    __syncthreads();   // ... and the next iteration overwrites them
  }
}

// dlogits *= g unless g == 1 (bf16 [n_rows, ld], all columns)
__global__ void __launch_bounds__(256) ce_scale_kernel(bf16* __restrict__ d, long long n8, const float* __restrict__ gscale) {
  const float g = __ldg(gscale);
  if (g == 1.0f) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 v = reinterpret_cast<uint4*>(d)[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(w[j]); w[j] = pack_bf16x2(f.x * g, f.y * g); }
    reinterpret_cast<uint4*>(d)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace neko

extern "C" {

int neko_masked_ce_fwd(const float* logits, int64_t ld_logits, int V, const int32_t* rows, int n_rows, const int64_t* tokens,
                       float* row_lse, float* row_loss, float* loss, int flags, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(logits && rows && tokens && row_lse && row_loss && loss, "masked_ce_fwd: null pointer");
  NEKO_REQUIRE(V > 0 && n_rows > 0 && ld_logits >= V, "masked_ce_fwd: bad sizes (V=%d n_rows=%d)", V, n_rows);
  ce_fwd_kernel<<<n_rows, CE_THREADS, 0, as_stream(stream)>>>(logits, ld_logits, V, rows, tokens, row_lse, row_loss, flags);
  NEKO_LAUNCH_CHECK("ce_fwd_kernel");
  ce_mean_kernel<<<1, 1024, 0, as_stream(stream)>>>(row_loss, n_rows, loss);
  NEKO_LAUNCH_CHECK("ce_mean_kernel");
  return NEKO_OK;
}

int neko_masked_ce_bwd(const float* logits, int64_t ld_logits, int V, const int32_t* rows, int n_rows, const int64_t* tokens,
                       const float* row_lse, const float* gscale, uint16_t* dlogits, int64_t ld_dlogits, int flags, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(logits && rows && tokens && row_lse && gscale && dlogits, "masked_ce_bwd: null pointer");
  NEKO_REQUIRE(V > 0 && n_rows > 0 && ld_logits >= V && ld_dlogits >= V, "masked_ce_bwd: bad sizes");
  const long long ldd = ld_dlogits;
  ce_bwd_kernel<<<n_rows, CE_THREADS, 0, as_stream(stream)>>>(logits, ld_logits, V, rows, n_rows, tokens, row_lse, gscale,
                                                             reinterpret_cast<bf16*>(dlogits), ldd, flags);
  NEKO_LAUNCH_CHECK("ce_bwd_kernel");
  return NEKO_OK;
}

int neko_masked_ce_fused(const float* logits, int64_t ld_logits, int V, const int32_t* rows, int n_rows, const int64_t* tokens,
                         float* row_lse, float* row_loss, float* loss, uint16_t* dlogits, int64_t ld_dlogits, int flags, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(logits && rows && tokens && row_lse && row_loss && loss && dlogits, "masked_ce_fused: null pointer");
  NEKO_REQUIRE(V > 0 && n_rows > 0 && ld_logits >= V && ld_dlogits >= V, "masked_ce_fused: bad sizes");
  const size_t smem = ((size_t)V + 3) / 4 * 16;
  // the row must fit the CTA's shared memory, and the vector paths need aligned pitches: otherwise report "not applicable"
  if (smem > 220 * 1024 || (ld_logits & 3) || (ld_dlogits & 3) || (reinterpret_cast<uintptr_t>(logits) & 15) || (reinterpret_cast<uintptr_t>(dlogits) & 7))
    return 1;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(ce_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(ce_fused)");
    attr = smem;
  }
  const int grid = n_rows < sm_count() ? n_rows : sm_count();
  ce_fused_kernel<<<grid, CEF_THREADS, smem, as_stream(stream)>>>(logits, ld_logits, V, rows, n_rows, tokens, row_lse, row_loss,
                                                                  reinterpret_cast<bf16*>(dlogits), ld_dlogits, flags);
  NEKO_LAUNCH_CHECK("ce_fused_kernel");
  ce_mean_kernel<<<1, 1024, 0, as_stream(stream)>>>(row_loss, n_rows, loss);
  NEKO_LAUNCH_CHECK("ce_mean_kernel");
  return NEKO_OK;
}

int neko_masked_ce_fused_f16(const uint16_t* logits_f16, int64_t ld_logits, int V, const int32_t* rows, int n_rows, const int64_t* tokens,
                             float* row_lse, float* row_loss, float* loss, uint16_t* dlogits, int64_t ld_dlogits, int flags, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(logits_f16 && rows && tokens && row_lse && row_loss && loss && dlogits, "masked_ce_fused_f16: null pointer");
  NEKO_REQUIRE(V > 0 && n_rows > 0 && ld_logits >= V && ld_dlogits >= V, "masked_ce_fused_f16: bad sizes");
  const size_t smem = ((size_t)V + 7) / 8 * 16;
  if (smem > 220 * 1024 || (ld_logits & 7) || (ld_dlogits & 7) || (reinterpret_cast<uintptr_t>(logits_f16) & 15) || (reinterpret_cast<uintptr_t>(dlogits) & 15))
    return 1;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(ce_fused_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(ce_fused_f16)");
    attr = smem;
  }
  const int grid = n_rows < sm_count() ? n_rows : sm_count();
  ce_fused_f16_kernel<<<grid, CEF_THREADS, smem, as_stream(stream)>>>(reinterpret_cast<const __half*>(logits_f16), ld_logits, V, rows, n_rows, tokens,
                                                                      row_lse, row_loss, reinterpret_cast<bf16*>(dlogits), ld_dlogits, flags);
  NEKO_LAUNCH_CHECK("ce_fused_f16_kernel");
  ce_mean_kernel<<<1, 1024, 0, as_stream(stream)>>>(row_loss, n_rows, loss);
  NEKO_LAUNCH_CHECK("ce_mean_kernel");
  return NEKO_OK;
}

int neko_ce_scale_grad(uint16_t* dlogits, int64_t n, const float* gscale, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(dlogits && gscale && n > 0 && n % 8 == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0, "ce_scale_grad: bad arguments");
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ce_scale_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<bf16*>(dlogits), n8, gscale);
  NEKO_LAUNCH_CHECK("ce_scale_kernel");
  return NEKO_OK;
}

}  // extern "C"
