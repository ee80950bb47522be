// Image patch embedding: ImageEmbedding.forward / ResidualBlock_V2 (embeddings.py:28-61,111-131).
//
//   x = (img/255*2 - 1)/sqrt(p)  ->  patchify 16x16  ->  h = conv1(GELU(x)) 3->C, 3x3, per-patch zero pad
//   -> GroupNorm(groups) -> GELU -> conv2 C->3 -> x + .   -> flatten (c p1 p2) -> fp16 row of the projection GEMM
//
// One CTA processes one patch at a time (persistent loop over patches, weights resident in shared memory),
// the C x 256 intermediate never leaves shared memory (the unfused torch path writes / re-reads 131 KB per
// patch four times).  fp32 CUDA-core arithmetic; thread = (2x2 pixel quad, quarter of the channels) so each
// weight fetched from shared memory feeds 4 FMAs.  Backward recomputes conv1 + GroupNorm from the saved
// (mean, rstd), accumulates all weight gradients in registers / shared memory across the CTA's patches and
// flushes once with atomics.
#include "common.cuh"

namespace neko {

constexpr int PE_P = 16;            // patch edge
constexpr int PE_PX = 256;          // pixels per patch
constexpr int PE_C = 128;           // mid channels (train.py always passes 128)
constexpr int PE_THREADS = 256;
constexpr int PE_HP = 257;          // hbuf pitch (odd: channel-owner sweeps are bank-conflict free)
constexpr int PE_PAD = 18;          // padded edge

struct PatchArgs {
  const void* images;
  int is_u8, n_img, Himg, Wimg, n_h, n_w, groups;
  const float *w1, *b1, *gw, *gb, *w2, *b2;
  uint16_t* out;        // fwd: [P,768] fp16
  uint16_t* out_bf;     // fwd: optional bf16 copy (wgrad operand)
  float* stats;         // [P, groups, 2]
  // backward only
  const bf16* dout;     // [P,768]
  float *dw1, *db1, *dgw, *dgb, *dw2, *db2;
};

// layout helpers: pixel (y,x) <-> plane index  sub*64 + quad   (sub = (y&1)*2 + (x&1), quad = (y>>1)*8 + (x>>1))
__device__ __forceinline__ int plane_index(int y, int x) { return (((y & 1) << 1) | (x & 1)) * 64 + ((y >> 1) << 3) + (x >> 1); }

struct PatchSmem {
  float* hbuf;    // [C][HP]
  float* w1s;     // [C][27]
  float* w2s;     // [3][C][9]
  float* b1s;     // [C]
  float* gws;     // [C]
  float* gbs;     // [C]
  float* gx;      // [3][18][18] gelu(x), zero padded
  float* xin;     // [3][256]    normalised x (standard pixel order)
  float* red;     // [4][3][256] conv2 partial sums / scratch
  float* gstat;   // [groups][2] sum, sumsq  -> mean, rstd
  float* dypad;   // [3][18][18] (backward)
  float* gsum;    // [groups][2] s1, s2 (backward)
  float* dgacc;   // [C] dgamma accumulators (backward)
  float* dbacc;   // [C]
  float* db2acc;  // [4]
};

__device__ __forceinline__ PatchSmem carve(float* base, bool bwd) {
  PatchSmem s;
  float* p = base;
  s.hbuf = p; p += PE_C * PE_HP;
  s.w1s = p; p += PE_C * 27;
  s.w2s = p; p += 3 * PE_C * 9;
  s.b1s = p; p += PE_C;
  s.gws = p; p += PE_C;
  s.gbs = p; p += PE_C;
  s.gx = p; p += 3 * PE_PAD * PE_PAD;
  s.xin = p; p += 3 * PE_PX;
  s.red = p; p += 4 * 3 * PE_PX;
  s.gstat = p; p += 2 * 64;
  s.dypad = p; p += bwd ? 3 * PE_PAD * PE_PAD : 0;
  s.gsum = p; p += bwd ? 2 * 64 : 0;
  s.dgacc = p; p += bwd ? PE_C : 0;
  s.dbacc = p; p += bwd ? PE_C : 0;
  s.db2acc = p;
  return s;
}
static size_t patch_smem_bytes(bool bwd) {
  size_t n = (size_t)PE_C * PE_HP + PE_C * 27 + 3 * PE_C * 9 + 3 * PE_C + 3 * PE_PAD * PE_PAD + 3 * PE_PX + 4 * 3 * PE_PX + 128;
  if (bwd) n += 3 * PE_PAD * PE_PAD + 128 + 2 * PE_C + 4;
  return n * sizeof(float);
}

__device__ __forceinline__ void load_weights(const PatchArgs& a, PatchSmem& s) {
  for (int i = threadIdx.x; i < PE_C * 27; i += PE_THREADS) s.w1s[i] = a.w1[i];
  for (int i = threadIdx.x; i < 3 * PE_C * 9; i += PE_THREADS) s.w2s[i] = a.w2[i];
  for (int i = threadIdx.x; i < PE_C; i += PE_THREADS) {
    s.b1s[i] = a.b1[i];
    s.gws[i] = a.gw[i];
    s.gbs[i] = a.gb[i];
  }
}

// load + normalise one patch: xin (standard order) and gx = gelu(x) zero-padded
__device__ __forceinline__ void load_patch(const PatchArgs& a, PatchSmem& s, int patch) {
  const int npp = a.n_h * a.n_w;
  const int n = patch / npp, r = patch - n * npp;
  const int ph = r / a.n_w, pw = r - ph * a.n_w;
  for (int i = threadIdx.x; i < 3 * PE_PAD * PE_PAD; i += PE_THREADS) s.gx[i] = 0.f;
  __syncthreads();
  const int t = threadIdx.x, y = t >> 4, x = t & 15;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const size_t idx = (((size_t)n * 3 + c) * a.Himg + (size_t)ph * PE_P + y) * a.Wimg + (size_t)pw * PE_P + x;
    float v = a.is_u8 ? (float)reinterpret_cast<const uint8_t*>(a.images)[idx] : reinterpret_cast<const float*>(a.images)[idx];
    v = v / 255.0f * 2.0f - 1.0f;   // embeddings.py:40
    v = v / 4.0f;                   // / sqrt(patch_size), :41
    s.xin[c * PE_PX + t] = v;
    s.gx[(c * PE_PAD + y + 1) * PE_PAD + x + 1] = gelu_erf(v);
  }
  __syncthreads();
}

// conv1 for this thread's quad and channel slice -> hbuf (raw, pre-GroupNorm)
__device__ __forceinline__ void conv1_quad(PatchSmem& s, int quad, int slice) {
  const int qy = quad >> 3, qx = quad & 7;
  float win[3][4][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) win[c][r][k] = s.gx[(c * PE_PAD + 2 * qy + r) * PE_PAD + 2 * qx + k];
  for (int cc = 0; cc < PE_C / 4; ++cc) {
    const int c = slice * (PE_C / 4) + cc;
    const float* w = s.w1s + c * 27;
    const float b = s.b1s[c];
    float o00 = b, o01 = b, o10 = b, o11 = b;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float wv = w[(ci * 3 + ky) * 3 + kx];
          o00 = fmaf(wv, win[ci][ky][kx], o00);
          o01 = fmaf(wv, win[ci][ky][kx + 1], o01);
          o10 = fmaf(wv, win[ci][ky + 1][kx], o10);
          o11 = fmaf(wv, win[ci][ky + 1][kx + 1], o11);
        }
    float* h = s.hbuf + c * PE_HP + quad;
    h[0] = o00; h[64] = o01; h[128] = o10; h[192] = o11;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PE_THREADS, 1) patch_resblock_fwd_kernel(PatchArgs a) {
  extern __shared__ __align__(16) float pe_smem[];
  PatchSmem s = carve(pe_smem, false);
  const int tid = threadIdx.x, lane = tid & 31;
  const int quad = tid & 63, slice = tid >> 6;
  const int P = a.n_img * a.n_h * a.n_w;
  const int gs = PE_C / a.groups;  // channels per group
  load_weights(a, s);
  for (int patch = blockIdx.x; patch < P; patch += gridDim.x) {
    load_patch(a, s, patch);
    if (tid < 2 * a.groups) s.gstat[tid] = 0.f;
    conv1_quad(s, quad, slice);
    __syncthreads();
    // GroupNorm statistics (biased variance over gs channels x 256 pixels)
    for (int g0 = 0; g0 < (PE_C / 4) / gs; ++g0) {
      const int g = slice * ((PE_C / 4) / gs) + g0;
      float sum = 0.f, sq = 0.f;
      for (int k = 0; k < gs; ++k) {
        const float* h = s.hbuf + (g * gs + k) * PE_HP + quad;
#pragma unroll
        for (int u = 0; u < 4; ++u) { const float v = h[u * 64]; sum += v; sq += v * v; }
      }
      sum = warp_sum(sum); sq = warp_sum(sq);
      if (lane == 0) { atomicAdd(&s.gstat[2 * g], sum); atomicAdd(&s.gstat[2 * g + 1], sq); }
    }
    __syncthreads();
    if (tid < a.groups) {
      const float n = (float)(gs * PE_PX);
      const float mean = s.gstat[2 * tid] / n;
      const float var = fmaxf(s.gstat[2 * tid + 1] / n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      s.gstat[2 * tid] = mean; s.gstat[2 * tid + 1] = rstd;
      a.stats[((size_t)patch * a.groups + tid) * 2] = mean;
      a.stats[((size_t)patch * a.groups + tid) * 2 + 1] = rstd;
    }
    __syncthreads();
    // normalise + GELU in place (own elements)
    for (int cc = 0; cc < PE_C / 4; ++cc) {
      const int c = slice * (PE_C / 4) + cc, g = c / gs;
      const float mean = s.gstat[2 * g], rstd = s.gstat[2 * g + 1], gw = s.gws[c], gb = s.gbs[c];
      float* h = s.hbuf + c * PE_HP + quad;
#pragma unroll
      for (int u = 0; u < 4; ++u) h[u * 64] = gelu_erf((h[u * 64] - mean) * rstd * gw + gb);
    }
    __syncthreads();
    // conv2 partial sums over this slice's channels
    const int qy = quad >> 3, qx = quad & 7;
    int off[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y = 2 * qy + r - 1, x = 2 * qx + k - 1;
        off[r][k] = (y >= 0 && y < PE_P && x >= 0 && x < PE_P) ? plane_index(y, x) : -1;
      }
    float acc[3][4];
#pragma unroll
    for (int co = 0; co < 3; ++co) acc[co][0] = acc[co][1] = acc[co][2] = acc[co][3] = 0.f;
    for (int cc = 0; cc < PE_C / 4; ++cc) {
      const int c = slice * (PE_C / 4) + cc;
      const float* h = s.hbuf + c * PE_HP;
      float win[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) win[r][k] = off[r][k] >= 0 ? h[off[r][k]] : 0.f;
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        const float* w = s.w2s + (co * PE_C + c) * 9;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float wv = w[ky * 3 + kx];
            acc[co][0] = fmaf(wv, win[ky][kx], acc[co][0]);
            acc[co][1] = fmaf(wv, win[ky][kx + 1], acc[co][1]);
            acc[co][2] = fmaf(wv, win[ky + 1][kx], acc[co][2]);
            acc[co][3] = fmaf(wv, win[ky + 1][kx + 1], acc[co][3]);
          }
      }
    }
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
      for (int u = 0; u < 4; ++u) s.red[(slice * 3 + co) * PE_PX + u * 64 + quad] = acc[co][u];
    __syncthreads();
    // residual + bias, write the (c p1 p2) row
    {
      const int y = tid >> 4, x = tid & 15, pi = plane_index(y, x);
      uint16_t* o = a.out + (size_t)patch * (3 * PE_PX);
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        float v = s.xin[co * PE_PX + tid] + a.b2[co];
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) v += s.red[(sl * 3 + co) * PE_PX + pi];
        o[co * PE_PX + tid] = cvt_16(v, true);
        if (a.out_bf) a.out_bf[(size_t)patch * (3 * PE_PX) + co * PE_PX + tid] = cvt_16(v, false);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PE_THREADS, 1) patch_resblock_bwd_kernel(PatchArgs a) {
  extern __shared__ __align__(16) float pe_smem[];
  PatchSmem s = carve(pe_smem, true);
  const int tid = threadIdx.x, lane = tid & 31;
  const int quad = tid & 63, slice = tid >> 6;
  const int qy = quad >> 3, qx = quad & 7;
  const int P = a.n_img * a.n_h * a.n_w;
  const int gs = PE_C / a.groups;
  // channel-owner mapping for the weight gradients: channel oc, taps [n0, n1)
  const int oc = tid & (PE_C - 1), half = tid >> 7;
  const int n0 = half ? 14 : 0, n1 = half ? 27 : 14;
  float dw2acc[14], dw1acc[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) dw2acc[i] = dw1acc[i] = 0.f;
  float db1acc = 0.f;
  // shared-memory offsets of this thread's taps (entry 13 of the upper half duplicates tap 26 and is dropped at the flush)
  int doff[14], goff[14];
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    const int n = min(n0 + j, 26);
    const int c3 = n / 9, tap = n - c3 * 9, ky = tap / 3, kx = tap - ky * 3;
    doff[j] = (c3 * PE_PAD + 2 - ky) * PE_PAD + 2 - kx;   // dY[co][y-ky+1][x-kx+1] in the padded tile
    goff[j] = (c3 * PE_PAD + ky) * PE_PAD + kx;           // gelu(x)[ci][y+ky-1][x+kx-1] in the padded tile
  }
  load_weights(a, s);
  for (int i = tid; i < PE_C; i += PE_THREADS) { s.dgacc[i] = 0.f; s.dbacc[i] = 0.f; }
  if (tid < 4) s.db2acc[tid] = 0.f;

  for (int patch = blockIdx.x; patch < P; patch += gridDim.x) {
    load_patch(a, s, patch);
    for (int i = tid; i < 3 * PE_PAD * PE_PAD; i += PE_THREADS) s.dypad[i] = 0.f;
    if (tid < a.groups) {
      s.gstat[2 * tid] = a.stats[((size_t)patch * a.groups + tid) * 2];
      s.gstat[2 * tid + 1] = a.stats[((size_t)patch * a.groups + tid) * 2 + 1];
    }
    if (tid < 2 * a.groups) s.gsum[tid] = 0.f;
    __syncthreads();
    {
      const int y = tid >> 4, x = tid & 15;
      const bf16* g = a.dout + (size_t)patch * (3 * PE_PX);
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        const float v = __bfloat162float(g[co * PE_PX + tid]);
        s.dypad[(co * PE_PAD + y + 1) * PE_PAD + x + 1] = v;
        const float t = warp_sum(v);
        if (lane == 0) atomicAdd(&s.db2acc[co], t);
      }
    }
    conv1_quad(s, quad, slice);
    __syncthreads();

    // ---- dW2[co][oc][tap] += sum_px h2[oc][px] * dY[co][px - tap + 1] ----
    {
      const int g = oc / gs;
      const float mean = s.gstat[2 * g], rstd = s.gstat[2 * g + 1], gw = s.gws[oc], gb = s.gbs[oc];
      const float* h = s.hbuf + oc * PE_HP;
      for (int i = 0; i < PE_PX; ++i) {
        const int sub = i >> 6, q = i & 63;
        const int y = 2 * (q >> 3) + (sub >> 1), x = 2 * (q & 7) + (sub & 1);
        const float h2 = gelu_erf((h[i] - mean) * rstd * gw + gb);
        const int base = y * PE_PAD + x;
#pragma unroll
        for (int j = 0; j < 14; ++j) dw2acc[j] = fmaf(h2, s.dypad[doff[j] + base], dw2acc[j]);
      }
    }
    // ---- pass A: dh2 -> dhn, group sums, dgamma / dbeta ----
    float dyw[3][4][4];
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) dyw[co][r][k] = s.dypad[(co * PE_PAD + 2 * qy + r) * PE_PAD + 2 * qx + k];
    auto dh2_of = [&](int c, float (&d)[4]) {
      d[0] = d[1] = d[2] = d[3] = 0.f;
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        const float* w = s.w2s + (co * PE_C + c) * 9;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float wv = w[ky * 3 + kx];
            // pixel (y,x) receives w[ky][kx] * dY[y-ky+1][x-kx+1]; window origin is (2qy-1, 2qx-1)
            d[0] = fmaf(wv, dyw[co][2 - ky][2 - kx], d[0]);
            d[1] = fmaf(wv, dyw[co][2 - ky][3 - kx], d[1]);
            d[2] = fmaf(wv, dyw[co][3 - ky][2 - kx], d[2]);
            d[3] = fmaf(wv, dyw[co][3 - ky][3 - kx], d[3]);
          }
      }
    };
    for (int g0 = 0; g0 < (PE_C / 4) / gs; ++g0) {
      const int g = slice * ((PE_C / 4) / gs) + g0;
      const float mean = s.gstat[2 * g], rstd = s.gstat[2 * g + 1];
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < gs; ++k) {
        const int c = g * gs + k;
        const float gw = s.gws[c], gb = s.gbs[c];
        float d[4];
        dh2_of(c, d);
        const float* h = s.hbuf + c * PE_HP + quad;
        float dg = 0.f, db = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float xh = (h[u * 64] - mean) * rstd;
          const float dhn = d[u] * gelu_erf_grad(xh * gw + gb);
          dg += dhn * xh; db += dhn;
        }
        s1 += gw * db; s2 += gw * dg;
        dg = warp_sum(dg); db = warp_sum(db);
        if (lane == 0) { atomicAdd(&s.dgacc[c], dg); atomicAdd(&s.dbacc[c], db); }
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { atomicAdd(&s.gsum[2 * g], s1); atomicAdd(&s.gsum[2 * g + 1], s2); }
    }
    __syncthreads();
    // ---- pass B: dh (gradient at the conv1 output) written over hbuf ----
    {
      const float inv_n = 1.0f / (float)(gs * PE_PX);
      for (int cc = 0; cc < PE_C / 4; ++cc) {
        const int c = slice * (PE_C / 4) + cc, g = c / gs;
        const float mean = s.gstat[2 * g], rstd = s.gstat[2 * g + 1], gw = s.gws[c], gb = s.gbs[c];
        const float m1 = s.gsum[2 * g] * inv_n, m2 = s.gsum[2 * g + 1] * inv_n;
        float d[4];
        dh2_of(c, d);
        float* h = s.hbuf + c * PE_HP + quad;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float xh = (h[u * 64] - mean) * rstd;
          const float dhn = d[u] * gelu_erf_grad(xh * gw + gb);
          h[u * 64] = rstd * (gw * dhn - m1 - xh * m2);
        }
      }
    }
    __syncthreads();
    // ---- dW1[oc][ci][tap] += sum_px dh[oc][px] * gelu(x)[ci][px + tap - 1];  db1[oc] += sum_px dh ----
    {
      const float* h = s.hbuf + oc * PE_HP;
      for (int i = 0; i < PE_PX; ++i) {
        const int sub = i >> 6, q = i & 63;
        const int y = 2 * (q >> 3) + (sub >> 1), x = 2 * (q & 7) + (sub & 1);
        const float dh = h[i];
        if (half == 0) db1acc += dh;
        const int base = y * PE_PAD + x;
#pragma unroll
        for (int j = 0; j < 14; ++j) dw1acc[j] = fmaf(dh, s.gx[goff[j] + base], dw1acc[j]);
      }
    }
    __syncthreads();
  }
  // flush
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    const int n = n0 + j;
    if (n < n1) {
      const int co = n / 9, tap = n - co * 9;
      atomicAdd(a.dw2 + ((size_t)co * PE_C + oc) * 9 + tap, dw2acc[j]);
      atomicAdd(a.dw1 + (size_t)oc * 27 + n, dw1acc[j]);
    }
  }
  if (half == 0) atomicAdd(a.db1 + oc, db1acc);
  __syncthreads();
  for (int i = tid; i < PE_C; i += PE_THREADS) { atomicAdd(a.dgw + i, s.dgacc[i]); atomicAdd(a.dgb + i, s.dbacc[i]); }
  if (tid < 3) atomicAdd(a.db2 + tid, s.db2acc[tid]);
}

// ---------------------------------------------------------------------------------------------
// patch position encodings
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_pos_add_kernel(float* __restrict__ x, int P, int d, const int32_t* __restrict__ row_bin,
                                                            const int32_t* __restrict__ col_bin, const float* __restrict__ row_tab,
                                                            const float* __restrict__ col_tab) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= P) return;
  float4* x4 = reinterpret_cast<float4*>(x + (size_t)w * d);
  const float4* r4 = reinterpret_cast<const float4*>(row_tab + (size_t)row_bin[w] * d);
  const float4* c4 = reinterpret_cast<const float4*>(col_tab + (size_t)col_bin[w] * d);
  for (int i = lane; i < (d >> 2); i += 32) {
    float4 v = x4[i];
    const float4 r = __ldg(r4 + i), c = __ldg(c4 + i);
    v.x += r.x + c.x; v.y += r.y + c.y; v.z += r.z + c.z; v.w += r.w + c.w;
    x4[i] = v;
  }
}

__global__ void __launch_bounds__(256) patch_pos_bwd_kernel(const float* __restrict__ dx, int P, int d, const int32_t* __restrict__ row_bin,
                                                            const int32_t* __restrict__ col_bin, float* __restrict__ d_row_tab,
                                                            float* __restrict__ d_col_tab) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= P) return;
  const float* g = dx + (size_t)w * d;
  float* r = d_row_tab + (size_t)row_bin[w] * d;
  float* c = d_col_tab + (size_t)col_bin[w] * d;
  for (int i = lane; i < d; i += 32) {
    const float v = g[i];
    atomicAdd(r + i, v);
    atomicAdd(c + i, v);
  }
}

static int check_patch_args(int n_img, int Himg, int Wimg, int patch, int C, int groups) {
  NEKO_REQUIRE(patch == PE_P, "patch_resblock: only patch_size 16 is implemented (got %d)", patch);
  NEKO_REQUIRE(C == PE_C, "patch_resblock: only resid_mid_channels 128 is implemented (got %d)", C);
  NEKO_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0 && (C / 4) % (C / groups) == 0, "patch_resblock: unsupported num_groups %d", groups);
  NEKO_REQUIRE(n_img > 0 && Himg > 0 && Wimg > 0 && Himg % patch == 0 && Wimg % patch == 0, "Image dimensions must be divisible by patch size");
  return NEKO_OK;
}

}  // namespace neko

extern "C" {

int neko_patch_resblock_fwd(const void* images, int is_u8, int n_img, int Himg, int Wimg, int patch, int C, int groups,
                            const float* conv1_w, const float* conv1_b, const float* gn_w, const float* gn_b, const float* conv2_w,
                            const float* conv2_b, uint16_t* patches_out, uint16_t* patches_out_bf16, float* gn_stats, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(images && conv1_w && conv1_b && gn_w && gn_b && conv2_w && conv2_b && patches_out && gn_stats, "patch_resblock_fwd: null pointer");
  int rc = check_patch_args(n_img, Himg, Wimg, patch, C, groups);
  if (rc != NEKO_OK) return rc;
  PatchArgs a{};
  a.images = images; a.is_u8 = is_u8; a.n_img = n_img; a.Himg = Himg; a.Wimg = Wimg; a.n_h = Himg / patch; a.n_w = Wimg / patch;
  a.groups = groups; a.w1 = conv1_w; a.b1 = conv1_b; a.gw = gn_w; a.gb = gn_b; a.w2 = conv2_w; a.b2 = conv2_b;
  a.out = patches_out; a.out_bf = patches_out_bf16; a.stats = gn_stats;
  const size_t smem = patch_smem_bytes(false);
  cudaError_t e = cudaFuncSetAttribute(patch_resblock_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(patch_fwd)");
  const int P = n_img * a.n_h * a.n_w;
  const int grid = P < sm_count() ? P : sm_count();
  patch_resblock_fwd_kernel<<<grid, PE_THREADS, smem, as_stream(stream)>>>(a);
  NEKO_LAUNCH_CHECK("patch_resblock_fwd_kernel");
  return NEKO_OK;
}

int neko_patch_resblock_bwd(const void* images, int is_u8, int n_img, int Himg, int Wimg, int patch, int C, int groups,
                            const float* conv1_w, const float* conv1_b, const float* gn_w, const float* gn_b, const float* conv2_w,
                            const float* gn_stats, const uint16_t* d_patches_bf16, float* d_conv1_w, float* d_conv1_b, float* d_gn_w,
                            float* d_gn_b, float* d_conv2_w, float* d_conv2_b, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(images && conv1_w && conv1_b && gn_w && gn_b && conv2_w && gn_stats && d_patches_bf16 && d_conv1_w && d_conv1_b && d_gn_w &&
               d_gn_b && d_conv2_w && d_conv2_b, "patch_resblock_bwd: null pointer");
  int rc = check_patch_args(n_img, Himg, Wimg, patch, C, groups);
  if (rc != NEKO_OK) return rc;
  PatchArgs a{};
  a.images = images; a.is_u8 = is_u8; a.n_img = n_img; a.Himg = Himg; a.Wimg = Wimg; a.n_h = Himg / patch; a.n_w = Wimg / patch;
  a.groups = groups; a.w1 = conv1_w; a.b1 = conv1_b; a.gw = gn_w; a.gb = gn_b; a.w2 = conv2_w; a.b2 = nullptr;
  a.stats = const_cast<float*>(gn_stats); a.dout = reinterpret_cast<const bf16*>(d_patches_bf16);
  a.dw1 = d_conv1_w; a.db1 = d_conv1_b; a.dgw = d_gn_w; a.dgb = d_gn_b; a.dw2 = d_conv2_w; a.db2 = d_conv2_b;
  const size_t smem = patch_smem_bytes(true);
  cudaError_t e = cudaFuncSetAttribute(patch_resblock_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(patch_bwd)");
  const int P = n_img * a.n_h * a.n_w;
  const int grid = P < sm_count() ? P : sm_count();
  patch_resblock_bwd_kernel<<<grid, PE_THREADS, smem, as_stream(stream)>>>(a);
  NEKO_LAUNCH_CHECK("patch_resblock_bwd_kernel");
  return NEKO_OK;
}

int neko_patch_pos_add(float* x, int P, int d, const int32_t* row_bin, const int32_t* col_bin, const float* row_tab,
                       const float* col_tab, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(x && row_bin && col_bin && row_tab && col_tab && P > 0 && d > 0 && d % 4 == 0, "patch_pos_add: bad arguments");
  patch_pos_add_kernel<<<(unsigned)(((long long)P * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(x, P, d, row_bin, col_bin, row_tab, col_tab);
  NEKO_LAUNCH_CHECK("patch_pos_add_kernel");
  return NEKO_OK;
}

int neko_patch_pos_bwd(const float* dx, int P, int d, const int32_t* row_bin, const int32_t* col_bin, float* d_row_tab,
                       float* d_col_tab, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(dx && row_bin && col_bin && d_row_tab && d_col_tab && P > 0 && d > 0, "patch_pos_bwd: bad arguments");
  patch_pos_bwd_kernel<<<(unsigned)(((long long)P * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(dx, P, d, row_bin, col_bin, d_row_tab, d_col_tab);
  NEKO_LAUNCH_CHECK("patch_pos_bwd_kernel");
  return NEKO_OK;
}

}  // extern "C"
