// Causal self-attention with left padding, forward and backward (flash style: the S x S score matrix of
// Attention._attn, trajectory_gpt2.py:163-188, is never materialised).
//
// Reference semantics: w = q k^T / sqrt(dh); w = where(tril, w, -1e4); w += (1 - token_mask) * -1e4;
// softmax; w v.  For a valid query row the -1e4 terms underflow to exactly 0 probability in fp32, so the
// row attends keys in [first_valid[b], query].  Padded query rows (left pad, or the right pad of --pad_seq)
// are never read by the loss (gato_policy.py:177-180) nor by valid rows; they are written as zeros.
//
// Tensor-core path: warp-level mma.sync m16n8k16 bf16 with fp32 accumulation, operands staged in shared
// memory so that every B fragment is a k-contiguous 32-bit load.  dh = 32 (24 heads x 32 at d=768) makes
// this op softmax/exp-bound rather than MMA-bound; it is <3% of the step FLOPs.
#include "common.cuh"

namespace neko {

constexpr int ATT_BLK = 64;       // queries per CTA = keys per tile
constexpr int ATT_THREADS = 128;  // 4 warps x 16 rows

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Shared-memory tile of `rows` x `cols` bf16 with an 8-element row pad (keeps 32-bit fragment loads
// conflict-free: the pitch in words is = 4 mod 8... cols/2 + 4).
template <int COLS>
struct Tile {
  static constexpr int PITCH = COLS + 8;
  bf16* p;
  __device__ __forceinline__ uint32_t ld32(int r, int c) const { return *reinterpret_cast<const uint32_t*>(p + r * PITCH + c); }
  __device__ __forceinline__ bf16& at(int r, int c) { return p[r * PITCH + c]; }
};

// A fragment (16 x 16 at column k0) of a row-major smem tile whose rows are the M index.
template <int COLS>
__device__ __forceinline__ void load_a_frag(const Tile<COLS>& t, int row0, int k0, int lane, uint32_t (&a)[4]) {
  const int g = lane >> 2, q = lane & 3;
  a[0] = t.ld32(row0 + g, k0 + 2 * q);
  a[1] = t.ld32(row0 + g + 8, k0 + 2 * q);
  a[2] = t.ld32(row0 + g, k0 + 8 + 2 * q);
  a[3] = t.ld32(row0 + g + 8, k0 + 8 + 2 * q);
}

// cooperative loads: rows [r0, r0+64) x DH of one head from the packed qkv / out / dout tensors
template <int DH>
__device__ __forceinline__ void load_tile(Tile<DH> dst, const bf16* __restrict__ src, long long row_pitch, int r0, int r_end) {
  constexpr int VPR = DH / 8;  // uint4 per row
  for (int i = threadIdx.x; i < ATT_BLK * VPR; i += ATT_THREADS) {
    const int r = i / VPR, v = i % VPR;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r0 + r < r_end) val = __ldg(reinterpret_cast<const uint4*>(src + (long long)(r0 + r) * row_pitch) + v);
    *reinterpret_cast<uint4*>(dst.p + r * Tile<DH>::PITCH + v * 8) = val;
  }
}
// transposed: dst[c][r] = src[r0 + r][c]
template <int DH>
__device__ __forceinline__ void load_tile_t(Tile<ATT_BLK> dst, const bf16* __restrict__ src, long long row_pitch, int r0, int r_end) {
  constexpr int VPR = DH / 8;
  for (int i = threadIdx.x; i < ATT_BLK * VPR; i += ATT_THREADS) {
    const int r = i % ATT_BLK, v = i / ATT_BLK;  // consecutive threads -> consecutive rows: conflict-free smem writes
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r0 + r < r_end) val = __ldg(reinterpret_cast<const uint4*>(src + (long long)(r0 + r) * row_pitch) + v);
    const bf16* e = reinterpret_cast<const bf16*>(&val);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst.at(v * 8 + j, r) = e[j];
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ first_valid,
                                                               bf16* __restrict__ out, bf16* __restrict__ out2, float* __restrict__ lse, int S,
                                                               int S_valid, int H, float scale_log2, int out_f16) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile<DH> sQ{reinterpret_cast<bf16*>(smem_raw)};
  Tile<DH> sK{sQ.p + ATT_BLK * Tile<DH>::PITCH};
  Tile<ATT_BLK> sVt{sK.p + ATT_BLK * Tile<DH>::PITCH};  // [DH][64 keys]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_BLK;
  const int d = H * DH;
  const long long pitch = 3LL * d;
  const bf16* base = qkv + (long long)b * S * pitch;
  const int lo = first_valid[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;

  const int q_hi = min(q0 + ATT_BLK, S_valid);  // queries in [max(q0,lo), q_hi) are live
  float o[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

  if (q_hi > lo && q_hi > q0) {
    load_tile<DH>(sQ, base + h * DH, pitch, q0, S);
    __syncthreads();
    uint32_t qa[DH / 16][4];
#pragma unroll
    for (int k = 0; k < DH / 16; ++k) load_a_frag<DH>(sQ, warp * 16, k * 16, lane, qa[k]);

    const int j_begin = (lo / ATT_BLK) * ATT_BLK;
    for (int j0 = j_begin; j0 < q_hi; j0 += ATT_BLK) {
      __syncthreads();
      load_tile<DH>(sK, base + d + h * DH, pitch, j0, S);
      load_tile_t<DH>(sVt, base + 2 * d + h * DH, pitch, j0, S);
      __syncthreads();
      float s[ATT_BLK / 8][4];
#pragma unroll
      for (int n = 0; n < ATT_BLK / 8; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          mma_bf16(s[n], qa[k], sK.ld32(n * 8 + g, k * 16 + 2 * q), sK.ld32(n * 8 + g, k * 16 + 8 + 2 * q));
      }
      // mask + online softmax (rows row_a, row_b; this thread holds keys j0 + n*8 + 2q, +1)
      float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
      for (int n = 0; n < ATT_BLK / 8; ++n) {
        const int key = j0 + n * 8 + 2 * q;
        if (key < lo || key > row_a || row_a >= S_valid) s[n][0] = -INFINITY;
        if (key + 1 < lo || key + 1 > row_a || row_a >= S_valid) s[n][1] = -INFINITY;
        if (key < lo || key > row_b || row_b >= S_valid) s[n][2] = -INFINITY;
        if (key + 1 < lo || key + 1 > row_b || row_b >= S_valid) s[n][3] = -INFINITY;
        mx_a = fmaxf(mx_a, fmaxf(s[n][0], s[n][1]));
        mx_b = fmaxf(mx_b, fmaxf(s[n][2], s[n][3]));
      }
      mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
      mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
      mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
      mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
      const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
      const float ref_a = (mn_a == -INFINITY) ? 0.f : mn_a * scale_log2;
      const float ref_b = (mn_b == -INFINITY) ? 0.f : mn_b * scale_log2;
      const float corr_a = exp2f(m_a * scale_log2 - ref_a), corr_b = exp2f(m_b * scale_log2 - ref_b);
      m_a = mn_a; m_b = mn_b;
      float sum_a = 0.f, sum_b = 0.f;
#pragma unroll
      for (int n = 0; n < ATT_BLK / 8; ++n) {
        s[n][0] = exp2f(s[n][0] * scale_log2 - ref_a);
        s[n][1] = exp2f(s[n][1] * scale_log2 - ref_a);
        s[n][2] = exp2f(s[n][2] * scale_log2 - ref_b);
        s[n][3] = exp2f(s[n][3] * scale_log2 - ref_b);
        sum_a += s[n][0] + s[n][1];
        sum_b += s[n][2] + s[n][3];
      }
      l_a = l_a * corr_a + sum_a;
      l_b = l_b * corr_b + sum_b;
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) {
        o[n][0] *= corr_a; o[n][1] *= corr_a; o[n][2] *= corr_b; o[n][3] *= corr_b;
      }
      // O += P V  (P from the score accumulators, V^T tile gives k-contiguous B fragments)
#pragma unroll
      for (int kk = 0; kk < ATT_BLK / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          mma_bf16(o[n], pa, sVt.ld32(n * 8 + g, kk * 16 + 2 * q), sVt.ld32(n * 8 + g, kk * 16 + 8 + 2 * q));
      }
    }
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
  }
  // write-out: dead rows (padding) -> zeros, lse = +inf so that backward sees p = 0
  const float inv_a = (l_a > 0.f) ? 1.f / l_a : 0.f, inv_b = (l_b > 0.f) ? 1.f / l_b : 0.f;
  const float kLn2 = 0.6931471805599453f;
  bf16* ob = out + (long long)b * S * d + h * DH;
  bf16* ob2 = out2 ? out2 + (long long)b * S * d + h * DH : nullptr;
  if (row_a < S) {
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(ob + (long long)row_a * d + n * 8 + 2 * q) = pack_16x2(o[n][0] * inv_a, o[n][1] * inv_a, out_f16 != 0);
      if (ob2) *reinterpret_cast<uint32_t*>(ob2 + (long long)row_a * d + n * 8 + 2 * q) = pack_bf16x2(o[n][0] * inv_a, o[n][1] * inv_a);
    }
    if (q == 0) lse[((long long)b * H + h) * S + row_a] = (l_a > 0.f) ? (m_a * scale_log2 + log2f(l_a)) * kLn2 : INFINITY;
  }
  if (row_b < S) {
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(ob + (long long)row_b * d + n * 8 + 2 * q) = pack_16x2(o[n][2] * inv_b, o[n][3] * inv_b, out_f16 != 0);
      if (ob2) *reinterpret_cast<uint32_t*>(ob2 + (long long)row_b * d + n * 8 + 2 * q) = pack_bf16x2(o[n][2] * inv_b, o[n][3] * inv_b);
    }
    if (q == 0) lse[((long long)b * H + h) * S + row_b] = (l_b > 0.f) ? (m_b * scale_log2 + log2f(l_b)) * kLn2 : INFINITY;
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_c dO[i,c] * O[i,c]
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta,
                                                         int B, int S, int H, int dh, int out_f16) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * S * H;
  if (w >= total) return;
  const int h = (int)(w % H);
  const long long bs = w / H;  // b*S + s
  const bf16* o = out + bs * (long long)H * dh + (long long)h * dh;
  const bf16* g = dout + bs * (long long)H * dh + (long long)h * dh;
  float acc = 0.f;
  for (int c = lane * 2; c < dh; c += 64) {
    const uint32_t ou = *reinterpret_cast<const uint32_t*>(o + c);
    const float2 a = out_f16 ? unpack_f16x2(ou) : unpack_bf16x2(ou);
    const float2 bb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(g + c));
    acc += a.x * bb.x + a.y * bb.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const int b = (int)(bs / S), s = (int)(bs % S);
    delta[((long long)b * H + h) * S + s] = acc;
  }
}

// dK, dV: one CTA owns 64 keys (4 warps x 16) of one head and sweeps the query tiles at or below it.
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                   const float* __restrict__ lse, const float* __restrict__ delta,
                                                                   const int32_t* __restrict__ first_valid, bf16* __restrict__ dqkv,
                                                                   int S, int S_valid, int H, float scale, float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile<DH> sK{reinterpret_cast<bf16*>(smem_raw)};
  Tile<DH> sV{sK.p + ATT_BLK * Tile<DH>::PITCH};
  Tile<DH> sQ{sV.p + ATT_BLK * Tile<DH>::PITCH};
  Tile<DH> sdO{sQ.p + ATT_BLK * Tile<DH>::PITCH};
  Tile<ATT_BLK> sQt{sdO.p + ATT_BLK * Tile<DH>::PITCH};            // [DH][64 queries]
  Tile<ATT_BLK> sdOt{sQt.p + DH * Tile<ATT_BLK>::PITCH};           // [DH][64 queries]
  float* s_lse = reinterpret_cast<float*>(sdOt.p + DH * Tile<ATT_BLK>::PITCH);
  float* s_delta = s_lse + ATT_BLK;

  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * ATT_BLK;
  const int d = H * DH;
  const long long pitch = 3LL * d;
  const bf16* base = qkv + (long long)b * S * pitch;
  const bf16* dob = dout + (long long)b * S * d + h * DH;
  const float* lse_b = lse + ((long long)b * H + h) * S;
  const float* delta_b = delta + ((long long)b * H + h) * S;
  const int lo = first_valid[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int key_a = j0 + warp * 16 + g, key_b = key_a + 8;

  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  const bool live = (j0 + ATT_BLK > lo) && (j0 < S_valid);
  if (live) {
    load_tile<DH>(sK, base + d + h * DH, pitch, j0, S);
    load_tile<DH>(sV, base + 2 * d + h * DH, pitch, j0, S);
    for (int i0 = j0; i0 < S_valid; i0 += ATT_BLK) {
      __syncthreads();
      load_tile<DH>(sQ, base + h * DH, pitch, i0, S);
      load_tile<DH>(sdO, dob, d, i0, S);
      load_tile_t<DH>(sQt, base + h * DH, pitch, i0, S);
      load_tile_t<DH>(sdOt, dob, d, i0, S);
      if (threadIdx.x < ATT_BLK) {
        const int i = i0 + threadIdx.x;
        s_lse[threadIdx.x] = (i < S) ? lse_b[i] : INFINITY;
        s_delta[threadIdx.x] = (i < S) ? delta_b[i] : 0.f;
      }
      __syncthreads();
      // S^T = K Q^T and dP^T = V dO^T   (rows = this warp's 16 keys, cols = 64 queries)
      float st[ATT_BLK / 8][4], dpt[ATT_BLK / 8][4];
#pragma unroll
      for (int n = 0; n < ATT_BLK / 8; ++n) {
        st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
        dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        uint32_t ka[4], va[4];
        load_a_frag<DH>(sK, warp * 16, k * 16, lane, ka);
        load_a_frag<DH>(sV, warp * 16, k * 16, lane, va);
#pragma unroll
        for (int n = 0; n < ATT_BLK / 8; ++n) {
          mma_bf16(st[n], ka, sQ.ld32(n * 8 + g, k * 16 + 2 * q), sQ.ld32(n * 8 + g, k * 16 + 8 + 2 * q));
          mma_bf16(dpt[n], va, sdO.ld32(n * 8 + g, k * 16 + 2 * q), sdO.ld32(n * 8 + g, k * 16 + 8 + 2 * q));
        }
      }
      // P^T = exp(S^T * scale - lse[query]); dS^T = P^T * (dP^T - delta[query]) * scale
#pragma unroll
      for (int n = 0; n < ATT_BLK / 8; ++n) {
        const int c0 = n * 8 + 2 * q;  // local query column
        const int qi0 = i0 + c0, qi1 = qi0 + 1;
        const float l0 = s_lse[c0], l1 = s_lse[c0 + 1];
        const float kLog2e = 1.4426950408889634f;
        const bool ok00 = (key_a >= lo) && (key_a <= qi0) && (qi0 < S_valid);
        const bool ok01 = (key_a >= lo) && (key_a <= qi1) && (qi1 < S_valid);
        const bool ok10 = (key_b >= lo) && (key_b <= qi0) && (qi0 < S_valid);
        const bool ok11 = (key_b >= lo) && (key_b <= qi1) && (qi1 < S_valid);
        const float p00 = ok00 ? exp2f(st[n][0] * scale_log2 - l0 * kLog2e) : 0.f;
        const float p01 = ok01 ? exp2f(st[n][1] * scale_log2 - l1 * kLog2e) : 0.f;
        const float p10 = ok10 ? exp2f(st[n][2] * scale_log2 - l0 * kLog2e) : 0.f;
        const float p11 = ok11 ? exp2f(st[n][3] * scale_log2 - l1 * kLog2e) : 0.f;
        const float d0 = s_delta[c0], d1 = s_delta[c0 + 1];
        dpt[n][0] = p00 * (dpt[n][0] - d0) * scale;
        dpt[n][1] = p01 * (dpt[n][1] - d1) * scale;
        dpt[n][2] = p10 * (dpt[n][2] - d0) * scale;
        dpt[n][3] = p11 * (dpt[n][3] - d1) * scale;
        st[n][0] = p00; st[n][1] = p01; st[n][2] = p10; st[n][3] = p11;
      }
      // dV += P^T dO ; dK += dS^T Q   (k index = query: B fragments from the transposed tiles)
#pragma unroll
      for (int kk = 0; kk < ATT_BLK / 16; ++kk) {
        uint32_t pa[4], sa[4];
        pa[0] = pack_bf16x2(st[2 * kk][0], st[2 * kk][1]);
        pa[1] = pack_bf16x2(st[2 * kk][2], st[2 * kk][3]);
        pa[2] = pack_bf16x2(st[2 * kk + 1][0], st[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
        sa[0] = pack_bf16x2(dpt[2 * kk][0], dpt[2 * kk][1]);
        sa[1] = pack_bf16x2(dpt[2 * kk][2], dpt[2 * kk][3]);
        sa[2] = pack_bf16x2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]);
        sa[3] = pack_bf16x2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) {
          mma_bf16(dv[n], pa, sdOt.ld32(n * 8 + g, kk * 16 + 2 * q), sdOt.ld32(n * 8 + g, kk * 16 + 8 + 2 * q));
          mma_bf16(dk[n], sa, sQt.ld32(n * 8 + g, kk * 16 + 2 * q), sQt.ld32(n * 8 + g, kk * 16 + 8 + 2 * q));
        }
      }
    }
  }
  bf16* dkb = dqkv + (long long)b * S * pitch + d + h * DH;
  bf16* dvb = dqkv + (long long)b * S * pitch + 2 * d + h * DH;
  if (key_a < S) {
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(dkb + (long long)key_a * pitch + n * 8 + 2 * q) = pack_bf16x2(dk[n][0], dk[n][1]);
      *reinterpret_cast<uint32_t*>(dvb + (long long)key_a * pitch + n * 8 + 2 * q) = pack_bf16x2(dv[n][0], dv[n][1]);
    }
  }
  if (key_b < S) {
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(dkb + (long long)key_b * pitch + n * 8 + 2 * q) = pack_bf16x2(dk[n][2], dk[n][3]);
      *reinterpret_cast<uint32_t*>(dvb + (long long)key_b * pitch + n * 8 + 2 * q) = pack_bf16x2(dv[n][2], dv[n][3]);
    }
  }
}

// dQ: one CTA owns 64 queries of one head and sweeps the key tiles at or above... below the diagonal.
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                  const float* __restrict__ lse, const float* __restrict__ delta,
                                                                  const int32_t* __restrict__ first_valid, bf16* __restrict__ dqkv,
                                                                  int S, int S_valid, int H, float scale, float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile<DH> sQ{reinterpret_cast<bf16*>(smem_raw)};
  Tile<DH> sdO{sQ.p + ATT_BLK * Tile<DH>::PITCH};
  Tile<DH> sK{sdO.p + ATT_BLK * Tile<DH>::PITCH};
  Tile<DH> sV{sK.p + ATT_BLK * Tile<DH>::PITCH};
  Tile<ATT_BLK> sKt{sV.p + ATT_BLK * Tile<DH>::PITCH};  // [DH][64 keys]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_BLK;
  const int d = H * DH;
  const long long pitch = 3LL * d;
  const bf16* base = qkv + (long long)b * S * pitch;
  const bf16* dob = dout + (long long)b * S * d + h * DH;
  const int lo = first_valid[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  const int q_hi = min(q0 + ATT_BLK, S_valid);
  const float kLog2e = 1.4426950408889634f;

  float dq[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;

  if (q_hi > lo && q_hi > q0) {
    const float* lse_b = lse + ((long long)b * H + h) * S;
    const float* delta_b = delta + ((long long)b * H + h) * S;
    const float lse_a = (row_a < S) ? lse_b[row_a] * kLog2e : INFINITY, lse_bb = (row_b < S) ? lse_b[row_b] * kLog2e : INFINITY;
    const float del_a = (row_a < S) ? delta_b[row_a] : 0.f, del_b = (row_b < S) ? delta_b[row_b] : 0.f;
    load_tile<DH>(sQ, base + h * DH, pitch, q0, S);
    load_tile<DH>(sdO, dob, d, q0, S);
    __syncthreads();
    uint32_t qa[DH / 16][4], doa[DH / 16][4];
#pragma unroll
    for (int k = 0; k < DH / 16; ++k) {
      load_a_frag<DH>(sQ, warp * 16, k * 16, lane, qa[k]);
      load_a_frag<DH>(sdO, warp * 16, k * 16, lane, doa[k]);
    }
    const int j_begin = (lo / ATT_BLK) * ATT_BLK;
    for (int j0 = j_begin; j0 < q_hi; j0 += ATT_BLK) {
      __syncthreads();
      load_tile<DH>(sK, base + d + h * DH, pitch, j0, S);
      load_tile<DH>(sV, base + 2 * d + h * DH, pitch, j0, S);
      load_tile_t<DH>(sKt, base + d + h * DH, pitch, j0, S);
      __syncthreads();
      float s[ATT_BLK / 8][4], dp[ATT_BLK / 8][4];
#pragma unroll
      for (int n = 0; n < ATT_BLK / 8; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          mma_bf16(s[n], qa[k], sK.ld32(n * 8 + g, k * 16 + 2 * q), sK.ld32(n * 8 + g, k * 16 + 8 + 2 * q));
          mma_bf16(dp[n], doa[k], sV.ld32(n * 8 + g, k * 16 + 2 * q), sV.ld32(n * 8 + g, k * 16 + 8 + 2 * q));
        }
      }
#pragma unroll
      for (int n = 0; n < ATT_BLK / 8; ++n) {
        const int key = j0 + n * 8 + 2 * q;
        const bool va = row_a < S_valid, vb = row_b < S_valid;
        const float p0 = (va && key >= lo && key <= row_a) ? exp2f(s[n][0] * scale_log2 - lse_a) : 0.f;
        const float p1 = (va && key + 1 >= lo && key + 1 <= row_a) ? exp2f(s[n][1] * scale_log2 - lse_a) : 0.f;
        const float p2 = (vb && key >= lo && key <= row_b) ? exp2f(s[n][2] * scale_log2 - lse_bb) : 0.f;
        const float p3 = (vb && key + 1 >= lo && key + 1 <= row_b) ? exp2f(s[n][3] * scale_log2 - lse_bb) : 0.f;
        s[n][0] = p0 * (dp[n][0] - del_a) * scale;
        s[n][1] = p1 * (dp[n][1] - del_a) * scale;
        s[n][2] = p2 * (dp[n][2] - del_b) * scale;
        s[n][3] = p3 * (dp[n][3] - del_b) * scale;
      }
#pragma unroll
      for (int kk = 0; kk < ATT_BLK / 16; ++kk) {
        uint32_t sa[4];
        sa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        sa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        sa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        sa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          mma_bf16(dq[n], sa, sKt.ld32(n * 8 + g, kk * 16 + 2 * q), sKt.ld32(n * 8 + g, kk * 16 + 8 + 2 * q));
      }
    }
  }
  bf16* dqb = dqkv + (long long)b * S * pitch + h * DH;
  if (row_a < S) {
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
      *reinterpret_cast<uint32_t*>(dqb + (long long)row_a * pitch + n * 8 + 2 * q) = pack_bf16x2(dq[n][0], dq[n][1]);
  }
  if (row_b < S) {
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
      *reinterpret_cast<uint32_t*>(dqb + (long long)row_b * pitch + n * 8 + 2 * q) = pack_bf16x2(dq[n][2], dq[n][3]);
  }
}

template <int DH>
static size_t fwd_smem() { return (size_t)(2 * ATT_BLK * Tile<DH>::PITCH + DH * Tile<ATT_BLK>::PITCH) * sizeof(bf16); }
template <int DH>
static size_t dkv_smem() { return (size_t)(4 * ATT_BLK * Tile<DH>::PITCH + 2 * DH * Tile<ATT_BLK>::PITCH) * sizeof(bf16) + 2 * ATT_BLK * sizeof(float); }
template <int DH>
static size_t dq_smem() { return (size_t)(4 * ATT_BLK * Tile<DH>::PITCH + DH * Tile<ATT_BLK>::PITCH) * sizeof(bf16); }

template <int DH>
static int launch_fwd(const bf16* qkv, const int32_t* fv, bf16* out, bf16* out2, float* lse, int B, int S, int S_valid, int H, int out_f16, cudaStream_t st) {
  const size_t smem = fwd_smem<DH>();
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attn_fwd)");
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)DH);
  dim3 grid((S + ATT_BLK - 1) / ATT_BLK, H, B);
  attn_fwd_kernel<DH><<<grid, ATT_THREADS, smem, st>>>(qkv, fv, out, out2, lse, S, S_valid, H, scale_log2, out_f16);
  NEKO_LAUNCH_CHECK("attn_fwd_kernel");
  return NEKO_OK;
}

template <int DH>
static int launch_bwd(const bf16* qkv, const bf16* dout, const float* lse, const float* delta, const int32_t* fv, bf16* dqkv, int B,
                      int S, int S_valid, int H, cudaStream_t st) {
  const size_t s1 = dkv_smem<DH>(), s2 = dq_smem<DH>();
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attn_bwd_dkv)");
  e = cudaFuncSetAttribute(attn_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attn_bwd_dq)");
  const float scale = 1.0f / sqrtf((float)DH);
  const float scale_log2 = 1.4426950408889634f * scale;
  dim3 grid((S + ATT_BLK - 1) / ATT_BLK, H, B);
  attn_bwd_dkv_kernel<DH><<<grid, ATT_THREADS, s1, st>>>(qkv, dout, lse, delta, fv, dqkv, S, S_valid, H, scale, scale_log2);
  NEKO_LAUNCH_CHECK("attn_bwd_dkv_kernel");
  attn_bwd_dq_kernel<DH><<<grid, ATT_THREADS, s2, st>>>(qkv, dout, lse, delta, fv, dqkv, S, S_valid, H, scale, scale_log2);
  NEKO_LAUNCH_CHECK("attn_bwd_dq_kernel");
  return NEKO_OK;
}

}  // namespace neko

extern "C" {

int neko_attention_fwd(const uint16_t* qkv, const int32_t* first_valid, uint16_t* out, uint16_t* out2_bf16, float* lse, int B, int S,
                       int S_valid, int H, int dh, int out_f16, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(qkv && first_valid && out && lse, "attention_fwd: null pointer");
  NEKO_REQUIRE(B > 0 && S > 0 && H > 0 && S_valid > 0 && S_valid <= S, "attention_fwd: bad sizes");
  NEKO_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention_fwd: misaligned");
  const bf16* x = reinterpret_cast<const bf16*>(qkv);
  bf16* o = reinterpret_cast<bf16*>(out);
  bf16* o2 = reinterpret_cast<bf16*>(out2_bf16);
  cudaStream_t st = as_stream(stream);
  switch (dh) {
    case 16: return launch_fwd<16>(x, first_valid, o, o2, lse, B, S, S_valid, H, out_f16, st);
    case 32: return launch_fwd<32>(x, first_valid, o, o2, lse, B, S, S_valid, H, out_f16, st);
    case 64: return launch_fwd<64>(x, first_valid, o, o2, lse, B, S, S_valid, H, out_f16, st);
    case 128: return launch_fwd<128>(x, first_valid, o, o2, lse, B, S, S_valid, H, out_f16, st);
    default: set_error("attention: head dim %d not supported (16, 32, 64, 128)", dh); return NEKO_EINVAL;
  }
}

int neko_attention_bwd(const uint16_t* qkv, const uint16_t* out, const uint16_t* dout, const float* lse, const int32_t* first_valid,
                       uint16_t* dqkv, float* delta, int B, int S, int S_valid, int H, int dh, int out_f16, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(qkv && out && dout && lse && first_valid && dqkv && delta, "attention_bwd: null pointer");
  NEKO_REQUIRE(B > 0 && S > 0 && H > 0 && S_valid > 0 && S_valid <= S, "attention_bwd: bad sizes");
  cudaStream_t st = as_stream(stream);
  const bf16* x = reinterpret_cast<const bf16*>(qkv);
  const bf16* o = reinterpret_cast<const bf16*>(out);
  const bf16* g = reinterpret_cast<const bf16*>(dout);
  bf16* dx = reinterpret_cast<bf16*>(dqkv);
  const long long warps = (long long)B * S * H;
  attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(o, g, delta, B, S, H, dh, out_f16);
  NEKO_LAUNCH_CHECK("attn_delta_kernel");
  switch (dh) {
    case 16: return launch_bwd<16>(x, g, lse, delta, first_valid, dx, B, S, S_valid, H, st);
    case 32: return launch_bwd<32>(x, g, lse, delta, first_valid, dx, B, S, S_valid, H, st);
    case 64: return launch_bwd<64>(x, g, lse, delta, first_valid, dx, B, S, S_valid, H, st);
    case 128: return launch_bwd<128>(x, g, lse, delta, first_valid, dx, B, S, S_valid, H, st);
    default: set_error("attention: head dim %d not supported (16, 32, 64, 128)", dh); return NEKO_EINVAL;
  }
}

}  // extern "C"
