// Counter-based dropout shared by every kernel that applies or re-applies a mask
// (nn.Dropout at trajectory_gpt2.py:179 attention weights, :254 / :278 residual branches, :707 embeddings).
//
// The keep decision of element (row, col) of dropout site `stream` is a pure function of (seed, stream, row, col), so the
// backward kernels regenerate the forward's mask instead of storing it, whatever their thread -> element mapping:
//   key    = H(seed[0] ^ H(seed[1] + stream * G))                 once per thread
//   rowkey = H(key + row * G2)                                     once per row
//   bits   = H(rowkey ^ ((col >> 1) * G))                          16 bits per element, two adjacent columns per hash
//   keep   = bits16 >= thr16            (thr16 = round(p * 65536); kept values are scaled by 65536 / (65536 - thr16))
// `seed` lives in device memory (two words written by the host before every forward) so a captured CUDA graph
// draws fresh masks on every replay.
#pragma once
#include <cstdint>
#include "../../include/neko_b200.h"

namespace neko {

struct DropCfg {
  const uint32_t* seed;  // device pointer, nullptr = dropout disabled
  uint32_t stream, thr16;
  float scale;
};
__host__ __device__ inline DropCfg drop_cfg(const neko_dropout* d) {
  DropCfg c{nullptr, 0u, 0u, 1.0f};
  if (d && d->seed && d->thr16) { c.seed = d->seed; c.stream = d->stream; c.thr16 = d->thr16; c.scale = d->scale; }
  return c;
}
__device__ __forceinline__ uint32_t drop_hash(uint32_t x) {  // "lowbias32" integer finaliser
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
  return x;
}
__device__ __forceinline__ uint32_t drop_key(const DropCfg& c) {
  return drop_hash(__ldg(c.seed) ^ drop_hash(__ldg(c.seed + 1) + c.stream * 0x9E3779B1u));
}
__device__ __forceinline__ uint32_t drop_rowkey(uint32_t key, uint32_t row) { return drop_hash(key + row * 0x85EBCA6Bu); }
// bits of columns (2*pair, 2*pair+1): low / high half-word
__device__ __forceinline__ uint32_t drop_bits(uint32_t rowkey, uint32_t pair) { return drop_hash(rowkey ^ (pair * 0x9E3779B1u)); }
// multipliers (0 or scale) of an even/odd column pair
__device__ __forceinline__ void drop_pair(uint32_t rowkey, uint32_t pair, uint32_t thr16, float scale, float& m0, float& m1) {
  const uint32_t b = drop_bits(rowkey, pair);
  m0 = (b & 0xffffu) >= thr16 ? scale : 0.f;
  m1 = (b >> 16) >= thr16 ? scale : 0.f;
}

}  // namespace neko
