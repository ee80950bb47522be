// Gradient all-reduce over NVLink peer memory, written to run NEXT TO the persistent GEMMs of backward.
//
// Replaces the reference's DDP bucket all-reduce (train.py:26-40,107; trainer.py:176-186: SUM then / world).
// Why not NCCL here: NCCL's kernels need whole SMs for as long as a bucket is in flight, and a persistent GEMM with one
// 227 KB CTA per SM cannot start its CTA on such an SM -- every GEMM of backward then ends only when the bucket does
// (measured: the "overlapped" NCCL path costs exactly as much as reducing everything after backward).
// This kernel is sized to CO-RESIDE with a GEMM CTA instead: 128 threads, <= 80 registers, no shared memory
// (the GEMM leaves 1.75 KB of shared memory and 11.7 K registers per SM), one CTA per SM.  It is bandwidth-light by
// design: the exchange only has to finish within backward (~100-200 GB/s of the 900 GB/s links).
//
// Algorithm (two-shot, in place, fp32, deterministic): every rank's gradient arena is mapped into every process
// (cudaIpc*).  For the bucket [lo, hi):
//   barrier A   all ranks' bucket data is final (each rank's launch is stream-ordered after its own last writer)
//   phase 1     rank r reads slice r of the bucket from EVERY rank (peer loads), sums in rank order, scales by 1/world
//               and overwrites slice r of its own arena                                     (reduce-scatter)
//   barrier B   all slices reduced
//   phase 2     rank r copies the reduced slices p != r out of rank p's arena into its own   (all-gather)
//   barrier C   nobody still reads this rank's arena: the kernel may exit (the next step will overwrite it)
// Cross-rank barriers: monotonically increasing sequence numbers in an IPC-mapped signal buffer (one word per source
// rank), st.release.sys / ld.acquire.sys; inside a rank the CTAs meet on a local atomic counter.  The launch counter the
// sequence numbers derive from is device state as well: launch arguments are constant, the kernel is CUDA-graph safe.
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace neko {

constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_THREADS = 128;

struct P2PArgs {
  float* buf[P2P_MAX_WORLD];      // every rank's arena (same layout), peer-mapped
  unsigned* sig[P2P_MAX_WORLD];   // every rank's signal words [world]
  float* stage[P2P_MAX_WORLD];    // push variant: every rank's staging buffer (world planes of plane4 float4), peer-mapped
  long long plane4, off4;         // staging plane size and this bucket's offset inside a plane (float4 units)
  unsigned* state;                // this rank's private words: [0] launches completed so far, [1] CTA arrival counter
  int rank, world;
  long long lo4, n4;              // bucket in float4 units
  float scale;
  int flags;                      // experiment switches (NEKO_P2P_FLAGS): 1 plain loads, 2 slow single-thread polling, 4 skip phase 1, 8 skip phase 2
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {   // peer data: never from a stale cache line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// all CTAs of all ranks: k-th barrier (0..2) of launch number `seq`.  Everything the barrier needs lives in device memory
// (the launch counter included), so the launch arguments never change and the kernel replays from a CUDA graph.
__device__ __forceinline__ void rank_barrier(const P2PArgs& a, unsigned seq, int k) {
  __syncthreads();
  const unsigned want = 3u * seq + (unsigned)k + 1u;
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(a.state + 1, 1u);
    if (prev + 1u == (unsigned)(k + 1) * gridDim.x) {   // last CTA of this rank to get here
      if (k == 2) {                                      // every CTA has read the launch counter long ago: advance it
        a.state[1] = 0u;
        a.state[0] = seq + 1u;
        __threadfence();
      }
      for (int p = 0; p < a.world; ++p) st_release_sys(a.sig[p] + a.rank, want);
    }
  }
  if (a.flags & 2) {
    if (threadIdx.x == 0) {
      for (int q = 0; q < a.world; ++q) {
        const unsigned* mine = a.sig[a.rank] + q;
        while ((int)(*reinterpret_cast<const volatile unsigned*>(mine) - want) < 0) __nanosleep(1000);
      }
      __threadfence_system();
    }
  } else if (threadIdx.x < a.world) {
    const unsigned* mine = a.sig[a.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(mine) - want) < 0) __nanosleep(64);
  }
  __syncthreads();
}

template <int WORLD>
__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_kernel(const P2PArgs a) {
  const long long per = (a.n4 + WORLD - 1) / WORLD;
  const long long tid = (long long)blockIdx.x * P2P_THREADS + threadIdx.x;
  const long long stride = (long long)gridDim.x * P2P_THREADS;
  const unsigned seq = *reinterpret_cast<volatile unsigned*>(a.state);
  rank_barrier(a, seq, 0);
  {  // phase 1: reduce slice `rank`.  U float4 positions per thread and pass: U * WORLD independent loads in flight (the
     // exchange is latency-bound: ~1.5 us per peer load, so bandwidth = bytes in flight / latency)
    constexpr int U = (16 / WORLD) < 1 ? 1 : (16 / WORLD);
    const long long s0 = a.lo4 + (long long)a.rank * per;
    const long long s1 = a.lo4 + min(a.n4, (long long)(a.rank + 1) * per);
    float4* mine = reinterpret_cast<float4*>(a.buf[a.rank]);
    long long i = s0 + tid;
    for (; i + (U - 1) * stride < s1; i += U * stride) {
      float4 v[U][WORLD];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int p = 0; p < WORLD; ++p) v[u][p] = ld_sys_f4(reinterpret_cast<const float4*>(a.buf[p]) + i + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 s = v[u][0];
#pragma unroll
        for (int p = 1; p < WORLD; ++p) { s.x += v[u][p].x; s.y += v[u][p].y; s.z += v[u][p].z; s.w += v[u][p].w; }
        s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
        mine[i + u * stride] = s;
      }
    }
    for (; i < s1; i += stride) {
      float4 v[WORLD];
#pragma unroll
      for (int p = 0; p < WORLD; ++p) v[p] = ld_sys_f4(reinterpret_cast<const float4*>(a.buf[p]) + i);
      float4 s = v[0];
#pragma unroll
      for (int p = 1; p < WORLD; ++p) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
      s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
      mine[i] = s;
    }
  }
  rank_barrier(a, seq, 1);
  {  // phase 2: gather the other ranks' reduced slices (staggered start so the ranks do not all hit one peer)
    float4* mine = reinterpret_cast<float4*>(a.buf[a.rank]);
#pragma unroll 1
    for (int pp = 1; pp < WORLD; ++pp) {
      const int p = (a.rank + pp) % WORLD;
      const long long s0 = a.lo4 + (long long)p * per;
      const long long s1 = a.lo4 + min(a.n4, (long long)(p + 1) * per);
      const float4* src = reinterpret_cast<const float4*>(a.buf[p]);
      long long i = s0 + tid;
      for (; i + 7 * stride < s1; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_sys_f4(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) mine[i + u * stride] = v[u];
      }
      for (; i < s1; i += stride) mine[i] = ld_sys_f4(src + i);
    }
  }
  rank_barrier(a, seq, 2);
}

// PUSH variant (the default): all traffic over NVLink is posted WRITES, as in NCCL's protocols.  Peer loads hold a
// request slot for the whole ~1.5 us round trip, and enough of them in flight to fill the link back up into the fabric
// the GEMMs' operand reads share (measured: a GEMM chain next to the pull kernel made almost no progress, tools/
// dp_overlap_probe.py); writes are fire-and-forget.
//   phase 1     rank r copies slice p of its bucket into plane r of rank p's staging buffer, for every p != r
//   barrier 0   all contributions have landed
//   phase 2     rank r sums its own slice r and the world-1 staged copies in rank order, scales, and writes the result
//               into slice r of EVERY rank's arena (its own included)
//   barrier 1   all reduced slices have landed everywhere
//   (barrier 2 only keeps the bookkeeping of rank_barrier uniform: one extra flag round, a few microseconds)
// Staging regions of consecutive buckets are disjoint (off4 advances), so a rank that runs ahead never overwrites what
// a slower peer is still reducing.
template <int WORLD>
__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_push_kernel(const P2PArgs a) {
  const long long per = (a.n4 + WORLD - 1) / WORLD;
  const long long tid = (long long)blockIdx.x * P2P_THREADS + threadIdx.x;
  const long long stride = (long long)gridDim.x * P2P_THREADS;
  const unsigned seq = *reinterpret_cast<volatile unsigned*>(a.state);
  float4* mine = reinterpret_cast<float4*>(a.buf[a.rank]);
  if (!(a.flags & 4)) {  // phase 1: push
#pragma unroll 1
    for (int pp = 1; pp < WORLD; ++pp) {
      const int p = (a.rank + pp) % WORLD;
      const long long len = min(a.n4, (long long)(p + 1) * per) - (long long)p * per;
      const float4* src = mine + a.lo4 + (long long)p * per;
      float4* dst = reinterpret_cast<float4*>(a.stage[p]) + (long long)a.rank * a.plane4 + a.off4;
      long long i = tid;
      for (; i + 3 * stride < len; i += 4 * stride) {
        const float4 v0 = src[i], v1 = src[i + stride], v2 = src[i + 2 * stride], v3 = src[i + 3 * stride];
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
      }
      for (; i < len; i += stride) dst[i] = src[i];
    }
  }
  rank_barrier(a, seq, 0);
  if (!(a.flags & 8)) {  // phase 2: reduce slice `rank` (local loads only), broadcast the result
    const long long s0 = (long long)a.rank * per;
    const long long len = min(a.n4, s0 + per) - s0;
    const float4* stg = reinterpret_cast<const float4*>(a.stage[a.rank]) + a.off4;
    for (long long i = tid; i < len; i += stride) {
      float4 v[WORLD];
#pragma unroll
      for (int q = 0; q < WORLD; ++q)
        v[q] = (q == a.rank || (a.flags & 1)) ? (q == a.rank ? mine[a.lo4 + s0 + i] : stg[(long long)q * a.plane4 + i])
                                              : ld_sys_f4(stg + (long long)q * a.plane4 + i);
      float4 s = v[0];
#pragma unroll
      for (int q = 1; q < WORLD; ++q) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
      s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
#pragma unroll
      for (int p = 0; p < WORLD; ++p) reinterpret_cast<float4*>(a.buf[p])[a.lo4 + s0 + i] = s;
    }
  }
  rank_barrier(a, seq, 1);
  rank_barrier(a, seq, 2);
}

// ---------------------------------------------------------------------------------------------
// COPY-ENGINE variant (the default): the SM-resident kernels above share their SM's load/store path with the GEMM CTA
// next to them, and while they stream data the GEMM stalls behind their memory instructions (measured: a GEMM chain
// loses ~80 % of the all-reduce's duration whether the exchange pulls or pushes, tools/dp_overlap_probe.py).  Here
// NVLink traffic is moved by the DMA engines (cudaMemcpyAsync between peer-mapped buffers, issued by the host side);
// the only kernels are a one-warp flag exchange and the local reduction of the staged planes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) p2p_signal_wait_kernel(const P2PArgs a) {
  // channel c = a.flags (0..3): independent barrier sequences, so that exchanges running on different streams do not have
  // to agree on one global order.  state[2 + c]: launches on this channel so far; signal words 32 + 8 c .. belong to it
  const int ch = a.flags;
  const unsigned want = *reinterpret_cast<volatile unsigned*>(a.state + 2 + ch) + 1u;
  __threadfence_system();
  if ((int)threadIdx.x < a.world) {
    st_release_sys(a.sig[threadIdx.x] + 32 + 8 * ch + a.rank, want);
    const unsigned* mine = a.sig[a.rank] + 32 + 8 * ch + threadIdx.x;
    while ((int)(ld_acquire_sys(mine) - want) < 0) __nanosleep(100);
  }
  __syncwarp();
  if (threadIdx.x == 0) a.state[2 + ch] = want;
}

// dst[i] = scale * sum over planes q in rank order (plane `self` is dst itself, the others are the staged contributions)
template <int WORLD>
__global__ void __launch_bounds__(256) reduce_planes_kernel(float4* __restrict__ dst, const float4* __restrict__ stage, long long plane4,
                                                            int self, long long n4, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v[WORLD];
#pragma unroll
    for (int q = 0; q < WORLD; ++q) v[q] = (q == self) ? dst[i] : ld_sys_f4(stage + (long long)q * plane4 + i);
    float4 s = v[0];
#pragma unroll
    for (int q = 1; q < WORLD; ++q) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    dst[i] = s;
  }
}

// occupies n CTAs of `threads` threads for `ns` nanoseconds (no shared memory, a handful of registers)
__global__ void spin_kernel(unsigned long long ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 >= ns) break;
    __nanosleep(200);
  }
}

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

}  // namespace neko

extern "C" {

int neko_debug_spin(int n_ctas, int threads, long long ns, int max_shared_carveout, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(n_ctas >= 1 && threads >= 1 && threads <= 1024 && ns >= 0, "debug_spin: bad arguments");
  cudaError_t e = cudaFuncSetAttribute(spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       max_shared_carveout ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(spin)");
  spin_kernel<<<n_ctas, threads, 0, as_stream(stream)>>>((unsigned long long)ns);
  NEKO_LAUNCH_CHECK("spin_kernel");
  return NEKO_OK;
}

int neko_memcpy_async(void* dst, const void* src, long long bytes, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(dst && src && bytes >= 0, "memcpy_async: bad arguments");
  if (bytes == 0) return NEKO_OK;
  return check_cuda(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, as_stream(stream)), "cudaMemcpyAsync");
}

int neko_p2p_signal_wait(void* const* host_sigs, unsigned* state, int rank, int world, int channel, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(host_sigs && state && world >= 2 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world && channel >= 0 && channel < 4,
               "p2p_signal_wait: bad arguments");
  P2PArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    NEKO_REQUIRE(host_sigs[p], "p2p_signal_wait: null signal pointer %d", p);
    a.sig[p] = static_cast<unsigned*>(host_sigs[p]);
  }
  a.state = state; a.rank = rank; a.world = world; a.flags = channel;
  static bool carve = false;
  if (!carve) {
    cudaError_t e = cudaFuncSetAttribute(p2p_signal_wait_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(signal_wait carveout)");
    carve = true;
  }
  p2p_signal_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(a);
  NEKO_LAUNCH_CHECK("p2p_signal_wait_kernel");
  return NEKO_OK;
}

int neko_reduce_planes_f32(float* dst, const float* stage, long long plane, int n_planes, int self, long long n, float scale, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(dst && stage && n >= 0 && n % 4 == 0 && plane % 4 == 0 && n_planes >= 2 && n_planes <= P2P_MAX_WORLD && self >= 0 && self < n_planes,
               "reduce_planes: bad arguments");
  NEKO_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(stage)) & 15) == 0, "reduce_planes: misaligned");
  if (n == 0) return NEKO_OK;
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 2LL * sm_count()) blocks = 2LL * sm_count();
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(stage);
  cudaStream_t s = as_stream(stream);
  switch (n_planes) {
    case 2: reduce_planes_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(d4, s4, plane / 4, self, n4, scale); break;
    case 3: reduce_planes_kernel<3><<<(unsigned)blocks, 256, 0, s>>>(d4, s4, plane / 4, self, n4, scale); break;
    case 4: reduce_planes_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(d4, s4, plane / 4, self, n4, scale); break;
    case 5: reduce_planes_kernel<5><<<(unsigned)blocks, 256, 0, s>>>(d4, s4, plane / 4, self, n4, scale); break;
    case 6: reduce_planes_kernel<6><<<(unsigned)blocks, 256, 0, s>>>(d4, s4, plane / 4, self, n4, scale); break;
    case 7: reduce_planes_kernel<7><<<(unsigned)blocks, 256, 0, s>>>(d4, s4, plane / 4, self, n4, scale); break;
    default: reduce_planes_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(d4, s4, plane / 4, self, n4, scale); break;
  }
  NEKO_LAUNCH_CHECK("reduce_planes_kernel");
  return NEKO_OK;
}

int neko_prefer_shared_carveout(int on) {
  // Device-wide default for kernels that state no preference of their own: the largest shared-memory split, i.e. the one
  // the GEMM and the all-reduce CTAs run with.  An SM re-partitions L1 / shared memory only when it is empty, so a kernel
  // asking for another split cannot join an SM on which an all-reduce CTA is resident: it would wait for the bucket.
  return neko::check_cuda(cudaDeviceSetCacheConfig(on ? cudaFuncCachePreferShared : cudaFuncCachePreferNone), "cudaDeviceSetCacheConfig");
}

int neko_ipc_export(const void* dev_ptr, unsigned char* handle_out, long long* offset_out) {
  using namespace neko;
  NEKO_REQUIRE(dev_ptr && handle_out && offset_out, "ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) { set_error("cuMemGetAddressRange not available"); return NEKO_ECUDA; }
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = reinterpret_cast<GetAddressRangeFn>(fn)(&base, &size, (CUdeviceptr)dev_ptr);
  if (r != CUDA_SUCCESS) { set_error("cuMemGetAddressRange failed (%d)", (int)r); return NEKO_ECUDA; }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) return check_cuda(e, "cudaIpcGetMemHandle (the buffer must come from cudaMalloc, not from a VMM / async pool)");
  memcpy(handle_out, &h, 64);
  *offset_out = (long long)((CUdeviceptr)dev_ptr - base);
  return NEKO_OK;
}

int neko_ipc_import(const unsigned char* handle, long long offset, void** dev_ptr_out) {
  using namespace neko;
  NEKO_REQUIRE(handle && dev_ptr_out && offset >= 0, "ipc_import: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* base = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return check_cuda(e, "cudaIpcOpenMemHandle");
  *dev_ptr_out = static_cast<char*>(base) + offset;
  return NEKO_OK;
}

int neko_ipc_close(void* dev_ptr, long long offset) {
  using namespace neko;
  if (!dev_ptr) return NEKO_OK;
  return check_cuda(cudaIpcCloseMemHandle(static_cast<char*>(dev_ptr) - offset), "cudaIpcCloseMemHandle");
}

int neko_p2p_allreduce_f32(void* const* host_bufs, void* const* host_sigs, void* const* host_stage, long long stage_plane, long long stage_off,
                           unsigned* state, int rank, int world, long long lo, long long hi, float scale, int n_ctas, void* stream) {
  using namespace neko;
  NEKO_REQUIRE(host_bufs && host_sigs && state, "p2p_allreduce: null argument");
  NEKO_REQUIRE(host_stage == nullptr || (stage_plane % 4 == 0 && stage_off % 4 == 0 && stage_off >= 0 &&
                                         stage_off + (hi - lo + world - 1) / world + 4 <= stage_plane),
               "p2p_allreduce: staging plane too small (plane %lld, offset %lld, bucket %lld elements)", stage_plane, stage_off, hi - lo);
  NEKO_REQUIRE(world >= 2 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "p2p_allreduce: world %d rank %d (2..%d ranks)", world, rank, P2P_MAX_WORLD);
  NEKO_REQUIRE(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, "p2p_allreduce: [lo, hi) must be multiples of 4 elements");
  NEKO_REQUIRE(n_ctas >= 1 && n_ctas <= sm_count(), "p2p_allreduce: 1..%d CTAs (one per SM: they spin on each other)", sm_count());
  if (hi == lo) return NEKO_OK;
  P2PArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    NEKO_REQUIRE(host_bufs[p] && host_sigs[p], "p2p_allreduce: null peer pointer %d", p);
    NEKO_REQUIRE((reinterpret_cast<uintptr_t>(host_bufs[p]) & 15) == 0, "p2p_allreduce: misaligned buffer");
    a.buf[p] = static_cast<float*>(host_bufs[p]);
    a.sig[p] = static_cast<unsigned*>(host_sigs[p]);
    if (host_stage) {
      NEKO_REQUIRE(host_stage[p] && (reinterpret_cast<uintptr_t>(host_stage[p]) & 15) == 0, "p2p_allreduce: bad staging pointer %d", p);
      a.stage[p] = static_cast<float*>(host_stage[p]);
    }
  }
  a.plane4 = stage_plane / 4; a.off4 = stage_off / 4;
  a.flags = getenv("NEKO_P2P_FLAGS") ? atoi(getenv("NEKO_P2P_FLAGS")) : 0;
  a.state = state; a.rank = rank; a.world = world; a.lo4 = lo / 4; a.n4 = (hi - lo) / 4;
  a.scale = scale;
  cudaStream_t s = as_stream(stream);
  // An SM changes its L1 / shared-memory split only when it is empty.  Ask for the same split the GEMM runs with (all
  // shared memory): otherwise a GEMM CTA cannot join an SM on which one of these CTAs sits -- and the other way round.
  static bool carveout_set = false;
  if (!carveout_set) {
    const void* fns[] = {(const void*)p2p_allreduce_push_kernel<2>, (const void*)p2p_allreduce_push_kernel<3>, (const void*)p2p_allreduce_push_kernel<4>,
                         (const void*)p2p_allreduce_push_kernel<5>, (const void*)p2p_allreduce_push_kernel<6>, (const void*)p2p_allreduce_push_kernel<7>,
                         (const void*)p2p_allreduce_push_kernel<8>, (const void*)p2p_allreduce_kernel<2>, (const void*)p2p_allreduce_kernel<3>, (const void*)p2p_allreduce_kernel<4>,
                         (const void*)p2p_allreduce_kernel<5>, (const void*)p2p_allreduce_kernel<6>, (const void*)p2p_allreduce_kernel<7>,
                         (const void*)p2p_allreduce_kernel<8>};
    for (const void* f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(p2p carveout)");
    }
    carveout_set = true;
  }
  if (host_stage) {
    switch (world) {
      case 2: p2p_allreduce_push_kernel<2><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
      case 3: p2p_allreduce_push_kernel<3><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
      case 4: p2p_allreduce_push_kernel<4><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
      case 5: p2p_allreduce_push_kernel<5><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
      case 6: p2p_allreduce_push_kernel<6><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
      case 7: p2p_allreduce_push_kernel<7><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
      default: p2p_allreduce_push_kernel<8><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
    }
    NEKO_LAUNCH_CHECK("p2p_allreduce_push_kernel");
    return NEKO_OK;
  }
  switch (world) {
    case 2: p2p_allreduce_kernel<2><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
    case 3: p2p_allreduce_kernel<3><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
    case 4: p2p_allreduce_kernel<4><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
    case 5: p2p_allreduce_kernel<5><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
    case 6: p2p_allreduce_kernel<6><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
    case 7: p2p_allreduce_kernel<7><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
    default: p2p_allreduce_kernel<8><<<n_ctas, P2P_THREADS, 0, s>>>(a); break;
  }
  NEKO_LAUNCH_CHECK("p2p_allreduce_kernel");
  return NEKO_OK;
}

}  // extern "C"
