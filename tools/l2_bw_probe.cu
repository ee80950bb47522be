// L2 -> shared-memory operand bandwidth probe: the ceiling of the GEMM main loop.
//
// One thread per CTA streams 32 KB stages (256 rows x 128 B, 128B-swizzled boxes like the GEMM's operand tiles) from an
// L2-resident buffer through a STAGES-deep mbarrier ring and consumes nothing.  With a cluster of CS CTAs every CTA loads
// 1/CS of the stage and multicasts it to all CTAs of the cluster, so each SM still RECEIVES 32 KB per stage but L2 is read
// once per cluster -- what a B tile shared by the m-neighbours of a cluster would do.  Slots are recycled through an
// empty barrier that collects one arrival from every CTA of the cluster (a peer writes into my shared memory).
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l2_bw_probe tools/l2_bw_probe.cu
// run  : tools/_build/l2_bw_probe [buffer MB = 32] [iterations = 4000]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int TILE_ROWS = 256;
constexpr int ROW_BYTES = 128;
constexpr int TILE_BYTES = TILE_ROWS * ROW_BYTES;  // 32 KB per stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// data and the complete_tx land at the same CTA-relative offsets in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int CS>
__global__ void __launch_bounds__(32, 1) probe_kernel(const __grid_constant__ CUtensorMap map, int stages, int iters, int tiles_total) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * TILE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
  const uint32_t cid = CS > 1 ? cluster_id_x() : blockIdx.x;
  constexpr int SLICE_ROWS = TILE_ROWS / CS;
  constexpr int SLICE_BYTES = TILE_BYTES / CS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar0 + 8u * s, 1);              // full: my expect_tx arrival + TILE_BYTES from all CTAs of the cluster
      mbar_init(bar0 + 8u * (stages + s), CS);  // empty: every CTA of the cluster consumed the slot
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (CS > 1) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < iters + stages - 1; ++i) {
      if (i < iters) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(bar0 + 8u * (stages + s), ph ^ 1u);
        mbar_expect_tx(bar0 + 8u * s, TILE_BYTES);
        const int tile = (int)(((long long)cid * 977 + i) % tiles_total);   // clusters walk different tiles of the buffer
        const int row0 = tile * TILE_ROWS + (int)rank * SLICE_ROWS;
        const uint32_t dst = smem_u32(smem + (size_t)s * TILE_BYTES) + rank * SLICE_BYTES;
        if (CS == 1) tma_load_2d(dst, &map, bar0 + 8u * s, 0, row0);
        else         tma_load_2d_mc(dst, &map, bar0 + 8u * s, 0, row0, (uint16_t)((1u << CS) - 1u));
      }
      const int j = i - (stages - 1);
      if (j >= 0) {
        const int sj = j % stages;
        mbar_wait(bar0 + 8u * sj, (uint32_t)(j / stages) & 1u);
        if (CS == 1) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (stages + sj)) : "memory");
        } else {
          for (uint32_t r = 0; r < (uint32_t)CS; ++r) mbar_arrive_cluster(mapa_shared(bar0 + 8u * (stages + sj), r));
        }
      }
    }
  }
  if (CS > 1) cluster_sync_all(); else __syncthreads();   // peers may still write my shared memory / barriers
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS>
static void run(void* buf, long long rows, int stages, int iters, EncodeTiledFn enc, int sms) {
  CUtensorMap map;
  const cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {ROW_BYTES};
  const cuuint32_t box[2] = {64, (cuuint32_t)(TILE_ROWS / CS)};
  const cuuint32_t es[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    fprintf(stderr, "tensor map encode failed\n");
    exit(1);
  }
  const size_t smem = (size_t)stages * TILE_BYTES + 2 * stages * 8 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 4) CK(cudaFuncSetAttribute(probe_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ctas = sms / CS * CS;
  if (CS > 1) {
    cfg.gridDim = dim3(ctas);
    int nclusters = 0;
    CK(cudaOccupancyMaxActiveClusters(&nclusters, probe_kernel<CS>, &cfg));
    if (nclusters * CS < ctas) ctas = nclusters * CS;   // one wave only
  }
  cfg.gridDim = dim3(ctas);
  const int tiles_total = (int)(rows / TILE_ROWS);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, probe_kernel<CS>, map, stages, iters, tiles_total));
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, probe_kernel<CS>, map, stages, iters, tiles_total));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double delivered = (double)ctas * iters * TILE_BYTES;
  printf("cluster %d  stages %d  ctas %3d  %.3f ms  delivered to SMs %.2f TB/s  read from L2 %.2f TB/s  (%.1f B/clk/SM at 1.9 GHz)\n", CS, stages,
         ctas, best, delivered / best * 1e-9, delivered / CS / best * 1e-9, delivered / best * 1e3 / ctas / 1.9e9);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const long long mb = argc > 1 ? atoll(argv[1]) : 32;
  const int iters = argc > 2 ? atoi(argv[2]) : 4000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  if (!ptr || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ptr);
  const long long rows = mb * 1024 * 1024 / ROW_BYTES / TILE_ROWS * TILE_ROWS;
  void* buf;
  CK(cudaMalloc(&buf, (size_t)rows * ROW_BYTES));
  CK(cudaMemset(buf, 1, (size_t)rows * ROW_BYTES));
  printf("%s, %d SMs, buffer %lld MB, %d stages of 32 KB per CTA and launch\n", prop.name, sms, mb, iters);
  for (int stages : {2, 4, 6}) run<1>(buf, rows, stages, iters, enc, sms);
  for (int stages : {2, 4, 6}) run<2>(buf, rows, stages, iters, enc, sms);
  for (int stages : {2, 4, 6}) run<4>(buf, rows, stages, iters, enc, sms);
  for (int stages : {4, 6}) run<8>(buf, rows, stages, iters, enc, sms);
  return 0;
}
