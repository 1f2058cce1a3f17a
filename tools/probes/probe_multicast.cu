// Probe: how many bytes per second can land in shared memory by TMA bulk copies, per SM and chip-wide, when
//   mode 0  every CTA streams DISTINCT data                                   (unicast, the L2 -> SM ceiling),
//   mode 1  all CTAs of a cluster stream the SAME data, each with its own copy (unicast, duplicates served by L2),
//   mode 2  each CTA of a cluster loads 1/CS of the data and MULTICASTS it to all CS CTAs (one L2 read, CS deliveries).
// The dense f16x3 GEMM sits at 55 % tensor-pipe activity while its operand tiles arrive at ~23 B/cycle/SM
// (DESIGN.md section 5): this measures whether a cluster that shares weight / activation tiles by multicast could feed
// an SM faster, i.e. whether the limit is L2 output or the SM's own ingest.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_multicast probe_multicast.cu
// Run:   ./probe_multicast            (prints one line per mode x cluster size)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kTile = 16 * 1024;     // bytes per tile per CTA
constexpr int kSlots = 8;            // tiles in flight per CTA and round (128 KB of shared memory)
constexpr int kThreads = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
  for (uint32_t spins = 0; spins < (1u << 26); ++spins)
    if (mbar_try(b, parity)) return true;
  return false;       // a protocol error must not hang the GPU
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// src: a buffer much larger than L2 is NOT wanted here - the GEMM's operands are L2 hits - so the host passes a window
// of `window` bytes (a few tens of MB, L2 resident after the first pass) that the tile index wraps around in.
__global__ void __launch_bounds__(kThreads, 1) probe_kernel(const unsigned char* __restrict__ src, size_t window, int mode,
                                                            int rounds, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kSlots];
  const uint32_t cs = cluster_size(), cr = cluster_rank();
  const size_t cluster_id = blockIdx.x / cs;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const size_t ntiles = window / kTile;
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < kSlots; ++s) {
        const size_t it = (size_t)r * kSlots + s;
        mbar_expect(&full[s], kTile);
        if (mode == 0) {
          const size_t t = ((size_t)blockIdx.x * 977 + it) % ntiles;               // distinct stream per CTA
          bulk_load(smem + s * kTile, src + t * kTile, kTile, &full[s]);
        } else if (mode == 1) {
          const size_t t = (cluster_id * 977 + it) % ntiles;                       // same stream for the whole cluster
          bulk_load(smem + s * kTile, src + t * kTile, kTile, &full[s]);
        } else {
          const size_t t = (cluster_id * 977 + it) % ntiles;
          const uint32_t slice = kTile / cs;                                       // this CTA's share, delivered to all
          bulk_load_mc(smem + s * kTile + cr * slice, src + t * kTile + cr * slice, slice, &full[s],
                       (uint16_t)((1u << cs) - 1));
        }
      }
      for (int s = 0; s < kSlots; ++s)
        if (!mbar_wait(&full[s], r & 1)) atomicExch(err, 1);
    }
    __syncthreads();
    if (mode == 2) cluster_sync();      // nobody refills a slot of a CTA that has not seen its current phase complete
  }
  cluster_sync();
}

static float run(const unsigned char* src, size_t window, int mode, int cs, int rounds, int* d_err, int sms) {
  int ctas = (sms / cs) * cs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSlots * kTile;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kTile);
  if (cs > 8) cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  // one wave only: as many clusters as can be resident at once (GPC sizes limit clusters of 8)
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, probe_kernel, &cfg) == cudaSuccess && max_clusters > 0 &&
      max_clusters * cs < ctas) {
    ctas = max_clusters * cs;
    cfg.gridDim = dim3(ctas);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) cudaLaunchKernelEx(&cfg, probe_kernel, src, window, mode, rounds, d_err);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, probe_kernel, src, window, mode, rounds, d_err);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mode %d cs %d: CUDA error %s\n", mode, cs, cudaGetErrorString(e));
    exit(1);
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double landed = (double)ctas * rounds * kSlots * kTile;     // bytes that arrived in shared memory, all SMs
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("mode %d cluster %2d (%3d CTAs): %7.3f ms  landed %7.1f GB/s chip, %6.1f GB/s per SM (%5.1f B/clk/SM at %d MHz nominal)\n", mode, cs,
         ctas, ms, landed / ms / 1e6, landed / ms / 1e6 / ctas, landed / ms / 1e6 / ctas * 1e9 / (clk_khz * 1e3), clk_khz / 1000);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t window = 64ull << 20;      // 64 MB: L2 resident (126 MB)
  unsigned char* src = nullptr;
  int* d_err = nullptr;
  cudaMalloc(&src, window);
  cudaMalloc(&d_err, sizeof(int));
  cudaMemset(src, 1, window);
  cudaMemset(d_err, 0, sizeof(int));
  const int rounds = 400;                 // 400 * 8 * 16 KB = 52 MB landed per SM
  printf("SMs %d, tile %d B, %d tiles in flight per CTA, window %zu MB\n", sms, kTile, kSlots, window >> 20);
  for (int cs : {1, 2, 4, 8}) {
    run(src, window, 0, cs, rounds, d_err, sms);
    if (cs > 1) {
      run(src, window, 1, cs, rounds, d_err, sms);
      run(src, window, 2, cs, rounds, d_err, sms);
    }
  }
  int herr = 0;
  cudaMemcpy(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost);
  printf("mbarrier timeouts: %s\n", herr ? "YES (results invalid)" : "none");
  return herr;
}
