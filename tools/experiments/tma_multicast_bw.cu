// Experiment: what bounds the L2 -> shared-memory operand stream of the GEMM kernels -- the SM's receive path or the L2 side?
// Every CTA of a cluster needs the SAME tile stream (the A operand shared by the N tiles of one row block).
//   unicast  : every CTA bulk-copies the whole tile itself                      (what gemm_tc.cu does today across N tiles)
//   multicast: every CTA bulk-copies 1/cs of the tile and multicasts it to all  (one L2 read feeds cs SMs)
// All source data is L2-resident (each cluster cycles through a 1 MB region).  Delivered bytes per SM per clock are compared:
// equal  => the SM's receive path is the limit and multicast buys nothing; multicast higher => the L2 side is the limit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/experiments/tma_multicast_bw tools/experiments/tma_multicast_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int TILE = 16 * 1024;      // bytes per tile
constexpr int BATCH = 4;             // tiles per batch; two batches (double buffer) = 128 KB of shared memory
constexpr int REGION = 1 << 20;      // bytes of source per cluster

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t phase) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(phase) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
  const long long t0 = clock64();
  while (!mbar_try(b, phase))
    if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void bulk_uni(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ void tma_2d(void* dst, const void* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// The GEMM's A operand as it is fetched today: a {64 fp16 x 128 rows} box (16 KB, 128-byte rows 1 KB apart) of a row-major
// [rows, 512] fp16 matrix through a tiled tensor map with SWIZZLE_128B.  `share` consecutive CTAs fetch the same boxes.
__global__ void __launch_bounds__(128) bw_tensor_kernel(const __grid_constant__ CUtensorMap map, int iters, int share, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  const int stream = blockIdx.x / share;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int buf, int batch) {
    mbar_expect(&bars[buf], BATCH * TILE);
    for (int t = 0; t < BATCH; ++t) {
      const int tile = batch * BATCH + t;
      const int kb = tile & 7, rb = (tile >> 3) & 7;          // 8 k-blocks x 8 row blocks = the stream's 1 MB region
      tma_2d(smem + (size_t)(buf * BATCH + t) * TILE, &map, &bars[buf], kb * 64, stream * 1024 + rb * 128);
    }
  };
  if (threadIdx.x == 0) {
    issue(0, 0);
    issue(1, 1);
  }
  unsigned acc = 0;
  for (int b = 0; b < iters; ++b) {
    const int buf = b & 1;
    mbar_wait(&bars[buf], (b >> 1) & 1);
    acc += smem[(size_t)buf * BATCH * TILE + threadIdx.x * 64];
    __syncthreads();
    if (threadIdx.x == 0 && b + 2 < iters) issue(buf, b + 2);
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
}

// mode 0: unicast, 1: multicast, 2: unicast WITHOUT a cluster launch -- `share` consecutive CTAs of a plain grid read the same stream
// (what the GEMM's N tiles of one row block do today).  cs = cluster size (1 = no cluster).
__global__ void __launch_bounds__(128) bw_kernel(const uint8_t* __restrict__ src, int iters, int mode, int cs, unsigned* sink, int share) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t r = cs > 1 ? cluster_rank() : 0;
  const int cluster = blockIdx.x / (mode == 2 ? share : cs);
  const uint8_t* base = src + (size_t)cluster * REGION;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cs > 1) cluster_sync();
  const uint16_t mask = (uint16_t)((1u << cs) - 1);
  auto issue = [&](int buf, int batch) {   // one thread: BATCH tiles of batch `batch` into buffer `buf`
    mbar_expect(&bars[buf], BATCH * TILE);
    for (int t = 0; t < BATCH; ++t) {
      const size_t off = ((size_t)(batch * BATCH + t) * TILE) % REGION;
      uint8_t* dst = smem + (size_t)(buf * BATCH + t) * TILE;
      if (mode == 0 || cs == 1) {
        bulk_uni(dst, base + off, TILE, &bars[buf]);
      } else {
        const uint32_t chunk = TILE / cs;
        bulk_mc(dst + r * chunk, base + off + r * chunk, chunk, &bars[buf], mask);
      }
    }
  };
  if (threadIdx.x == 0) {
    issue(0, 0);
    issue(1, 1);
  }
  unsigned acc = 0;
  for (int b = 0; b < iters; ++b) {
    const int buf = b & 1;
    mbar_wait(&bars[buf], (b >> 1) & 1);
    acc += smem[(size_t)buf * BATCH * TILE + threadIdx.x * 64];   // touch the data
    __syncthreads();
    if (cs > 1 && mode == 1) cluster_sync();   // every CTA of the cluster is done with the buffer before anyone refills it
    if (threadIdx.x == 0 && b + 2 < iters) issue(buf, b + 2);
  }
  if (cs > 1) cluster_sync();
  if (acc == 0xdeadbeef) sink[0] = acc;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const size_t bytes = (size_t)sms * REGION;
  uint8_t* src;
  unsigned* sink;
  cudaMalloc(&src, bytes);
  cudaMalloc(&sink, 4);
  cudaMemset(src, 1, bytes);
  const int smem = 2 * BATCH * TILE;
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 4000;
  printf("# %d SMs, nominal %d MHz; tile %d B, %d tiles per batch, %d batches per CTA, source L2-resident (%zu MB)\n", sms, khz / 1000, TILE,
         BATCH, iters, bytes >> 20);
  printf("# cs mode     CTAs   ms      delivered GB/s (all SMs)   per SM GB/s   L2 read GB/s\n");
  for (int cs : {1, 2, 4, 8}) {
    for (int mode : {0, 1}) {
      if (cs == 1 && mode == 1) continue;
      for (int ctas : {sms / cs * cs, 128 / cs * cs, 32 / cs * cs > 0 ? 32 / cs * cs : cs}) {
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
        float best = 1e30f;
        cudaError_t e = cudaSuccess;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          e = cudaLaunchKernelEx(&cfg, bw_kernel, (const uint8_t*)src, iters, mode, cs, sink, 1);
          cudaEventRecord(e1);
          cudaError_t e2 = cudaDeviceSynchronize();
          if (e == cudaSuccess) e = e2;
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          if (rep > 0 && ms < best) best = ms;
          cudaEventDestroy(e0); cudaEventDestroy(e1);
          if (e != cudaSuccess) break;
        }
        if (e != cudaSuccess) {
          printf("%4d %-9s %4d  failed: %s\n", cs, mode ? "multicast" : "unicast", ctas, cudaGetErrorString(e));
          cudaGetLastError();
          continue;
        }
        const double delivered = (double)ctas * iters * BATCH * TILE;          // bytes landing in shared memory
        const double l2 = (mode == 1 && cs > 1) ? delivered / cs : delivered;   // bytes read from L2
        printf("%4d %-9s %4d  %7.3f  %12.0f %20.1f %14.0f\n", cs, mode ? "multicast" : "unicast", ctas, best, delivered / best / 1e6,
               delivered / best / 1e6 / ctas, l2 / best / 1e6);
      }
    }
  }
  printf("# plain grid (no cluster): `share` consecutive CTAs read the same stream\n# share          CTAs   ms      delivered GB/s (all SMs)   per SM GB/s\n");
  for (int share : {1, 2, 4, 8, 16}) {
    for (int ctas : {sms / share * share, 128}) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = nullptr; cfg.numAttrs = 0;
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, bw_kernel, (const uint8_t*)src, iters, 2, 1, sink, share);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
      }
      const double delivered = (double)ctas * iters * BATCH * TILE;
      printf("%4d %-9s %4d  %7.3f  %12.0f %20.1f\n", share, "plain", ctas, best, delivered / best / 1e6, delivered / best / 1e6 / ctas);
    }
  }
  // tensor-map (tiled, swizzled) fetches of the same bytes
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  if (fn) {
    CUtensorMap map;
    const cuuint64_t dims[2] = {512, (cuuint64_t)sms * 1024};
    const cuuint64_t strides[1] = {1024};
    const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("# tiled tensor map, box 64 fp16 x 128 rows (128-byte rows, 1 KB apart), SWIZZLE_128B; encode rc %d\n# share          CTAs   ms      delivered GB/s (all SMs)   per SM GB/s\n", (int)r);
    cudaFuncSetAttribute(bw_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + 1024);
    for (int share : {1, 2, 4, 8}) {
      for (int ctas : {sms / share * share, 128, 32}) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          bw_tensor_kernel<<<ctas, 128, smem + 1024>>>(map, iters, share, sink);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          if (rep > 0 && ms < best) best = ms;
          cudaEventDestroy(e0); cudaEventDestroy(e1);
          if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
        }
        const double delivered = (double)ctas * iters * BATCH * TILE;
        printf("%4d %-9s %4d  %7.3f  %12.0f %20.1f\n", share, "tensor", ctas, best, delivered / best / 1e6, delivered / best / 1e6 / ctas);
      }
    }
  }
  return 0;
}
