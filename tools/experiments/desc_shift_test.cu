// Experiment: can a tcgen05 shared-memory matrix descriptor (K-major, SWIZZLE_128B) start at an arbitrary 128-byte ROW of a
// TMA-written tile (start address not 1024-aligned)?  That is what a halo-reusing 3x3 convolution needs: one NHWC halo patch
// in shared memory, nine A operands that are the same patch shifted by (dy*pitch + dx) pixel rows.
// For every shift 0..17 and both settings of the descriptor's "matrix base offset" field (0, and (addr >> 7) & 7) the kernel
// computes D = A[shift : shift+128, :] . B^T and the host compares with a CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I mage_b200/csrc -o tools/experiments/desc_shift_test tools/experiments/desc_shift_test.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "tc_common.cuh"
using namespace tc;

constexpr int ROWS = 192, BOXR = 184, K = 64, N = 64, NSHIFT = 18, NVAR = 3;

__device__ __forceinline__ uint64_t desc_with_base(uint32_t addr, uint32_t base_off, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 24 * 1024, bar = base + 34 * 1024, bar2 = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, BOXR * K * 2 + N * K * 2);
    tma_load_3d(sA, &mapA, bar, 0, 0, 0);
    tma_load_3d(sB, &mapB, bar, 0, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t phase = 0;
  for (int variant = 0; variant < NVAR; ++variant) {
    for (int shift = 0; shift < NSHIFT; ++shift) {
      if (threadIdx.x == 0) {
        const uint32_t a_addr = sA + shift * 128;
        const uint32_t bo = variant == 1 ? ((a_addr >> 7) & 7) : 0;
        const uint32_t sbo = variant == 2 ? 1280 : 1024;   // variant 2: 8-row groups 10 rows apart (a 16x8-pixel tile of a 10-pixel-pitch halo)
        constexpr uint32_t idesc = umma_idesc_f16(N, 128);
        for (int kk = 0; kk < K / 16; ++kk) {
          const uint64_t ad = desc_with_base(a_addr + kk * 32, bo, sbo);
          const uint64_t bd = desc_with_base(sB + kk * 32, 0, 1024);
          umma_f16(tmem, ad, bd, idesc, kk > 0);
        }
        umma_commit(bar2);
      }
      mbar_wait(bar2, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t r[32];
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        tmem_wait_ld();
        float* dst = out + (((size_t)variant * NSHIFT + shift) * 128 + warp * 32 + lane) * N + c * 32;
        for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  std::vector<__half> hA(ROWS * K), hB(N * K);
  std::vector<float> fA(ROWS * K), fB(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, sizeof(float) * NVAR * NSHIFT * 128 * N);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0, sizeof(float) * NVAR * NSHIFT * 128 * N);
  CUtensorMap mA, mB;
  cuuint32_t es[3] = {1, 1, 1};
  {
    cuuint64_t dims[3] = {K, ROWS, 1}; cuuint64_t str[2] = {K * 2, (cuuint64_t)K * 2 * ROWS}; cuuint32_t box[3] = {K, BOXR, 1};
    CUresult r = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[3] = {K, N, 1}; cuuint64_t str[2] = {K * 2, (cuuint64_t)K * 2 * N}; cuuint32_t box[3] = {K, N, 1};
    CUresult r = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 44 * 1024);
  k<<<1, 128, 44 * 1024>>>(mA, mB, dOut);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> out(NVAR * NSHIFT * 128 * N);
  cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
  for (int variant = 0; variant < NVAR; ++variant)
    for (int shift = 0; shift < NSHIFT; ++shift) {
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          const int row = (variant == 2 ? (m / 8) * 10 + m % 8 : m) + shift;
          for (int kk = 0; kk < K; ++kk) s += (double)fA[row * K + kk] * fB[n * K + kk];
          maxerr = fmax(maxerr, fabs(s - out[((size_t)(variant * NSHIFT + shift) * 128 + m) * N + n]));
        }
      printf("variant=%s shift=%2d  max|err| = %.3e  %s\n", variant == 1 ? "base_offset=(addr>>7)&7" : variant == 2 ? "base_offset=0,SBO=1280" : "base_offset=0", shift, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
