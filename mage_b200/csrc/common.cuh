// Shared helpers for libmage_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mage_b200.h"

extern int64_t g_mage_launches;  // defined in misc.cu

#define MAGE_CHECK_ARG(cond) \
  do {                       \
    if (!(cond)) return MAGE_EINVAL; \
  } while (0)

// Count the launch and surface launch-time errors (never sync here: callers own the stream).
static inline int mage_post_launch() {
  ++g_mage_launches;
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

__device__ __forceinline__ float mage_act(float x, int act) {
  switch (act) {
    case MAGE_ACT_RELU: return fmaxf(x, 0.f);
    case MAGE_ACT_QUICKGELU: return x * (1.f / (1.f + expf(-1.702f * x)));
    case MAGE_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case MAGE_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
