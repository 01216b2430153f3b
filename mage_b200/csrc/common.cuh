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

// Programmatic dependent launch (PDL).  Kernels of the per-step path call pdl_launch_dependents() first thing -- the next
// kernel's CTAs may then be scheduled as soon as SM resources free up, i.e. its prologue (barrier init, TMEM allocation,
// descriptor prefetch) and its launch latency overlap this kernel's tail -- and pdl_wait() before their first global-memory
// access, which blocks until the preceding grid has completed and flushed.  Both are no-ops for a kernel launched without the
// attribute.  MAGE_PDL=0 launches everything with plain stream order.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern int g_mage_pdl;  // defined in misc.cu

template <typename... KArgs, typename... Args>
static inline cudaError_t mage_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                          Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (g_mage_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
