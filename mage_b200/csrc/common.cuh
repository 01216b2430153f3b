// Shared helpers for libmage_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mage_b200.h"

// The opaque handle of the C ABI (include/mage_b200.h: mage_ctx_create / mage_ctx_destroy): everything the library remembers
// between calls lives here -- the device it was created for, that device's SM count and per-kernel launch configuration
// (opt-in shared-memory size, co-resident CTA-pair count), the tile-selection / launch tuning switches, the launch counter.
// No process-global mutable state: two handles (two devices, two threads) never see each other.
struct mage_ctx {
  int device = 0;
  int sms = 148;
  int forced_bn = 0, forced_pair = -1;   // mage_tc_tuning
  int ns = 1;                            // mage_tc_nsplit
  int halo = 1;                          // mage_tc_conv_halo
  int small = 1;                         // one-tile cost model for sub-2-wave GEMMs (MAGE_TC_SMALL)
  int resident = 1;                      // resident weight slots in the halo convolution when they fit (MAGE_TC_RESIDENT)
  int pdl = 1;                           // mage_pdl
  int tattn_ring = 0;                    // temporal attention step as the persistent ring kernel (MAGE_TATTN_RING; 0 = one CTA per unit)
  int sm_share = 0;                      // mage_sm_share: persistent kernels use at most this many SMs (0 = all of them)
  int eff_sms() const { return sm_share > 0 && sm_share < sms ? sm_share : sms; }
  int64_t launches = 0;
  static constexpr int kMaxKernels = 64;
  const void* cfg_fn[kMaxKernels] = {};  // kernels whose attributes have been set on `device`
  int cfg_units[kMaxKernels] = {};       // ... and how many CTAs / CTA pairs of each can be co-resident
  size_t cfg_smem[kMaxKernels] = {};
  int n_cfg = 0;
  int find(const void* fn) const {
    for (int i = 0; i < n_cfg; ++i)
      if (cfg_fn[i] == fn) return i;
    return -1;
  }
  int add(const void* fn, int units, size_t smem) {
    if (n_cfg >= kMaxKernels) return -1;
    cfg_fn[n_cfg] = fn; cfg_units[n_cfg] = units; cfg_smem[n_cfg] = smem;
    return n_cfg++;
  }
};

#define MAGE_CHECK_ARG(cond) \
  do {                       \
    if (!(cond)) return MAGE_EINVAL; \
  } while (0)
#define MAGE_CHECK_CTX(ctx) \
  do {                      \
    if ((ctx) == nullptr) return MAGE_EINVAL; \
  } while (0)

// Count the launch and surface launch-time errors (never sync here: callers own the stream).
static inline int mage_post_launch(mage_ctx* ctx) {
  ++ctx->launches;
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
// attribute.  On by default (mage_ctx::pdl); MAGE_PDL=0 launches everything with plain stream order.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t mage_launch_pdl(const mage_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (ctx->pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
