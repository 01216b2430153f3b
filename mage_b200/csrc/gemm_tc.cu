// tcgen05 GEMM back end -- placeholder until the TMA/TMEM kernel lands (returns ENOTSUP so the
// dispatcher uses the FFMA kernel).
#include "common.cuh"

int mage_gemm_tc(const float*, int64_t, const float*, int64_t, const float*, const float*, int64_t, int, float*, int64_t,
                 int, int, int, int, int, cudaStream_t) {
  return MAGE_ENOTSUP;
}
