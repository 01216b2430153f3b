// mage_gemm_f32: picks the GEMM back end.
#include "common.cuh"

int mage_gemm_simt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* residual,
                   int64_t ldr, int res_mod, float* C, int64_t ldc, int M, int N, int K, int act, int relu_a, cudaStream_t st);
int mage_gemm_tc(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* residual,
                 int64_t ldr, int res_mod, float* C, int64_t ldc, int M, int N, int K, int act, int relu_a, cudaStream_t st);

static int g_backend = MAGE_GEMM_SIMT;

extern "C" int mage_set_gemm_backend(int backend) {
  if (backend != MAGE_GEMM_SIMT && backend != MAGE_GEMM_TCGEN05) return MAGE_EINVAL;
  g_backend = backend;
  return 0;
}
extern "C" int mage_get_gemm_backend(void) { return g_backend; }

extern "C" int mage_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                             const float* residual, int64_t ldr, int res_mod, float* C, int64_t ldc, int M, int N, int K,
                             int act, int relu_a, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (g_backend == MAGE_GEMM_TCGEN05) {
    const int r = mage_gemm_tc(A, lda, W, ldw, bias, residual, ldr, res_mod, C, ldc, M, N, K, act, relu_a, st);
    if (r != MAGE_ENOTSUP) return r;  // shapes the tensor-core kernel does not take fall through to FFMA
  }
  return mage_gemm_simt(A, lda, W, ldw, bias, residual, ldr, res_mod, C, ldc, M, N, K, act, relu_a, st);
}
