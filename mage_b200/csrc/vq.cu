// VectorQuantizer nearest-code search (vqvae_model.py:8-25) as one fused kernel:
// distance GEMM tile in registers + running (min, argmin) + half-warp shuffle reduction.
//   d[n,k] = (|c_k|^2 + |z_n|^2) - 2 z_n.c_k      (same association as the reference's addmm)
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int D) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(x + (int64_t)row * D);
  float s = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = __ldg(p + i);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

constexpr int VBM = 64, VBN = 128, VBK = 16;

__global__ void __launch_bounds__(256) vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                        const float* __restrict__ csq, int64_t* __restrict__ idx,
                                                        int N, int D, int K) {
  __shared__ __align__(16) float As[2][VBK][VBM + 4];
  __shared__ __align__(16) float Bs[2][VBK][VBN + 4];
  __shared__ float zsq[VBM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * VBM;

  // |z|^2 of the CTA's 64 rows: 8 warps x 8 rows, coalesced float4 reads
  for (int r = warp * 8; r < warp * 8 + 8; ++r) {
    float s = 0.f;
    if (m0 + r < N) {
      const float4* p = reinterpret_cast<const float4*>(z + (int64_t)(m0 + r) * D);
      for (int i = lane; i < D / 4; i += 32) {
        const float4 v = __ldg(p + i);
        s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    }
    s = warp_sum(s);
    if (lane == 0) zsq[r] = s;
  }

  const int a_row = tid >> 2, kv = (tid & 3) * 4;
  const bool a_ok = (m0 + a_row) < N;
  const float* a_ptr = z + (int64_t)(m0 + a_row) * D;

  float best[4];
  int bidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = INFINITY; bidx[i] = 0x7fffffff; }

  const int nk = D / VBK;
  for (int ct = 0; ct < K / VBN; ++ct) {
    const float* b_ptr0 = cb + (int64_t)(ct * VBN + a_row) * D;
    const float* b_ptr1 = cb + (int64_t)(ct * VBN + 64 + a_row) * D;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float4 ra, rb0, rb1;
    auto load_global = [&](int kb) {
      const int k0 = kb * VBK + kv;
      ra = a_ok ? __ldg(reinterpret_cast<const float4*>(a_ptr + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb0 = __ldg(reinterpret_cast<const float4*>(b_ptr0 + k0));
      rb1 = __ldg(reinterpret_cast<const float4*>(b_ptr1 + k0));
    };
    auto store_smem = [&](int buf) {
      As[buf][kv + 0][a_row] = ra.x; As[buf][kv + 1][a_row] = ra.y; As[buf][kv + 2][a_row] = ra.z; As[buf][kv + 3][a_row] = ra.w;
      Bs[buf][kv + 0][a_row] = rb0.x; Bs[buf][kv + 1][a_row] = rb0.y; Bs[buf][kv + 2][a_row] = rb0.z; Bs[buf][kv + 3][a_row] = rb0.w;
      Bs[buf][kv + 0][64 + a_row] = rb1.x; Bs[buf][kv + 1][64 + a_row] = rb1.y;
      Bs[buf][kv + 2][64 + a_row] = rb1.z; Bs[buf][kv + 3][64 + a_row] = rb1.w;
    };
    load_global(0);
    store_smem(0);
    __syncthreads();
    int cur = 0;
    for (int kb = 0; kb < nk; ++kb) {
      if (kb + 1 < nk) load_global(kb + 1);
#pragma unroll
      for (int k = 0; k < VBK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
        const float a[4] = {av.x, av.y, av.z, av.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (kb + 1 < nk) store_smem(cur ^ 1);
      __syncthreads();
      cur ^= 1;
    }
    // distances of this 64x128 tile -> running minima (codes visited in increasing index per thread)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float zs = zsq[ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = ct * VBN + (j >> 2) * 64 + tx * 4 + (j & 3);
        const float d = fmaf(-2.f, acc[i][j], __ldg(csq + c) + zs);
        if (d < best[i] || (d == best[i] && c < bidx[i])) { best[i] = d; bidx[i] = c; }
      }
    }
  }
  // reduce over the 16 threads (tx) that share a row: they are one half-warp
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best[i], o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], o);
      if (ov < best[i] || (ov == best[i] && oi < bidx[i])) { best[i] = ov; bidx[i] = oi; }
    }
    const int m = m0 + ty * 4 + i;
    if (tx == 0 && m < N) idx[m] = bidx[i];
  }
}

}  // namespace

extern "C" int mage_vq_argmin_f32(mage_ctx* ctx, const float* z, const float* codebook, float* csq_scratch, int64_t* idx, int N, int D, int K,
                                  void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(N > 0 && D > 0 && D % VBK == 0 && K > 0 && K % VBN == 0 && aligned16(z) && aligned16(codebook) && csq_scratch);
  cudaStream_t st = as_stream(stream);
  row_sqnorm_kernel<<<(K + 7) / 8, 256, 0, st>>>(codebook, csq_scratch, K, D);
  int e = mage_post_launch(ctx);
  if (e) return e;
  vq_argmin_kernel<<<(N + VBM - 1) / VBM, 256, 0, st>>>(z, codebook, csq_scratch, idx, N, D, K);
  return mage_post_launch(ctx);
}
