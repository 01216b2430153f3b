// Bandwidth-bound kernels of the stage-2 objective's forward half (MAGE.forward in eval mode, mage_model.py:575-639; SURVEY.md §8
// row N2): GroupNorm of the 3-D convolutional video posterior, the reparameterisation + KL term, per-row cross-entropy, and a
// fixed-order mean.  The dense work of that pass (3x3x3 convolutions with the temporal taps folded into the channel axis, the
// teacher-forced decoder) runs on the tensor-core kernels of gemm_tc.cu.  All fp32, channels-last, deterministic.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

// (mean, rstd) of every (sample, group) from the per-(slot, sample, group) partial sums of mage_gn_partial_f32, combined over the
// slots in slot order (double): one thread per (sample, group).
__global__ void gn_stats_kernel(const double* __restrict__ part, float2* __restrict__ stat, int n_slots, int B, int G, double count,
                                float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * G) return;
  double s = 0.0, q = 0.0;
  for (int sl = 0; sl < n_slots; ++sl) {
    s += part[((int64_t)sl * B * G + i) * 2];
    q += part[((int64_t)sl * B * G + i) * 2 + 1];
  }
  const double mean = s / count;
  double var = q / count - mean * mean;   // biased, like nn.GroupNorm
  if (var < 0.0) var = 0.0;
  stat[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// y = (x - mean) * rstd * gamma + beta (+ residual) (ReLU) for rows of x [n_slots*B*HW, 512]; row -> sample (row / HW) % B.
// One warp per row, lane holds 4 float4 (float4 index i*32 + lane -> channels 4*(i*32+lane) .. +3, all in one group: cpg % 4 == 0).
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const float2* __restrict__ stat,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ res, float* __restrict__ out,
                                                       __half* __restrict__ split, int64_t plane, int* flag, int rows, int B, int HW,
                                                       int G, int cpg, int relu) {
  constexpr int C = 512, NV = 4;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = (row / HW) % B;
  const float4* src = reinterpret_cast<const float4*>(x + (int64_t)row * C);
  const float4* rs = res ? reinterpret_cast<const float4*>(res + (int64_t)row * C) : nullptr;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = i * 32 + lane;
    const float2 st = __ldg(stat + b * G + (c4 * 4) / cpg);
    const float4 v = src[c4];
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 o;
    o.x = (v.x - st.x) * st.y * g.x + bt.x;
    o.y = (v.y - st.x) * st.y * g.y + bt.y;
    o.z = (v.z - st.x) * st.y * g.z + bt.z;
    o.w = (v.w - st.x) * st.y * g.w + bt.w;
    if (rs) {
      const float4 r = rs[c4];
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    if (out) reinterpret_cast<float4*>(out + (int64_t)row * C)[c4] = o;
    if (split) {
      uint2 hi, lo;
      bad |= tc::split4(o, hi, lo);
      const int64_t e = (int64_t)row * C + c4 * 4;
      *reinterpret_cast<uint2*>(split + e) = hi;
      *reinterpret_cast<uint2*>(split + plane + e) = lo;
    }
  }
  if (bad && flag) atomicOr(flag, 1);
}

// loss[row] = logsumexp(logits[row, :]) - logits[row, target[row]]  (F.cross_entropy, reduction='none'): one warp per row.
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ target,
                                                      float* __restrict__ loss, int rows, int K, int* flag) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* r = logits + (int64_t)row * ld;
  float m = -INFINITY;
  for (int k = lane; k < K; k += 32) m = fmaxf(m, r[k]);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += expf(r[k] - m);
  s = warp_sum(s);
  if (lane == 0) {
    const int64_t t = target[row];
    if (t < 0 || t >= K) {   // F.cross_entropy raises on a class index outside [0, K): flagged, never dereferenced
      if (flag) atomicOr(flag, 4);
      loss[row] = 0.f;
    } else {
      loss[row] = (m + logf(s)) - r[t];
    }
  }
}

// reparameterize (mage_model.py:569-573) with the draw given, and the KL integrand (:624-625):
//   z[b, c, p] (NCHW, like the sampling path's noise) = eps[b, c, p] * exp(0.5 * logvar[b, p, c]) + mu[b, p, c]
//   kl_rows[b] = sum_{c,p} 1 + logvar - mu^2 - exp(logvar)
// mu / logvar are the two column halves of one convolution output ml [B, HW, 2*Cz].  One block per sample, fixed-order tree.
__global__ void __launch_bounds__(256) reparam_kl_kernel(const float* __restrict__ ml, const float* __restrict__ eps, float* __restrict__ z,
                                                         float* __restrict__ kl_rows, int HW, int Cz) {
  __shared__ double red[256];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < HW * Cz; i += 256) {
    const int p = i / Cz, c = i - p * Cz;
    const float mu = ml[((int64_t)b * HW + p) * 2 * Cz + c];
    const float lv = ml[((int64_t)b * HW + p) * 2 * Cz + Cz + c];
    if (z) {
      const int64_t o = ((int64_t)b * Cz + c) * HW + p;
      z[o] = eps[o] * expf(0.5f * lv) + mu;
    }
    acc += (double)(1.f + lv - mu * mu - expf(lv));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) kl_rows[b] = (float)red[0];
}

// out[0] = scale * sum(x[0..n)): one block, double accumulation, fixed order (thread-strided partials, then a tree).
__global__ void __launch_bounds__(1024) scaled_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, double scale) {
  __shared__ double red[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)x[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * scale);
}

// out[0] = scale * sum((a[i] - b[i])^2): F.mse_loss's reduction, one block, double accumulation, fixed order.
__global__ void __launch_bounds__(1024) scaled_sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                 float* __restrict__ out, int64_t n, double scale) {
  __shared__ double red[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float d = a[i] - b[i];
    acc += (double)(d * d);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * scale);
}

}  // namespace

extern "C" int mage_gn_apply_f32(mage_ctx* ctx, const float* x, const double* part, float* stat, const float* gamma, const float* beta,
                                 const float* residual, float* out, void* out_split, int64_t split_plane, int* flag, int n_slots, int B,
                                 int HW, int C, int groups, int relu, float eps, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_slots > 0 && B > 0 && HW > 0 && C == 512 && groups > 0 && C % groups == 0 && (C / groups) % 4 == 0 && stat);
  MAGE_CHECK_ARG(aligned16(x) && aligned16(gamma) && aligned16(beta) && (!residual || aligned16(residual)) && (!out || aligned16(out)) &&
                 (out || out_split) && (reinterpret_cast<uintptr_t>(out_split) & 7) == 0 && split_plane % 4 == 0 &&
                 (reinterpret_cast<uintptr_t>(stat) & 7) == 0);
  const int64_t rows = (int64_t)n_slots * B * HW;
  MAGE_CHECK_ARG(rows < ((int64_t)1 << 31));
  const double count = (double)n_slots * HW * (C / groups);
  gn_stats_kernel<<<(B * groups + 255) / 256, 256, 0, as_stream(stream)>>>(part, reinterpret_cast<float2*>(stat), n_slots, B, groups, count, eps);
  if (int r = mage_post_launch(ctx)) return r;
  gn_apply_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<const float2*>(stat), gamma, beta, residual, out,
                                                                            reinterpret_cast<__half*>(out_split), split_plane, flag,
                                                                            (int)rows, B, HW, groups, C / groups, relu);
  return mage_post_launch(ctx);
}

extern "C" int mage_cross_entropy_rows_f32(mage_ctx* ctx, const float* logits, int64_t ld, const int64_t* target, float* loss, int rows,
                                           int K, int* flag, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(rows > 0 && K > 0 && ld >= K && logits && target && loss);
  ce_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(logits, ld, target, loss, rows, K, flag);
  return mage_post_launch(ctx);
}

extern "C" int mage_reparam_kl_f32(mage_ctx* ctx, const float* mu_logvar, const float* eps, float* z, float* kl_rows, int B, int HW,
                                   int Cz, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(B > 0 && HW > 0 && Cz > 0 && mu_logvar && kl_rows && (z == nullptr || eps != nullptr));
  reparam_kl_kernel<<<B, 256, 0, as_stream(stream)>>>(mu_logvar, eps, z, kl_rows, HW, Cz);
  return mage_post_launch(ctx);
}

extern "C" int mage_scaled_sqdiff_sum_f32(mage_ctx* ctx, const float* a, const float* b, float* out, int64_t n, double scale, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n > 0 && a && b && out);
  scaled_sqdiff_sum_kernel<<<1, 1024, 0, as_stream(stream)>>>(a, b, out, n, scale);
  return mage_post_launch(ctx);
}

extern "C" int mage_scaled_sum_f32(mage_ctx* ctx, const float* x, float* out, int64_t n, double scale, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n > 0 && x && out);
  scaled_sum_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, out, n, scale);
  return mage_post_launch(ctx);
}
