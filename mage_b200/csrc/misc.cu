// Bandwidth-bound helper kernels of the MAGE sampling path: LayerNorm, greedy argmax, embedding
// gathers, pooling, AdaIN, first/last VQ-VAE layers.  All fp32, channels-last, float4 accesses.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

extern "C" int mage_abi_version(void) { return 5; }

// One handle per (process, device).  Environment variables only seed a new handle's switches: MAGE_PDL (programmatic dependent
// launch of the per-step kernels: ON -- same-box A/B, profiles/r02z_pdl_ab.txt: 22.29 -> 21.56 ms per generate at 8 prompts, 36.3 ->
// 35.6 at 16, neutral at 64), MAGE_TC_BN / MAGE_TC_PAIR / MAGE_TC_NS / MAGE_TC_HALO / MAGE_TC_SMALL / MAGE_TC_RESIDENT (tile
// selection, see gemm_tc.cu).
extern "C" int mage_ctx_create(int device, mage_ctx** out) {
  if (!out) return MAGE_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return (int)e;
  if (device < 0 || device >= n) return MAGE_EINVAL;
  int major = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (major != 10) return MAGE_ENOTSUP;   // sm_100a code only: no fallback for other architectures
  mage_ctx* c = new mage_ctx();
  c->device = device;
  c->sms = sms > 0 ? sms : 148;
  auto env = [](const char* k, int d) { const char* v = getenv(k); return v ? atoi(v) : d; };
  c->pdl = env("MAGE_PDL", 1);
  c->forced_bn = env("MAGE_TC_BN", 0);
  c->forced_pair = env("MAGE_TC_PAIR", -1);
  c->ns = env("MAGE_TC_NS", 1);
  c->halo = env("MAGE_TC_HALO", 1);
  c->small = env("MAGE_TC_SMALL", 1);
  c->resident = env("MAGE_TC_RESIDENT", 1);
  c->tattn_ring = env("MAGE_TATTN_RING", 0);
  *out = c;
  return 0;
}
extern "C" int mage_ctx_destroy(mage_ctx* ctx) {
  delete ctx;
  return 0;
}
extern "C" int mage_ctx_device(mage_ctx* ctx) { return ctx ? ctx->device : MAGE_EINVAL; }
extern "C" int64_t mage_launch_count(mage_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int mage_temporal_attn_ring(mage_ctx* ctx, int enable) {
  MAGE_CHECK_CTX(ctx);
  ctx->tattn_ring = enable != 0;
  return 0;
}
extern "C" int mage_sm_share(mage_ctx* ctx, int sms) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(sms >= 0);
  ctx->sm_share = sms;
  return 0;
}
extern "C" int mage_pdl(mage_ctx* ctx, int enable) {
  MAGE_CHECK_CTX(ctx);
  ctx->pdl = enable != 0;
  return 0;
}

namespace {

// ------------------------------------------------------------------ LayerNorm
// one warp per row; the row lives in registers (NV float4 per lane), two-pass statistics.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ out,
                                                        __half* __restrict__ split, int64_t plane, int* flag,
                                                        int rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C = NV * 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(in + (int64_t)row * C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = src[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
  float4* dst = reinterpret_cast<float4*>(out + (int64_t)row * C);
  bool bad = false;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out) dst[i * 32 + lane] = o;
    if (split) {  // the next tensor-core GEMM's operand format
      uint2 hi, lo;
      bad |= tc::split4(o, hi, lo);
      const int64_t e = (int64_t)row * C + (i * 32 + lane) * 4;
      *reinterpret_cast<uint2*>(split + e) = hi;
      *reinterpret_cast<uint2*>(split + plane + e) = lo;
    }
  }
  if (bad && flag) atomicOr(flag, 1);
}

// ------------------------------------------------------------------ argmax over rows
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, int64_t ldx,
                                                          int64_t* __restrict__ idx, int rows, int N) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* r = x + (int64_t)row * ldx;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int n = lane; n < N; n += 32) {
    const float v = r[n];
    if (v > best || (v == best && n < bi) || bi == 0x7fffffff) { best = v; bi = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) idx[row] = bi;
}

// ------------------------------------------------------------------ embedding gather
__global__ void __launch_bounds__(256) embedding_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table,
                                                        float* __restrict__ out, int rows, int C4) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)rows * C4) return;
  const int row = (int)(t / C4), c = (int)(t - (int64_t)row * C4);
  reinterpret_cast<float4*>(out)[t] = __ldg(reinterpret_cast<const float4*>(table) + idx[row] * C4 + c);
}

// ------------------------------------------------------------------ convolution of a codebook-embedded token map as table lookups
// out[b,y,x,:] = sum_{ky,kx} table[ky*KW+kx][tok[b, y+ky-KH/2, x+kx-KW/2]][:] + pos_bias[y*R+x][:] + bias[:]
// (taps outside the map contribute nothing = zero padding).  One warp per pixel, NV float4 per lane, fixed tap order.
template <int NV>
__global__ void __launch_bounds__(256) token_taps_kernel(const int64_t* __restrict__ tok, const float* __restrict__ table,
                                                         const float* __restrict__ pos_bias, const float* __restrict__ bias,
                                                         float* __restrict__ out, int n_pix, int R, int K, int KH, int KW,
                                                         const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta,
                                                         float ln_eps, __half* __restrict__ ln_split, int64_t ln_plane, int* flag) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C4 = NV * 32;
  const int pix = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= n_pix) return;
  const int b = pix / (R * R), p = pix - b * R * R, y = p / R, x = p - y * R;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 pb = __ldg(reinterpret_cast<const float4*>(pos_bias) + (int64_t)p * C4 + i * 32 + lane);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + i * 32 + lane);
    acc[i] = make_float4(pb.x + bb.x, pb.y + bb.y, pb.z + bb.z, pb.w + bb.w);
  }
  for (int ky = 0; ky < KH; ++ky) {
    const int yy = y + ky - KH / 2;
    if (yy < 0 || yy >= R) continue;
    for (int kx = 0; kx < KW; ++kx) {
      const int xx = x + kx - KW / 2;
      if (xx < 0 || xx >= R) continue;
      const int64_t code = tok[(int64_t)b * R * R + yy * R + xx];
      const float4* row = reinterpret_cast<const float4*>(table) + ((int64_t)(ky * KW + kx) * K + code) * C4;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 v = __ldg(row + i * 32 + lane);
        acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(out) + (int64_t)pix * C4;
#pragma unroll
  for (int i = 0; i < NV; ++i) dst[i * 32 + lane] = acc[i];
  if (ln_split) {
    // the row is complete in this warp's registers: the first block's ln_1 (mage_model.py:49) right here, same math as
    // layernorm_kernel (two-pass statistics), emitted as the split operand of the QKV projection -- one launch less per step
    constexpr int C = NV * 128;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + ln_eps);
    bool bad = false;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(ln_gamma) + i * 32 + lane);
      const float4 b = __ldg(reinterpret_cast<const float4*>(ln_beta) + i * 32 + lane);
      float4 o;
      o.x = (acc[i].x - mean) * rstd * g.x + b.x;
      o.y = (acc[i].y - mean) * rstd * g.y + b.y;
      o.z = (acc[i].z - mean) * rstd * g.z + b.z;
      o.w = (acc[i].w - mean) * rstd * g.w + b.w;
      uint2 hi, lo;
      bad |= tc::split4(o, hi, lo);
      const int64_t e = (int64_t)pix * C + (i * 32 + lane) * 4;
      *reinterpret_cast<uint2*>(ln_split + e) = hi;
      *reinterpret_cast<uint2*>(ln_split + ln_plane + e) = lo;
    }
    if (bad && flag) atomicOr(flag, 1);
  }
}

// ------------------------------------------------------------------ 2x2 max pool NHWC
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                      int n_img, int Hin, int Win, int C4) {
  const int Ho = Hin >> 1, Wo = Win >> 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_img * Ho * Wo * C4) return;
  const int c = (int)(t % C4);
  int64_t r = t / C4;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int n = (int)(r / Ho);
  const float4* src = reinterpret_cast<const float4*>(in) + (((int64_t)n * Hin + oy * 2) * Win + ox * 2) * C4 + c;
  const float4 a = __ldg(src), b = __ldg(src + C4), d = __ldg(src + (int64_t)Win * C4), e = __ldg(src + (int64_t)Win * C4 + C4);
  float4 o;
  o.x = fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x));
  o.y = fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y));
  o.z = fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z));
  o.w = fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w));
  reinterpret_cast<float4*>(out)[t] = o;
}

// ------------------------------------------------------------------ text-encoder front end
// one warp per (b,t): gather + add + LayerNorm(eps) + pad masking; warp 0 of each b also counts tokens.
__global__ void __launch_bounds__(256) text_embed_kernel(const int64_t* __restrict__ text, const float* __restrict__ tok_emb,
                                                         const float* __restrict__ pos_emb, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ x,
                                                         int32_t* __restrict__ key_len, int B, int T, int pad_idx, float eps,
                                                         int vocab, int* flag) {
  constexpr int NV = 4, C = 512;
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * T) return;
  const int b = w / T, t = w - b * T;
  int64_t tok = text[w];
  if (tok < 0 || tok >= vocab) {
    // nn.Embedding raises on an id outside the table (mage_model.py:228); never read out of bounds: flag it (the host raises
    // IndexError after the call) and embed the padding row instead
    if (lane == 0 && flag) atomicOr(flag, 2);
    tok = pad_idx;
  }
  if (t == 0) {
    int cnt = 0;
    for (int i = lane; i < T; i += 32) cnt += (text[b * T + i] != pad_idx) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) key_len[b] = cnt;
  }
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(tok_emb + tok * C) + i * 32 + lane);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos_emb + (int64_t)t * C) + i * 32 + lane);
    v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + bb * bb) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
  const float keep = (tok != pad_idx) ? 1.f : 0.f;
  float4* dst = reinterpret_cast<float4*>(x + (int64_t)w * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    float4 o;
    o.x = ((v[i].x - mean) * rstd * g.x + be.x) * keep;
    o.y = ((v[i].y - mean) * rstd * g.y + be.y) * keep;
    o.z = ((v[i].z - mean) * rstd * g.z + be.z) * keep;
    o.w = ((v[i].w - mean) * rstd * g.w + be.w) * keep;
    dst[i * 32 + lane] = o;
  }
}

// ------------------------------------------------------------------ AdaIN (instance norm + modulation)
// block = 32 channels x 8 position groups; statistics over the HW positions of one (image, channel).
__global__ void __launch_bounds__(256) adain_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ out,
                                                    int HW, int C, float eps) {
  __shared__ float red[8][33];
  const int n = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int g = threadIdx.x >> 5;
  const float* xb = x + (int64_t)n * HW * C + c;
  float s = 0.f;
  for (int p = g; p < HW; p += 8) s += xb[(int64_t)p * C];
  red[g][threadIdx.x & 31] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x & 31];
  const float mean = tot / HW;
  __syncthreads();
  float q = 0.f;
  for (int p = g; p < HW; p += 8) {
    const float d = xb[(int64_t)p * C] - mean;
    q += d * d;
  }
  red[g][threadIdx.x & 31] = q;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x & 31];
  const float rstd = rsqrtf(tot / HW + eps);
  const float* gb = gamma + (int64_t)n * HW * C + c;
  const float* bb = beta + (int64_t)n * HW * C + c;
  float* ob = out + (int64_t)n * HW * C + c;
  for (int p = g; p < HW; p += 8) {
    const int64_t o = (int64_t)p * C;
    ob[o] = gb[o] * ((xb[o] - mean) * rstd) + bb[o];
  }
}

__global__ void __launch_bounds__(256) add_scaled_vec_kernel(float* __restrict__ x, const float* __restrict__ s,
                                                             const float* __restrict__ vec, int HW, int C, int64_t total) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const int n = (int)(t / ((int64_t)HW * C));
  x[t] = x[t] + __fmul_rn(s[n], vec[c]);
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int C, int HW, int64_t total) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const int64_t r = t / C;
  const int p = (int)(r % HW);
  const int64_t n = r / HW;
  out[t] = in[(n * C + c) * HW + p];
}


// ------------------------------------------------------------------ MAGE+ continuous head: GroupNorm(32) -> SiLU -> 1x1x1 conv
// (mage_model.py:349-354, :386-388).  The GroupNorm statistics of a sample span its 16 channels x ALL temporal slots x H x W,
// so they are kept as per-(slot, sample, group) partial sums in double (a slot's partials are recomputed only when its hidden
// state changes) and combined over the slots by the consumer.
// gn_partial: one block per (slot, sample, group): 256 threads = the HW positions (strided if HW != 256), 16 channels each.
__global__ void __launch_bounds__(256) gn_partial_kernel(const float* __restrict__ x, double* __restrict__ part, int B, int HW,
                                                         int C, int cpg) {
  __shared__ double red[2][256];
  const int g = blockIdx.x % (C / cpg);
  const int sb = blockIdx.x / (C / cpg);               // slot * B + sample
  const float* base = x + (int64_t)sb * HW * C + g * cpg;
  double s = 0.0, q = 0.0;
  for (int p = threadIdx.x; p < HW; p += 256) {
    const float* r = base + (int64_t)p * C;
    for (int c = 0; c < cpg; c += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(r + c));
      s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {   // fixed-order tree: deterministic
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[(int64_t)blockIdx.x * 2] = red[0][0];
    part[(int64_t)blockIdx.x * 2 + 1] = red[1][0];
  }
}

// gn_head: one warp per row (slot, sample, position) of `x` [rows, 512]; lane = group (32 groups of 16 channels).
//   out[row, co] = bias[co] + sum_c w[co, c] * silu((x[row, c] - mean_g) * rstd_g * gamma[c] + beta[c]),   co < cout <= 8
// mean/rstd of (sample, group) from the partial sums of ALL n_slots slots (part [n_slots, B, 32, 2]).
__global__ void __launch_bounds__(256) gn_head_kernel(const float* __restrict__ x, const double* __restrict__ part,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      float* __restrict__ out, int rows, int B, int HW, int n_slots, int cout,
                                                      float eps) {
  constexpr int C = 512, CPG = 16;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = (row / HW) % B;
  double s = 0.0, q = 0.0;
  for (int sl = 0; sl < n_slots; ++sl) {
    const double* pp = part + (((int64_t)sl * B + b) * 32 + lane) * 2;
    s += pp[0];
    q += pp[1];
  }
  const double n = (double)n_slots * HW * CPG;
  const double mean_d = s / n;
  const float mean = (float)mean_d;
  const float rstd = rsqrtf((float)(q / n - mean_d * mean_d) + eps);
  float y[CPG];
  const float* xr = x + (int64_t)row * C + lane * CPG;
#pragma unroll
  for (int c = 0; c < CPG; c += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + lane * CPG + c));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + lane * CPG + c));
    const float t0 = (v.x - mean) * rstd * gm.x + bt.x, t1 = (v.y - mean) * rstd * gm.y + bt.y;
    const float t2 = (v.z - mean) * rstd * gm.z + bt.z, t3 = (v.w - mean) * rstd * gm.w + bt.w;
    y[c] = t0 / (1.f + expf(-t0)); y[c + 1] = t1 / (1.f + expf(-t1));
    y[c + 2] = t2 / (1.f + expf(-t2)); y[c + 3] = t3 / (1.f + expf(-t3));
  }
  for (int co = 0; co < cout; ++co) {
    const float* wr = w + (int64_t)co * C + lane * CPG;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) acc = fmaf(y[c], __ldg(wr + c), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)row * cout + co] = acc + __ldg(bias + co);
  }
}

// ------------------------------------------------------------------ first conv layer (tiny Cin, planar input)
// block: one output row segment of PX pixels x all Cout channels; the input patch sits in shared memory,
// each thread owns one output channel and PX accumulators; weights stream through L1 (transposed [K][Cout]).
template <int PX>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ in, const float* __restrict__ w_t,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int Cin, int H, int W, int Hout, int Wout, int Cout,
                                                         int KH, int KW, int stride, int pad, int act) {
  extern __shared__ float patch[];  // [Cin][KH][PW]
  const int PW = (PX - 1) * stride + KW;
  const int segs = (Wout + PX - 1) / PX;
  int b = blockIdx.x;
  const int seg = b % segs; b /= segs;
  const int oy = b % Hout;
  const int n = b / Hout;
  const int ox0 = seg * PX;
  const int iy0 = oy * stride - pad, ix0 = ox0 * stride - pad;
  for (int i = threadIdx.x; i < Cin * KH * PW; i += blockDim.x) {
    const int px = i % PW;
    const int ky = (i / PW) % KH;
    const int c = i / (PW * KH);
    const int iy = iy0 + ky, ix = ix0 + px;
    patch[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? in[(((int64_t)n * Cin + c) * H + iy) * W + ix] : 0.f;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float acc[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[p] = 0.f;
    for (int c = 0; c < Cin; ++c)
      for (int ky = 0; ky < KH; ++ky) {
        const float* prow = patch + (c * KH + ky) * PW;
        for (int kx = 0; kx < KW; ++kx) {
          const float wv = __ldg(w_t + (int64_t)((c * KH + ky) * KW + kx) * Cout + co);
#pragma unroll
          for (int p = 0; p < PX; ++p) acc[p] = fmaf(wv, prow[p * stride + kx], acc[p]);
        }
      }
    const float bv = bias ? bias[co] : 0.f;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const int ox = ox0 + p;
      if (ox < Wout) out[(((int64_t)n * Hout + oy) * Wout + ox) * Cout + co] = mage_act(acc[p] + bv, act);
    }
  }
}

// ------------------------------------------------------------------ last f8 decoder layer
// warp per pixel: relu -> 1x1 conv to <=4 channels -> tanh, planar output.
__global__ void __launch_bounds__(256) conv1x1_tanh_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           int64_t n_pix, int HW, int Cin, int Cout, int64_t out_img_stride) {
  const int64_t pix = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= n_pix) return;
  const float4* src = reinterpret_cast<const float4*>(in + pix * Cin);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = lane; i < Cin / 4; i += 32) {
    float4 v = __ldg(src + i);
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < Cout) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (int64_t)c * Cin) + i);
        acc[c] += (v.x * wv.x + v.y * wv.y) + (v.z * wv.z + v.w * wv.w);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = warp_sum(acc[c]);
  if (lane < Cout) {
    const int64_t n = pix / HW, p = pix - n * HW;
    float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
    out[n * out_img_stride + (int64_t)lane * HW + p] = tanhf(v + bias[lane]);
  }
}

__global__ void __launch_bounds__(256) kv_append_kernel(const float* __restrict__ qkv, float* __restrict__ kc,
                                                        float* __restrict__ vc, int M, int C4, int pos, int Lmax) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)M * C4) return;
  const int m = (int)(t / C4), c = (int)(t - (int64_t)m * C4);
  const float4* src = reinterpret_cast<const float4*>(qkv) + (int64_t)m * 3 * C4;
  const int64_t dst = ((int64_t)m * Lmax + pos) * C4 + c;
  reinterpret_cast<float4*>(kc)[dst] = src[C4 + c];
  reinterpret_cast<float4*>(vc)[dst] = src[2 * C4 + c];
}

}  // namespace

extern "C" int mage_layernorm_f32(mage_ctx* ctx, const float* in, const float* gamma, const float* beta, float* out, void* out_split,
                                  int64_t split_plane, int* flag, int rows, int C, float eps, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(rows > 0 && C % 128 == 0 && C <= 1024 && aligned16(in) && aligned16(out) && aligned16(gamma) && aligned16(beta));
  MAGE_CHECK_ARG((out || out_split) && aligned16(out_split) && split_plane % 4 == 0);
  const dim3 g((rows + 7) / 8);
  cudaStream_t st = as_stream(stream);
  __half* sp = reinterpret_cast<__half*>(out_split);
  switch (C / 128) {
    case 1: mage_launch_pdl(ctx, layernorm_kernel<1>, g, 256, 0, st, 1, in, gamma, beta, out, sp, split_plane, flag, rows, eps); break;
    case 2: mage_launch_pdl(ctx, layernorm_kernel<2>, g, 256, 0, st, 1, in, gamma, beta, out, sp, split_plane, flag, rows, eps); break;
    case 4: mage_launch_pdl(ctx, layernorm_kernel<4>, g, 256, 0, st, 1, in, gamma, beta, out, sp, split_plane, flag, rows, eps); break;
    case 8: mage_launch_pdl(ctx, layernorm_kernel<8>, g, 256, 0, st, 1, in, gamma, beta, out, sp, split_plane, flag, rows, eps); break;
    default: return MAGE_EINVAL;
  }
  return mage_post_launch(ctx);
}

extern "C" int mage_argmax_rows_f32(mage_ctx* ctx, const float* x, int64_t ldx, int64_t* idx, int rows, int N, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(rows > 0 && N > 0);
  mage_launch_pdl(ctx, argmax_rows_kernel, (rows + 7) / 8, 256, 0, as_stream(stream), 1, x, ldx, idx, rows, N);
  return mage_post_launch(ctx);
}

extern "C" int mage_embedding_f32(mage_ctx* ctx, const int64_t* idx, const float* table, float* out, int rows, int C, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(rows > 0 && C % 4 == 0 && aligned16(table) && aligned16(out));
  const int64_t total = (int64_t)rows * (C / 4);
  embedding_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(idx, table, out, rows, C / 4);
  return mage_post_launch(ctx);
}

extern "C" int mage_token_taps_f32(mage_ctx* ctx, const int64_t* tok, const float* table, const float* pos_bias, const float* bias, float* out,
                                   int n_img, int R, int K, int C, int KH, int KW, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && R > 0 && K > 0 && C == 512 && KH > 0 && KW > 0 && (KH & 1) && (KW & 1));
  MAGE_CHECK_ARG(aligned16(table) && aligned16(pos_bias) && aligned16(bias) && aligned16(out));
  const int n_pix = n_img * R * R;
  cudaError_t e = mage_launch_pdl(ctx, token_taps_kernel<4>, dim3((n_pix + 7) / 8), dim3(256), 0, as_stream(stream), 1, tok, table, pos_bias,
                                  bias, out, n_pix, R, K, KH, KW, (const float*)nullptr, (const float*)nullptr, 0.f, (__half*)nullptr,
                                  (int64_t)0, (int*)nullptr);
  if (e != cudaSuccess) return (int)e;
  return mage_post_launch(ctx);
}

extern "C" int mage_token_taps_ln_f32(mage_ctx* ctx, const int64_t* tok, const float* table, const float* pos_bias, const float* bias,
                                      float* out, int n_img, int R, int K, int C, int KH, int KW, const float* ln_gamma,
                                      const float* ln_beta, float ln_eps, void* ln_split, int64_t ln_plane, int* flag, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && R > 0 && K > 0 && C == 512 && KH > 0 && KW > 0 && (KH & 1) && (KW & 1));
  MAGE_CHECK_ARG(aligned16(table) && aligned16(pos_bias) && aligned16(bias) && aligned16(out) && ln_gamma && ln_beta && ln_split &&
                 aligned16(ln_gamma) && aligned16(ln_beta) && (reinterpret_cast<uintptr_t>(ln_split) & 7) == 0 && ln_plane % 4 == 0);
  const int n_pix = n_img * R * R;
  cudaError_t e = mage_launch_pdl(ctx, token_taps_kernel<4>, dim3((n_pix + 7) / 8), dim3(256), 0, as_stream(stream), 1, tok, table, pos_bias,
                                  bias, out, n_pix, R, K, KH, KW, ln_gamma, ln_beta, ln_eps, reinterpret_cast<__half*>(ln_split), ln_plane,
                                  flag);
  if (e != cudaSuccess) return (int)e;
  return mage_post_launch(ctx);
}

extern "C" int mage_maxpool2x2_nhwc_f32(mage_ctx* ctx, const float* in, float* out, int n_img, int Hin, int Win, int C, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && Hin % 2 == 0 && Win % 2 == 0 && C % 4 == 0 && aligned16(in) && aligned16(out));
  const int64_t total = (int64_t)n_img * (Hin / 2) * (Win / 2) * (C / 4);
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(in, out, n_img, Hin, Win, C / 4);
  return mage_post_launch(ctx);
}

extern "C" int mage_text_embed_f32(mage_ctx* ctx, const int64_t* text, const float* tok_emb, const float* pos_emb, const float* gamma,
                                   const float* beta, float* x, int32_t* key_len, int B, int T, int C, int pad_idx,
                                   float eps, int vocab, int* flag, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(B > 0 && T > 0 && C == 512 && aligned16(tok_emb) && aligned16(pos_emb) && aligned16(x));
  MAGE_CHECK_ARG(vocab > 0 && pad_idx >= 0 && pad_idx < vocab);
  text_embed_kernel<<<(B * T + 7) / 8, 256, 0, as_stream(stream)>>>(text, tok_emb, pos_emb, gamma, beta, x, key_len, B, T,
                                                                   pad_idx, eps, vocab, flag);
  return mage_post_launch(ctx);
}

extern "C" int mage_adain_nhwc_f32(mage_ctx* ctx, const float* x, const float* gamma, const float* beta, float* out, int n_img, int HW,
                                   int C, float eps, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && HW > 0 && C % 32 == 0);
  adain_kernel<<<dim3(C / 32, n_img), 256, 0, as_stream(stream)>>>(x, gamma, beta, out, HW, C, eps);
  return mage_post_launch(ctx);
}

extern "C" int mage_add_scaled_vec_f32(mage_ctx* ctx, float* x, const float* s, const float* vec, int n_img, int HW, int C, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && HW > 0 && C > 0);
  const int64_t total = (int64_t)n_img * HW * C;
  add_scaled_vec_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, s, vec, HW, C, total);
  return mage_post_launch(ctx);
}

extern "C" int mage_nchw_to_nhwc_f32(mage_ctx* ctx, const float* in, float* out, int n_img, int C, int HW, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && C > 0 && HW > 0);
  const int64_t total = (int64_t)n_img * HW * C;
  nchw_to_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(in, out, C, HW, total);
  return mage_post_launch(ctx);
}

extern "C" int mage_conv2d_first_f32(mage_ctx* ctx, const float* in, const float* w_t, const float* bias, float* out, int n_img, int Cin,
                                     int H, int W, int Hout, int Wout, int Cout, int KH, int KW, int stride, int pad,
                                     int act, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && Cin > 0 && Cin <= 4 && Cout > 0 && KH > 0 && KW > 0 && stride > 0);
  constexpr int PX = 16;
  const int PW = (PX - 1) * stride + KW;
  const size_t smem = (size_t)Cin * KH * PW * sizeof(float);
  MAGE_CHECK_ARG(smem <= 48 * 1024);
  const int segs = (Wout + PX - 1) / PX;
  const int64_t blocks = (int64_t)n_img * Hout * segs;
  MAGE_CHECK_ARG(blocks < ((int64_t)1 << 31));
  conv_first_kernel<PX><<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(in, w_t, bias, out, Cin, H, W, Hout, Wout, Cout,
                                                                           KH, KW, stride, pad, act);
  return mage_post_launch(ctx);
}

extern "C" int mage_conv1x1_tanh_nchw_f32(mage_ctx* ctx, const float* in, const float* w, const float* bias, float* out, int n_img, int HW,
                                          int Cin, int Cout, int64_t out_img_stride, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && HW > 0 && Cin % 4 == 0 && Cout >= 1 && Cout <= 4 && aligned16(in) && aligned16(w) && bias);
  const int64_t n_pix = (int64_t)n_img * HW;
  conv1x1_tanh_kernel<<<(unsigned)((n_pix + 7) / 8), 256, 0, as_stream(stream)>>>(in, w, bias, out, n_pix, HW, Cin, Cout,
                                                                                 out_img_stride);
  return mage_post_launch(ctx);
}

extern "C" int mage_kv_append_f32(mage_ctx* ctx, const float* qkv, float* kcache, float* vcache, int M, int C, int pos, int Lmax,
                                  void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(M > 0 && C % 4 == 0 && pos >= 0 && pos < Lmax && aligned16(qkv) && aligned16(kcache) && aligned16(vcache));
  const int64_t total = (int64_t)M * (C / 4);
  kv_append_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(qkv, kcache, vcache, M, C / 4, pos, Lmax);
  return mage_post_launch(ctx);
}

extern "C" int mage_gn_partial_f32(mage_ctx* ctx, const float* x, double* part, int n_slots, int B, int HW, int C, int groups, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_slots > 0 && B > 0 && HW > 0 && groups > 0 && C % groups == 0 && (C / groups) % 4 == 0 && aligned16(x));
  gn_partial_kernel<<<(unsigned)((int64_t)n_slots * B * groups), 256, 0, as_stream(stream)>>>(x, part, B, HW, C, C / groups);
  return mage_post_launch(ctx);
}

extern "C" int mage_gn_silu_head_f32(mage_ctx* ctx, const float* x, const double* part, const float* gamma, const float* beta, const float* w,
                                     const float* bias, float* out, int rows, int B, int HW, int n_slots, int C, int groups,
                                     int cout, float eps, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(rows > 0 && B > 0 && HW > 0 && n_slots > 0 && C == 512 && groups == 32 && cout >= 1 && cout <= 8);
  MAGE_CHECK_ARG(aligned16(x) && aligned16(gamma) && aligned16(beta));
  gn_head_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, part, gamma, beta, w, bias, out, rows, B, HW, n_slots,
                                                                            cout, eps);
  return mage_post_launch(ctx);
}
