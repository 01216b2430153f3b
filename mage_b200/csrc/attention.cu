// Multi-head attention cores (head_dim 32, fp32).
//
//  * mha_kernel: generic strided SDPA, one warp per (outer, inner, head, query), keys spread over
//    lanes.  Serves the H-/W-axial blocks, the text self-attention (key padding) and the motion-anchor
//    cross-attention -- all of them short sequences (<= 64 keys) with no reuse worth a tile.
//  * temporal_attn_kernel: the one bandwidth-bound attention on the path -- one query per location
//    against the growing K/V cache.  K/V rows are staged into shared memory with bulk async copies
//    (cp.async.bulk, the TMA engine) that complete on an mbarrier, so the copy engine streams HBM
//    while other resident CTAs compute.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

struct MhaArgs {
  const float* q; const float* k; const float* v; float* out;
  int n_outer, n_inner, n_head, Sq, Sk;
  int64_t q_outer, q_inner, q_seq, k_outer, k_inner, k_seq, v_outer, v_inner, v_seq, o_outer, o_inner, o_seq;
  const int32_t* key_len;
  float scale;
  int64_t total;
  __half* split;  // optional split (fp16 hi/lo) copy of `out`, same element offsets
  int64_t split_plane;
  int* flag;
};

__device__ __forceinline__ void store_attn(float* out, __half* split, int64_t plane, int* flag, int64_t e, float v) {
  if (out) out[e] = v;
  if (split) {
    __half hi, lo;
    tc::split_one(v, hi, lo);
    split[e] = hi;
    split[plane + e] = lo;
    if (!(fabsf(v) <= 65504.f) && flag) atomicOr(flag, 1);
  }
}

__global__ void __launch_bounds__(256) mha_kernel(const MhaArgs p) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= p.total) return;
  const int qi = (int)(w % p.Sq);
  int64_t r = w / p.Sq;
  const int h = (int)(r % p.n_head); r /= p.n_head;
  const int inner = (int)(r % p.n_inner);
  const int outer = (int)(r / p.n_inner);

  const float4* qp = reinterpret_cast<const float4*>(p.q + outer * p.q_outer + inner * p.q_inner + qi * p.q_seq + h * 32);
  float4 qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = __ldg(qp + i);

  int klen = p.Sk;
  if (p.key_len) klen = min(klen, p.key_len[outer]);
  const float* kb = p.k + outer * p.k_outer + inner * p.k_inner + h * 32;
  const float* vb = p.v + outer * p.v_outer + inner * p.v_inner + h * 32;

  float s[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = lane + t * 32;
    s[t] = -INFINITY;
    if (j < klen) {
      const float4* kp = reinterpret_cast<const float4*>(kb + j * p.k_seq);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 kv = __ldg(kp + i);
        d = fmaf(qv[i].x, kv.x, d); d = fmaf(qv[i].y, kv.y, d);
        d = fmaf(qv[i].z, kv.z, d); d = fmaf(qv[i].w, kv.w, d);
      }
      s[t] = d * p.scale;
    }
  }
  const float m = warp_max(fmaxf(s[0], s[1]));
  const float p0 = (lane < klen) ? expf(s[0] - m) : 0.f;
  const float p1 = (lane + 32 < klen) ? expf(s[1] - m) : 0.f;
  const float inv = 1.f / warp_sum(p0 + p1);
  float o = 0.f;
  for (int j = 0; j < klen; ++j) {
    const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
    o = fmaf(pj, __ldg(vb + j * p.v_seq + lane), o);
  }
  store_attn(p.out, p.split, p.split_plane, p.flag, outer * p.o_outer + inner * p.o_inner + qi * p.o_seq + h * 32 + lane, o * inv);
}

// ---------------------------------------------------------------------------------------------
// Temporal attention step with bulk-async (TMA) staging of the K/V cache.
//   grid  = M locations x 2 head-halves,  block = 256 threads = 8 warps = 8 heads x 32 lanes
//   smem  = K[Lmax][256] + V[Lmax][256] floats (the 8 heads' 1 KB slice of every cached position)
// Positions 0..pos-1 come from the cache through cp.async.bulk (one 1 KB copy per position and
// tensor, issued by the lanes of warp 0 and tracked by one mbarrier each for K and V); position
// `pos` (this step's k,v) is taken from qkv, written to shared memory for the math and to the
// cache for later steps.  Scores: lane = key, skewed over the 32 dims so the shared-memory reads
// are conflict free; output: lane = dim.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int TA_C = 512, TA_HALF = 256;

__global__ void __launch_bounds__(256) temporal_attn_kernel(const float* __restrict__ qkv, float* __restrict__ kcache,
                                                            float* __restrict__ vcache, float* __restrict__ out,
                                                            __half* __restrict__ split, int64_t split_plane, int* flag,
                                                            int pos, int Lmax, float scale) {
  extern __shared__ __align__(128) float smem[];
  float* Ks = smem;                       // [Lmax][256]
  float* Vs = smem + (size_t)Lmax * TA_HALF;
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ __align__(16) float qs[TA_HALF];

  const int m = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* row = qkv + (int64_t)m * 3 * TA_C + half * TA_HALF;
  float* kc = kcache + (int64_t)m * Lmax * TA_C + half * TA_HALF;
  float* vc = vcache + (int64_t)m * Lmax * TA_C + half * TA_HALF;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0 && pos > 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars[0], (uint32_t)pos * TA_HALF * 4);
      mbar_expect_tx(&bars[1], (uint32_t)pos * TA_HALF * 4);
    }
    __syncwarp();
    for (int j = lane; j < pos; j += 32) bulk_g2s(Ks + (size_t)j * TA_HALF, kc + (int64_t)j * TA_C, TA_HALF * 4, &bars[0]);
    for (int j = lane; j < pos; j += 32) bulk_g2s(Vs + (size_t)j * TA_HALF, vc + (int64_t)j * TA_C, TA_HALF * 4, &bars[1]);
  }
  // this position's q, k, v: registers -> shared (math) and -> cache (future steps)
  {
    const float qv = row[tid], kv = row[TA_C + tid], vv = row[2 * TA_C + tid];
    qs[tid] = qv;
    Ks[(size_t)pos * TA_HALF + tid] = kv;
    Vs[(size_t)pos * TA_HALF + tid] = vv;
    kc[(int64_t)pos * TA_C + tid] = kv;
    vc[(int64_t)pos * TA_C + tid] = vv;
  }
  __syncthreads();
  if (pos > 0) mbar_wait(&bars[0], 0);

  const int S = pos + 1;
  const float* qh = qs + warp * 32;
  float s[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = lane + t * 32;
    s[t] = -INFINITY;
    if (j < S) {
      const float* kr = Ks + (size_t)j * TA_HALF + warp * 32;
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int dd = (i + lane) & 31;  // skew: lane j starts at dim j -> distinct banks
        d = fmaf(qh[dd], kr[dd], d);
      }
      s[t] = d * scale;
    }
  }
  const float mx = warp_max(fmaxf(s[0], s[1]));
  const float p0 = (lane < S) ? expf(s[0] - mx) : 0.f;
  const float p1 = (lane + 32 < S) ? expf(s[1] - mx) : 0.f;
  const float inv = 1.f / warp_sum(p0 + p1);
  if (pos > 0) mbar_wait(&bars[1], 0);
  float o = 0.f;
  for (int j = 0; j < S; ++j) {
    const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
    o = fmaf(pj, Vs[(size_t)j * TA_HALF + warp * 32 + lane], o);
  }
  store_attn(out, split, split_plane, flag, (int64_t)m * TA_C + half * TA_HALF + tid, o * inv);
}

}  // namespace

extern "C" int mage_mha_f32(const float* q, const float* k, const float* v, float* out, int n_outer, int n_inner, int n_head,
                            int Sq, int Sk, int64_t q_outer, int64_t q_inner, int64_t q_seq, int64_t k_outer, int64_t k_inner,
                            int64_t k_seq, int64_t v_outer, int64_t v_inner, int64_t v_seq, int64_t o_outer, int64_t o_inner,
                            int64_t o_seq, const int32_t* key_len, float scale, void* out_split, int64_t split_plane,
                            int* flag, void* stream) {
  MAGE_CHECK_ARG(n_outer > 0 && n_inner > 0 && n_head > 0 && Sq > 0 && Sk > 0 && Sk <= 64);
  MAGE_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && (out || out_split));
  MAGE_CHECK_ARG(((q_outer | q_inner | q_seq | k_outer | k_inner | k_seq) & 3) == 0);
  MhaArgs a{q, k, v, out, n_outer, n_inner, n_head, Sq, Sk, q_outer, q_inner, q_seq, k_outer, k_inner, k_seq,
            v_outer, v_inner, v_seq, o_outer, o_inner, o_seq, key_len, scale, 0, reinterpret_cast<__half*>(out_split),
            split_plane, flag};
  a.total = (int64_t)n_outer * n_inner * n_head * Sq;
  const int64_t blocks = (a.total + 7) / 8;
  MAGE_CHECK_ARG(blocks < ((int64_t)1 << 31));
  mha_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a);
  return mage_post_launch();
}

extern "C" int mage_temporal_attn_step_f32(const float* qkv, float* kcache, float* vcache, float* out, void* out_split,
                                           int64_t split_plane, int* flag, int M, int pos, int Lmax, float scale,
                                           void* stream) {
  MAGE_CHECK_ARG(M > 0 && pos >= 0 && pos < Lmax && Lmax <= 64);
  MAGE_CHECK_ARG(aligned16(qkv) && aligned16(kcache) && aligned16(vcache) && aligned16(out) && (out || out_split));
  const size_t smem = (size_t)2 * Lmax * TA_HALF * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  temporal_attn_kernel<<<(unsigned)M * 2, 256, smem, as_stream(stream)>>>(qkv, kcache, vcache, out, reinterpret_cast<__half*>(out_split),
                                                                          split_plane, flag, pos, Lmax, scale);
  return mage_post_launch();
}
