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
// Axial (H / W) attention of one decode position: 16 queries x 16 keys x 16 heads x 32 dims per line.
// One warp per (image, line, head): the 16x32 Q, K, V tiles of the line are read ONCE with coalesced
// 128-byte row segments into shared memory (the generic kernel re-reads K/V per query), scores and
// P.V run out of shared memory with broadcast reads, the 16x32 output tile goes back through shared
// memory so every store is a full 128-byte row segment.  HBM-bound: qkv in, attention out, once each.
//   lane = (qi = lane & 15, half = lane >> 4): scores for keys half*8..+7, output dims half*16..+15.
// ---------------------------------------------------------------------------------------------
constexpr int AX_S = 16, AX_D = 32, AX_QLD = 36;  // Q rows padded to 36 floats (2-way instead of 16-way conflicts)
constexpr int AX_WARPS = 4;  // 4 x 6.3 KB of tiles per block

struct AxialArgs {
  const float* qkv;  // [rows, 3C]: q | k | v
  float* out;        // optional fp32 [rows, C]
  __half* split;     // optional split copy
  int64_t split_plane;
  int* flag;
  int n_lines;       // B * 16 lines
  int n_head, C;
  int64_t line_outer, line_inner, seq;  // in rows: row(b, line, s) = b*line_outer + line*line_inner + s*seq
  int lines_per_img;
  float scale;
};

__global__ void __launch_bounds__(AX_WARPS * 32) axial_attn_kernel(const AxialArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float sq[AX_WARPS][AX_S * AX_QLD];
  __shared__ __align__(16) float sk[AX_WARPS][AX_S * AX_D];
  __shared__ __align__(16) float sv[AX_WARPS][AX_S * AX_D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t unit = (int64_t)blockIdx.x * AX_WARPS + warp;  // (line, head)
  if (unit >= (int64_t)p.n_lines * p.n_head) return;
  const int head = (int)(unit % p.n_head);
  const int line = (int)(unit / p.n_head);
  const int b = line / p.lines_per_img, l = line - b * p.lines_per_img;
  const int64_t row0 = b * p.line_outer + l * p.line_inner;
  const int C3 = 3 * p.C;

  // coalesced tile loads: 8 lanes cover one 128-byte row segment, 4 rows per instruction
  {
    const int r4 = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int s = it * 4 + r4;
      const float* src = p.qkv + (row0 + s * p.seq) * C3 + head * AX_D + c4;
      const float4 q = __ldg(reinterpret_cast<const float4*>(src));
      const float4 k = __ldg(reinterpret_cast<const float4*>(src + p.C));
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + 2 * p.C));
      *reinterpret_cast<float4*>(&sq[warp][s * AX_QLD + c4]) = q;
      *reinterpret_cast<float4*>(&sk[warp][s * AX_D + c4]) = k;
      *reinterpret_cast<float4*>(&sv[warp][s * AX_D + c4]) = v;
    }
  }
  __syncwarp();
  const int qi = lane & 15, half = lane >> 4;
  float sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[j] = 0.f;
#pragma unroll
  for (int d4 = 0; d4 < AX_D / 4; ++d4) {
    const float4 q = *reinterpret_cast<const float4*>(&sq[warp][qi * AX_QLD + d4 * 4]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 k = *reinterpret_cast<const float4*>(&sk[warp][(half * 8 + j) * AX_D + d4 * 4]);  // broadcast
      sc[j] = fmaf(q.x, k.x, sc[j]); sc[j] = fmaf(q.y, k.y, sc[j]);
      sc[j] = fmaf(q.z, k.z, sc[j]); sc[j] = fmaf(q.w, k.w, sc[j]);
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] *= p.scale; m = fmaxf(m, sc[j]); }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = expf(sc[j] - m); sum += sc[j]; }
  sum += __shfl_xor_sync(0xffffffffu, sum, 16);
  const float inv = 1.f / sum;
  // all 16 probabilities of this query: own 8 + the partner lane's 8 (selects keep the arrays in registers)
  float plo[8], phi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float other = __shfl_xor_sync(0xffffffffu, sc[j], 16);
    plo[j] = half ? other : sc[j];
    phi[j] = half ? sc[j] : other;
  }
  float4 o[4];
#pragma unroll
  for (int d4 = 0; d4 < 4; ++d4) o[d4] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < AX_S; ++j) {
    const float pj = (j < 8) ? plo[j & 7] : phi[j & 7];
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 v = *reinterpret_cast<const float4*>(&sv[warp][j * AX_D + half * 16 + d4 * 4]);  // broadcast per half
      o[d4].x = fmaf(pj, v.x, o[d4].x); o[d4].y = fmaf(pj, v.y, o[d4].y);
      o[d4].z = fmaf(pj, v.z, o[d4].z); o[d4].w = fmaf(pj, v.w, o[d4].w);
    }
  }
  __syncwarp();  // everyone is done reading sq before it becomes the output staging tile
#pragma unroll
  for (int d4 = 0; d4 < 4; ++d4) {
    float4 v = o[d4];
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    *reinterpret_cast<float4*>(&sq[warp][qi * AX_QLD + half * 16 + d4 * 4]) = v;
  }
  __syncwarp();
  {
    const int r4 = lane >> 3, c4 = (lane & 7) * 4;
    bool bad = false;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int s = it * 4 + r4;
      const float4 v = *reinterpret_cast<const float4*>(&sq[warp][s * AX_QLD + c4]);
      const int64_t e = (row0 + s * p.seq) * p.C + head * AX_D + c4;
      if (p.out) *reinterpret_cast<float4*>(p.out + e) = v;
      if (p.split) {
        uint2 hi, lo;
        bad |= tc::split4(v, hi, lo);
        *reinterpret_cast<uint2*>(p.split + e) = hi;
        *reinterpret_cast<uint2*>(p.split + p.split_plane + e) = lo;
      }
    }
    if (bad && p.flag) atomicOr(p.flag, 1);
  }
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
// Bounded like tc::mbar_wait (tc_common.cuh): a copy that never lands must surface as a trap (launch error), not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  if (tc::mbar_try_wait(smem_u32(bar), phase)) return;
  const long long t0 = clock64();
  while (!tc::mbar_try_wait(smem_u32(bar), phase)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 2 GHz
  }
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
  pdl_launch_dependents();
  extern __shared__ __align__(128) float smem[];
  float* Ks = smem;                       // [pos+1][256]: sized to the live prefix so early steps fit more CTAs per SM
  float* Vs = smem + (size_t)(pos + 1) * TA_HALF;
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ __align__(16) float qs[TA_HALF];

  const int m = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* row = qkv + (int64_t)m * 3 * TA_C + half * TA_HALF;
  // cache layout [M][2 head-halves][Lmax][256]: the positions 0..pos-1 of this (row, half) are ONE contiguous block, so each
  // of K and V arrives with a single bulk copy of pos KB instead of pos copies of 1 KB
  float* kc = kcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;
  float* vc = vcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();   // the preceding grid (this step's QKV GEMM) has completed before anything global is touched
  __syncthreads();
  if (tid == 0 && pos > 0) {
    const uint32_t bytes = (uint32_t)pos * TA_HALF * 4;
    mbar_expect_tx(&bars[0], bytes);
    bulk_g2s(Ks, kc, bytes, &bars[0]);
    mbar_expect_tx(&bars[1], bytes);
    bulk_g2s(Vs, vc, bytes, &bars[1]);
  }
  // this position's q, k, v: registers -> shared (math) and -> cache (future steps)
  {
    const float qv = row[tid], kv = row[TA_C + tid], vv = row[2 * TA_C + tid];
    qs[tid] = qv;
    Ks[(size_t)pos * TA_HALF + tid] = kv;
    Vs[(size_t)pos * TA_HALF + tid] = vv;
    kc[(int64_t)pos * TA_HALF + tid] = kv;
    vc[(int64_t)pos * TA_HALF + tid] = vv;
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


// ---------------------------------------------------------------------------------------------
// The same step as a PERSISTENT kernel with a ring of staging slots.  The one-shot form above launches M*2 short-lived CTAs whose
// life is a chain of latencies (launch, barrier set-up, the q/k/v row, the bulk copies): ~130 us for 32768 CTAs at position 0
// whatever the traffic, 68 % of the copy bandwidth at 8 prompts.  Here a few CTAs per SM each walk over units (location, head
// half) with NS slots in flight: slot = this unit's q row, K[0..pos] and V[0..pos] (prefix by bulk copy from the cache, the new
// k / v rows by bulk copy from qkv straight into row `pos`), one mbarrier per slot; after a unit's math the block syncs and one
// thread refills the slot with the unit NS ahead.  The math per unit is the one-shot kernel's, statement for statement, so the
// two forms give the same bits (tests compare them).
// MEASURED (profiles/r02ai_temporal_attn_ring_ab.txt, same box, three repetitions): the ring form is 1.2-1.4 % SLOWER per generate
// at 8, 16 and 64 prompts -- thousands of small resident CTAs already keep more copies in flight than a ring per CTA does, and the
// block-wide hand-over per unit costs more than the CTA launches it saves.  OFF by default (MAGE_TATTN_RING=1 / mage_temporal_attn_ring).
// ---------------------------------------------------------------------------------------------
constexpr int TA_MAX_SLOTS = 8;

__global__ void __launch_bounds__(256, 4) temporal_attn_ring_kernel(const float* __restrict__ qkv, float* __restrict__ kcache,
                                                                 float* __restrict__ vcache, float* __restrict__ out,
                                                                 __half* __restrict__ split, int64_t split_plane, int* flag,
                                                                 int pos, int Lmax, float scale, int units, int ns) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[TA_MAX_SLOTS];
  const int S = pos + 1;
  const int slot_floats = (2 * S + 1) * TA_HALF;          // q | K[0..pos] | V[0..pos]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  auto fill = [&](int slot, int u) {                       // one thread: everything unit u needs, tracked by the slot's barrier
    const int m = u >> 1, half = u & 1;
    float* base = smem + (size_t)slot * slot_floats;
    float* Ks = base + TA_HALF;
    float* Vs = Ks + (size_t)S * TA_HALF;
    const float* row = qkv + (int64_t)m * 3 * TA_C + half * TA_HALF;
    const float* kc = kcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;
    const float* vc = vcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;
    const uint32_t prefix = (uint32_t)pos * TA_HALF * 4;
    mbar_expect_tx(&bars[slot], 2 * prefix + 3 * TA_HALF * 4);
    bulk_g2s(base, row, TA_HALF * 4, &bars[slot]);
    bulk_g2s(Ks + (size_t)pos * TA_HALF, row + TA_C, TA_HALF * 4, &bars[slot]);
    bulk_g2s(Vs + (size_t)pos * TA_HALF, row + 2 * TA_C, TA_HALF * 4, &bars[slot]);
    if (pos > 0) {
      bulk_g2s(Ks, kc, prefix, &bars[slot]);
      bulk_g2s(Vs, vc, prefix, &bars[slot]);
    }
  };

  if (tid == 0) {
    for (int i = 0; i < ns; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();   // the preceding grid (this step's QKV GEMM) has completed before anything global is touched
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < ns; ++i) {
      const int u = blockIdx.x + i * gridDim.x;
      if (u < units) fill(i, u);
    }
  }
  int it = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
    const int slot = it % ns;
    mbar_wait(&bars[slot], (it / ns) & 1);
    const int m = u >> 1, half = u & 1;
    const float* qs = smem + (size_t)slot * slot_floats;
    const float* Ks = qs + TA_HALF;
    const float* Vs = Ks + (size_t)S * TA_HALF;
    {   // this position's k, v join the cache for the later steps
      float* kc = kcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;
      float* vc = vcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;
      kc[(int64_t)pos * TA_HALF + tid] = Ks[(size_t)pos * TA_HALF + tid];
      vc[(int64_t)pos * TA_HALF + tid] = Vs[(size_t)pos * TA_HALF + tid];
    }
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
    float o = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
      o = fmaf(pj, Vs[(size_t)j * TA_HALF + warp * 32 + lane], o);
    }
    store_attn(out, split, split_plane, flag, (int64_t)m * TA_C + half * TA_HALF + tid, o * inv);
    __syncthreads();   // every thread is done with the slot
    if (tid == 0) {
      const int un = u + ns * gridDim.x;
      if (un < units) fill(slot, un);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Temporal attention over SEVERAL consecutive positions at once (the full-sequence form of the MAGE+ suffix re-evaluation):
// the queries of positions pos0 .. pos0+n_pos-1 of one (location, head-half) share one K/V prefix, so it is staged ONCE --
// cached positions 0..pos0-1 by bulk copy, the n_pos new k,v rows from qkv (also appended to the cache) -- and every query
// attends keys 0..its own position (causal).  Same grid / warp roles / math order as temporal_attn_kernel, looped over queries.
//   qkv rows: position s of location m at row s*M + m.
// ---------------------------------------------------------------------------------------------
constexpr int TAS_GROUPS = 1;   // query groups: warp = (head, group); group g takes queries g, g+G, ... (measured: 4 groups = 1024-thread
                                // CTAs are SLOWER, 305 vs 250 ms per MAGE+ generate: the kernel is instruction-bound, not latency-bound)
__global__ void __launch_bounds__(256 * TAS_GROUPS) temporal_attn_seq_kernel(const float* __restrict__ qkv, float* __restrict__ kcache,
                                                                float* __restrict__ vcache, __half* __restrict__ split,
                                                                int64_t split_plane, int* flag, int M, int pos0, int n_pos,
                                                                int Lmax, float scale) {
  extern __shared__ __align__(128) float smem[];
  const int S_all = pos0 + n_pos;
  float* Ks = smem;
  float* Vs = smem + (size_t)S_all * TA_HALF;
  float* Qs = Vs + (size_t)S_all * TA_HALF;      // [n_pos][256]: every query of this (location, half), staged up front so that the
  __shared__ __align__(8) uint64_t bars[2];      // eight head-warps run through their queries without block-wide barriers
  const int m = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int tid = threadIdx.x & 255, grp = threadIdx.x >> 8, warp = tid >> 5, lane = tid & 31;
  float* kc = kcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;
  float* vc = vcache + ((int64_t)m * 2 + half) * Lmax * TA_HALF;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && pos0 > 0) {
    const uint32_t bytes = (uint32_t)pos0 * TA_HALF * 4;
    mbar_expect_tx(&bars[0], bytes);
    bulk_g2s(Ks, kc, bytes, &bars[0]);
    mbar_expect_tx(&bars[1], bytes);
    bulk_g2s(Vs, vc, bytes, &bars[1]);
  }
  for (int s = grp; s < n_pos; s += TAS_GROUPS) {   // staging is shared out over the groups too
    const float* row = qkv + ((int64_t)s * M + m) * 3 * TA_C + half * TA_HALF;
    const float kv = row[TA_C + tid], vv = row[2 * TA_C + tid];
    Qs[(size_t)s * TA_HALF + tid] = row[tid];
    Ks[(size_t)(pos0 + s) * TA_HALF + tid] = kv;
    Vs[(size_t)(pos0 + s) * TA_HALF + tid] = vv;
    kc[(int64_t)(pos0 + s) * TA_HALF + tid] = kv;
    vc[(int64_t)(pos0 + s) * TA_HALF + tid] = vv;
  }
  __syncthreads();
  if (pos0 > 0) {
    mbar_wait(&bars[0], 0);
    mbar_wait(&bars[1], 0);
  }
  for (int s = grp; s < n_pos; s += TAS_GROUPS) {
    const int S = pos0 + s + 1;
    const float* qh = Qs + (size_t)s * TA_HALF + warp * 32;
    float sc[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = lane + t * 32;
      sc[t] = -INFINITY;
      if (j < S) {
        const float* kr = Ks + (size_t)j * TA_HALF + warp * 32;
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int dd = (i + lane) & 31;
          d = fmaf(qh[dd], kr[dd], d);
        }
        sc[t] = d * scale;
      }
    }
    const float mx = warp_max(fmaxf(sc[0], sc[1]));
    const float p0 = (lane < S) ? expf(sc[0] - mx) : 0.f;
    const float p1 = (lane + 32 < S) ? expf(sc[1] - mx) : 0.f;
    const float inv = 1.f / warp_sum(p0 + p1);
    float o = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
      o = fmaf(pj, Vs[(size_t)j * TA_HALF + warp * 32 + lane], o);
    }
    store_attn(nullptr, split, split_plane, flag, ((int64_t)s * M + m) * TA_C + half * TA_HALF + tid, o * inv);
  }
}

}  // namespace

extern "C" int mage_mha_f32(mage_ctx* ctx, const float* q, const float* k, const float* v, float* out, int n_outer, int n_inner, int n_head,
                            int Sq, int Sk, int64_t q_outer, int64_t q_inner, int64_t q_seq, int64_t k_outer, int64_t k_inner,
                            int64_t k_seq, int64_t v_outer, int64_t v_inner, int64_t v_seq, int64_t o_outer, int64_t o_inner,
                            int64_t o_seq, const int32_t* key_len, float scale, void* out_split, int64_t split_plane,
                            int* flag, void* stream) {
  MAGE_CHECK_CTX(ctx);
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
  return mage_post_launch(ctx);
}

extern "C" int mage_axial_attn_f32(mage_ctx* ctx, const float* qkv, float* out, void* out_split, int64_t split_plane, int* flag, int B, int R,
                                   int n_head, int axis, float scale, void* stream) {
  MAGE_CHECK_CTX(ctx);
  // qkv [B*R*R, 3C] rows ordered (b, h, w); axis 1: sequences run over h for fixed (b, w); axis 2: over w for fixed (b, h)
  MAGE_CHECK_ARG(B > 0 && R == AX_S && n_head > 0 && (axis == 1 || axis == 2) && (out || out_split));
  MAGE_CHECK_ARG(aligned16(qkv) && aligned16(out) && aligned16(out_split) && split_plane % 4 == 0);
  AxialArgs a{};
  a.qkv = qkv; a.out = out; a.split = reinterpret_cast<__half*>(out_split); a.split_plane = split_plane; a.flag = flag;
  a.n_lines = B * R; a.n_head = n_head; a.C = n_head * AX_D; a.lines_per_img = R; a.scale = scale;
  a.line_outer = (int64_t)R * R;
  a.line_inner = axis == 1 ? 1 : R;
  a.seq = axis == 1 ? R : 1;
  const int64_t units = (int64_t)a.n_lines * n_head;
  mage_launch_pdl(ctx, axial_attn_kernel, (unsigned)((units + AX_WARPS - 1) / AX_WARPS), AX_WARPS * 32, 0, as_stream(stream), 1, a);
  return mage_post_launch(ctx);
}

extern "C" int mage_temporal_attn_step_f32(mage_ctx* ctx, const float* qkv, float* kcache, float* vcache, float* out, void* out_split,
                                           int64_t split_plane, int* flag, int M, int pos, int Lmax, float scale,
                                           void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(M > 0 && pos >= 0 && pos < Lmax && Lmax <= 64);
  MAGE_CHECK_ARG(aligned16(qkv) && aligned16(kcache) && aligned16(vcache) && aligned16(out) && (out || out_split));
  // ring form: NS slots of (q | K[0..pos] | V[0..pos]) per CTA -- about 96 KB in flight per CTA, at least two slots -- and as many
  // CTAs per SM as fit (occupancy calculator); the grid never exceeds the units.  Very long clips (two slots > 220 KB) keep the
  // one-shot form.
  const size_t ring_slot = (size_t)(2 * (pos + 1) + 1) * TA_HALF * sizeof(float);
  int ns = (int)((size_t)96 * 1024 / ring_slot);
  ns = ns < 2 ? 2 : (ns > TA_MAX_SLOTS ? TA_MAX_SLOTS : ns);
  if (ctx->tattn_ring && (size_t)ns * ring_slot <= (size_t)220 * 1024) {
    const size_t smem = (size_t)ns * ring_slot;
    size_t smem_max = (size_t)2 * (2 * Lmax + 1) * TA_HALF * sizeof(float);      // two slots at the last position
    if (smem_max < (size_t)96 * 1024) smem_max = (size_t)96 * 1024;
    if (smem_max > (size_t)220 * 1024) smem_max = (size_t)220 * 1024;
    const void* fn = reinterpret_cast<const void*>(&temporal_attn_ring_kernel);
    int k = ctx->find(fn);
    if (k < 0 || ctx->cfg_smem[k] < smem_max) {
      cudaError_t e = cudaFuncSetAttribute(temporal_attn_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
      if (e != cudaSuccess) return (int)e;
      if (k < 0) k = ctx->add(fn, 0, smem_max);
      if (k >= 0) ctx->cfg_smem[k] = smem_max;
    }
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, temporal_attn_ring_kernel, 256, smem) != cudaSuccess || per_sm < 1) {
      (void)cudaGetLastError();
      per_sm = 1;
    }
    const int units = M * 2;
    int64_t grid = (int64_t)ctx->sms * per_sm;
    if (grid > units) grid = units;
    mage_launch_pdl(ctx, temporal_attn_ring_kernel, (unsigned)grid, 256, smem, as_stream(stream), 1, qkv, kcache, vcache, out,
                    reinterpret_cast<__half*>(out_split), split_plane, flag, pos, Lmax, scale, units, ns);
    return mage_post_launch(ctx);
  }
  const size_t smem = (size_t)2 * (pos + 1) * TA_HALF * sizeof(float);
  const size_t smem_max = (size_t)2 * Lmax * TA_HALF * sizeof(float);
  if (smem_max > 40 * 1024) {   // opt in early: static shared memory counts against the 48 KB default too.  Per-device attribute: remembered in the handle
    const void* fn = reinterpret_cast<const void*>(&temporal_attn_kernel);
    int k = ctx->find(fn);
    if (k < 0 || ctx->cfg_smem[k] < smem_max) {
      cudaError_t e = cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
      if (e != cudaSuccess) return (int)e;
      if (k < 0) k = ctx->add(fn, 0, smem_max);
      if (k >= 0) ctx->cfg_smem[k] = smem_max;
    }
  }
  mage_launch_pdl(ctx, temporal_attn_kernel, (unsigned)M * 2, 256, smem, as_stream(stream), 1, qkv, kcache, vcache, out,
                  reinterpret_cast<__half*>(out_split), split_plane, flag, pos, Lmax, scale);
  return mage_post_launch(ctx);
}

extern "C" int mage_temporal_attn_seq_f32(mage_ctx* ctx, const float* qkv, float* kcache, float* vcache, void* out_split,
                                          int64_t split_plane, int* flag, int M, int pos0, int n_pos, int Lmax, float scale,
                                          void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(M > 0 && pos0 >= 0 && n_pos > 0 && pos0 + n_pos <= Lmax && Lmax <= 64 && out_split != nullptr);
  MAGE_CHECK_ARG(aligned16(qkv) && aligned16(kcache) && aligned16(vcache));
  const size_t smem = (size_t)(2 * (pos0 + n_pos) + n_pos) * TA_HALF * sizeof(float);
  const size_t smem_max = (size_t)3 * Lmax * TA_HALF * sizeof(float);
  if (smem_max > 40 * 1024) {   // static shared memory counts against the 48 KB default as well
    const void* fn = reinterpret_cast<const void*>(&temporal_attn_seq_kernel);
    int k = ctx->find(fn);
    if (k < 0 || ctx->cfg_smem[k] < smem_max) {
      cudaError_t e = cudaFuncSetAttribute(temporal_attn_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
      if (e != cudaSuccess) return (int)e;
      if (k < 0) k = ctx->add(fn, 0, smem_max);
      if (k >= 0) ctx->cfg_smem[k] = smem_max;
    }
  }
  temporal_attn_seq_kernel<<<(unsigned)M * 2, 256 * TAS_GROUPS, smem, as_stream(stream)>>>(qkv, kcache, vcache, reinterpret_cast<__half*>(out_split),
                                                                           split_plane, flag, M, pos0, n_pos, Lmax, scale);
  return mage_post_launch(ctx);
}
