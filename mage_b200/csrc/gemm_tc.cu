// Tensor-core back end: fp32-grade GEMM and implicit-GEMM convolution on tcgen05 (sm_100a).
//
//   C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual)
//
// Operands arrive in the fp16 hi/lo split format (tc_common.cuh): every k-step issues three
// tcgen05.mma kind::f16 instructions (hi*hi -> main, lo*hi + hi*lo -> corr), both accumulators live
// in TMEM, the epilogue forms main + corr*2^-11 -- fp32-grade accuracy at a third of the fp16 rate.
//
// One persistent CTA per SM (or CTA pair per TPC), 320 threads, warp-specialised; every role loop runs warp-wide in uniform
// control flow and only the issue is predicated on elect.sync (tc_common.cuh: elect_one):
//   warp 0   TMA producer: one 5-D box for the A tile (both planes) + one 3-D box for the W tile per k-block into a
//            SWIZZLE_128B ring (mbarrier full/empty)
//   warp 1   TMEM allocator + MMA issuer (tcgen05.mma on uniform-register descriptors, tcgen05.commit)
//   warps 2-9 epilogue, row per thread (two warps per TMEM lane quadrant): tcgen05.ld -> main + corr*2^-11 -> bias /
//            activation / residual -> swizzled 4 KB staging tile -> one TMA store per output kind (fp32 tile, or the hi+lo
//            planes of the split format = the next consumer's operand)
// Kernels: tc_gemm_kernel (GEMM; convolution with one TMA box per tap), tc_conv_halo_kernel (convolution with one input patch
// per channel block and shifted descriptors per tap, optional fused pixel head).  For a convolution the A tile is an NHWC
// patch and TMA's out-of-bounds zero fill is the padding, so im2col never exists.
//
// CG = 2 (CTA pair, cluster 2x1x1, tcgen05 cta_group::2): the two CTAs of a TPC share one 256 x BN tile.  Each loads its own
// 128 A rows and HALF of the W tile (BN/2 rows); the leader issues M = 256 MMAs that read B from both CTAs' shared memory and
// write each CTA's own TMEM.  256 x 128 with two accumulator stages is the default shape of the path.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128;      // rows per tile = TMEM lanes
constexpr int BK = 64;       // halves per k-block = one 128-byte swizzle row
constexpr int UK = 16;       // K of one tcgen05.mma kind::f16
constexpr int NTHREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int SMEM_BUDGET = 227 * 1024;
// TIMING EXPERIMENT ONLY (tools/experiments/build_oneacc.sh builds a separate library with -DMAGE_EXPERIMENT_ONEACC): all three
// MMAs of a product accumulate into ONE TMEM accumulator, which frees the columns for double-buffered 256-wide tiles.  With the
// 2^11-scaled lo plane of the shipped split format the RESULTS ARE WRONG; the build exists to measure what a single-accumulator
// format (unscaled lo plane, power-of-two pre-scaled tensors) would buy before committing to it.  Never part of libmage_sm100.so.
#ifdef MAGE_EXPERIMENT_ONEACC
constexpr bool kOneAcc = true;
#else
constexpr bool kOneAcc = false;
#endif

struct TcParams {
  const float* bias;
  const float* res;
  float* out;
  __half* split;
  __half* split_relu;
  int* flag;
  int64_t ldr, ldc, split_plane, split_relu_plane, out_img_stride;
  int M, N, act, res_mod;
  int m_tiles, n_tiles, k_iters;
  // convolution geometry (conv != 0)
  int conv, Hout, Wout, Wb, Hb, KW, cin_blocks, cin, pad_y, pad_x;
  int res_mode, out_sy, out_sx, out_oy, out_ox, Hfull, Wfull;
  int wb_shift, tiles_x, tiles_img;   // log2(Wb); row tiles per image row / per image (Wb, Hb are powers of two)
  // fused pixel head (last f8 decoder layer): head_out[img, c, pix] = tanh(head_b[c] + sum_n relu(result[row, n]) * head_w[c, n])
  const float* head_w;
  const float* head_b;
  float* head_out;
  int64_t head_img_stride;
  int head_cout;
  // tiles are handed to a CTA (pair) in groups of `group` consecutive tile indices (= all N tiles of one row tile when the
  // pixel head accumulates across them); 1 = plain round-robin
  int group;
  // halo mode (tc_conv_halo_kernel): taps = KH*KW, halo patch pitch in pixels, bytes of one plane / of the whole patch
  int taps, halo_w, a_plane_bytes, a_tx_bytes;
  // passes = 3: fp32-grade product (hi*hi + lo*hi + hi*lo).  passes = 1 (halo kernel only): hi*hi alone -- plain fp16 operands
  // with fp32 accumulation, for layers whose result feeds no token (the last VQ-VAE decoder block: DESIGN.md "decoder
  // precision budget"); only the hi planes are fetched (w_tx_bytes / a_tx_bytes halve) and the corr accumulator is never read.
  int passes, w_tx_bytes;
  // fused LayerNorm of the rows this GEMM completes (mage_gemm_tc_ln): the CTA whose tile is the LAST of a 128-row block to become
  // globally visible (ln_count[row block] reaches n_tiles) normalises those rows of `out` into ln_split (the next GEMM's operand)
  const float* ln_gamma;
  const float* ln_beta;
  __half* ln_split;
  int64_t ln_plane;
  int* ln_count;
  float ln_eps;
  // fused QKV projection + axial attention (mage_qkv_axial_attn_tc): the N tile holds [q|k|v] x 32 columns of two heads, the
  // epilogue runs softmax(q k^T * attn_scale) v over the 16 positions of each line and stores only the attention output
  int attn;
  float attn_scale;
  // halo kernel shared-memory plan of this launch (HaloCfg): A ring depth / stage bytes, W slots / slot bytes, resident weights
  int sa, sw, a_stage_bytes, w_slot_bytes, w_resident, smem_bytes;
};

template <int BN, int CG>
struct Cfg {
  static constexpr int A_BYTES = 2 * BM * BK * 2;   // hi + lo planes of this CTA's 128 rows
  static constexpr int W_ROWS = BN / CG;            // W rows this CTA loads (the pair splits the N tile)
  static constexpr int W_BYTES = 2 * W_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int STAGING_BYTES = 8 * 32 * 32 * 4;   // one XOR-swizzled 32x32 fp32 transpose tile per epilogue warp
  static constexpr int STAGES_RAW = (SMEM_BUDGET - 1024 - STAGING_BYTES - 256) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int ACC_COLS = kOneAcc ? BN : 2 * BN;           // main+corr per accumulator stage
  static constexpr int ACC_STAGES = (ACC_COLS * 2 <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + 256;
};

// i-th tile of unit `unit`: groups of G consecutive tile indices are dealt round-robin to the units
__device__ __forceinline__ int tile_of(int i, int unit, int n_units, int G) {
  return G == 1 ? unit + i * n_units : (unit + (i / G) * n_units) * G + i % G;
}

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if (ACT == MAGE_ACT_RELU) return fmaxf(x, 0.f);
  // x * sigmoid(1.702 x) with ex2.approx / rcp.approx (a few ulp, branch-free: the epilogue is latency-bound)
  if (ACT == MAGE_ACT_QUICKGELU) return __fdividef(x, 1.f + __expf(-1.702f * x));
  if (ACT == MAGE_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  if (ACT == MAGE_ACT_TANH) return tanhf(x);
  return x;
}

// Epilogue of one CTA: 8 warps, two per TMEM lane quadrant (warp % 4), the pair splitting the 32-column chunks.
// Row-per-thread: lane l owns output row quad*32 + l (its TMEM lane).  Per 32-column chunk: tcgen05.ld main + corr ->
// v = main + corr*2^-11 -> bias (uniform loads), activation, residual (the thread's own 128-byte row segment) -> the warp's
// swizzled 4 KB staging tile -> ONE TMA store per output kind (the fp32 tile, or the hi+lo planes of the split format).
// No per-row global address arithmetic and no transposes: the k loops on this path are short (K = 512..1152), so the
// epilogue's instruction count is what paces the kernel (ncu: ~2200 instructions per warp and tile before this form).
// Rows beyond M are clipped by TMA.
// NS ("N-split", BN = 256 pair tiles): the tile's two 128-column halves are separate accumulators [main_h | corr_h] with their
// own full/empty barriers; epilogue warp (quadrant, hw) drains half hw, so the MMA issuer can start the next tile's half 0
// while half 1 is still being drained -- a 256-wide tile (A fetched once per 256 output columns) without giving up overlap.
// LayerNorm of NR rows (C = 512) by one warp, the math of layernorm_kernel<4> (misc.cu) -- two-pass statistics in registers, same
// element-to-lane layout and summation order, so the result is bit-identical to the separate kernel's.  The rows were written by
// OTHER CTAs' TMA stores, which this SM's L1 may not have seen: ld.global.cg.  NR rows are loaded before any is reduced, so the
// L2 round trips overlap (one row at a time would put 16 dependent round trips on the tail of the GEMM).
template <int NR>
__device__ __forceinline__ bool ln_rows_512(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            __half* __restrict__ split, int64_t plane, int64_t row0, int M, float eps, int lane) {
  constexpr int C = 512, NV = 4;
  float4 v[NR][NV];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int64_t row = row0 + r < M ? row0 + r : M - 1;
    const float4* src = reinterpret_cast<const float4*>(x + row * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[r][i] = __ldcg(src + i * 32 + lane);
  }
  bool bad = false;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[r][i].x - mean, b = v[r][i].y - mean, c = v[r][i].z - mean, d = v[r][i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / C) + eps);
    if (row0 + r < M) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
        float4 o;
        o.x = (v[r][i].x - mean) * rstd * g.x + b.x;
        o.y = (v[r][i].y - mean) * rstd * g.y + b.y;
        o.z = (v[r][i].z - mean) * rstd * g.z + b.z;
        o.w = (v[r][i].w - mean) * rstd * g.w + b.w;
        uint2 hi, lo;
        bad |= split4(o, hi, lo);
        const int64_t e = (row0 + r) * C + (i * 32 + lane) * 4;
        *reinterpret_cast<uint2*>(split + e) = hi;
        *reinterpret_cast<uint2*>(split + plane + e) = lo;
      }
    }
  }
  return bad;
}

template <int BN, int CG, int ACT, bool HEAD = false, bool NS = false>
__device__ __forceinline__ void epilogue_loop(const TcParams& p, const CUtensorMap* mapO, const CUtensorMap* mapS,
                                              const CUtensorMap* mapR, uint32_t tmem_base, uint8_t* staging, uint32_t tfull0,
                                              uint32_t tempty0, volatile int* s_flag = nullptr) {
  using C = Cfg<BN, CG>;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;   // a unit = the CTA (pair) that owns a tile
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int quad = warp & 3, half = (warp - 2) >> 2;
  uint8_t* const stg = staging + (warp - 2) * 4096;
  const uint32_t stg_s = smem_u32(stg);
  const bool res_relu = (p.act & MAGE_RES_RELU) != 0, post = (p.act & MAGE_ACT_POST_RES) != 0;
  const bool single = kOneAcc || p.passes == 1;   // no corr accumulator to fold in
  const float* const __restrict__ bias = p.bias;
  const float* const __restrict__ res = p.res;
  const bool has_out = p.out != nullptr, has_split = p.split != nullptr, has_relu = p.split_relu != nullptr;
  const int n_tiles = p.n_tiles, M = p.M;
  const int num_tiles = (p.m_tiles / CG) * p.n_tiles;
  // staging addresses of this lane's row: fp32 tile rows are 128 B (SWIZZLE_128B: 16-byte chunk ^ (row & 7)), split planes
  // have 64-byte rows (SWIZZLE_64B: chunk ^ ((row >> 1) & 3)), lo plane 2 KB after the hi plane
  const uint32_t row32 = stg_s + lane * 128, sw32 = lane & 7;
  const uint32_t row16 = stg_s + lane * 64, sw16 = (lane >> 1) & 3;
  int tcount = 0;
  float amax = 0.f;
  bool pending = false;   // a TMA store of this warp may still be reading the staging tile
  // Fused LayerNorm (p.ln_count): called one tile LATE (at the top of the next tile, and once after the loop), when the stores of
  // row block `mt_done` issued by this CTA have long been in flight: wait for their completion, make them visible device-wide,
  // bump the row block's counter; whichever CTA brings it to n_tiles has every column of those 128 rows visible and normalises
  // them (8 warps x 16 rows).  Which CTA that is does not matter for the result: LayerNorm is per row, the order inside a row fixed.
  bool ln_bad = false;
  auto ln_finish = [&](int mt_done) {
    if (mt_done * BM >= p.M) return;   // CTA-uniform (the padding row tile of an odd pair)
    if (lane == 0) bulk_wait0();
    __syncwarp();
    pending = false;
    fence_proxy_async();
    __threadfence();
    asm volatile("bar.sync 6, 256;" ::: "memory");
    if (warp == 2 && lane == 0) {
      const int old = atomicAdd(p.ln_count + mt_done, 1);
      const int last = old == p.n_tiles - 1;
      if (last) p.ln_count[mt_done] = 0;   // self-resetting: every contribution of this launch has been counted
      *s_flag = last;
    }
    asm volatile("bar.sync 6, 256;" ::: "memory");
    if (*s_flag) {
      __threadfence();
      const int r0 = mt_done * BM + (warp - 2) * 16;
#pragma unroll 1
      for (int r = r0; r < r0 + 16 && r < p.M; r += 4)
        ln_bad |= ln_rows_512<4>(p.out, p.ln_gamma, p.ln_beta, p.ln_split, p.ln_plane, r, p.M, p.ln_eps, lane);
    }
  };
  int mt_prev = -1;
  // HEAD (fused pixel head): nothing is staged for a store, so the staging region holds (a) the eight warps' 512-byte slots of
  // the final two-warp combine and (b) a [256][4] table (head_w[0][n], head_w[1][n], head_w[2][n], bias[n]): per column ONE
  // broadcast 16-byte shared-memory read replaces the bias load and three global weight loads, whose latency sat on the
  // epilogue's dependent chain (ncu, profiles/r02j_*: ~15 % of the stall samples on the first FFMA of each output channel).
  uint8_t* const hslot = staging + (warp - 2) * 512;
  const float4* const htab = reinterpret_cast<const float4*>(staging + 4096);
  if (HEAD) {
    const int n = (warp - 2) * 32 + lane;   // 256 epilogue threads = the 256 conv channels
    float4 t;
    t.x = p.head_cout > 0 ? __ldg(p.head_w + n) : 0.f;
    t.y = p.head_cout > 1 ? __ldg(p.head_w + p.N + n) : 0.f;
    t.z = p.head_cout > 2 ? __ldg(p.head_w + 2 * p.N + n) : 0.f;
    t.w = bias ? __ldg(bias + n) : 0.f;
    reinterpret_cast<float4*>(staging + 4096)[n] = t;
    asm volatile("bar.sync 5, 256;" ::: "memory");
  }
  auto staging_free = [&]() {
    if (pending) {
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      pending = false;
    }
  };
  float hacc[3] = {0.f, 0.f, 0.f};
  for (;; ++tcount) {
    const int tile = tile_of(tcount, unit, n_units, p.group);
    if (tile >= num_tiles) break;
    const int mt = (tile / n_tiles) * CG + cta_rank, nt = tile % n_tiles;
    const int acc = NS ? half : tcount % C::ACC_STAGES;
    const uint32_t aph = NS ? (tcount & 1) : (tcount / C::ACC_STAGES) & 1;
    if (!HEAD && p.ln_count && mt_prev >= 0) ln_finish(mt_prev);
    mt_prev = mt;
    // the thread's row, the warp's store-box origin and the residual offset: once per tile
    const int r = quad * 32 + lane, m = mt * BM + r;
    int sx0 = mt * BM + quad * 32, sy0 = 0, simg = 0, oy = 0, ox = 0;
    int64_t res_off = -1;
    if (p.conv) {
      simg = mt / p.tiles_img;
      const int rr = mt - simg * p.tiles_img;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int r0 = quad * 32;
      sx0 = tx * p.Wb + (r0 & (p.Wb - 1));
      sy0 = ty * p.Hb + (r0 >> p.wb_shift);
      oy = ty * p.Hb + (r >> p.wb_shift);
      ox = tx * p.Wb + (r & (p.Wb - 1));
      if (p.res_mode == 1) res_off = (((int64_t)simg * p.Hout + oy) * p.Wout + ox) * p.ldr;
      else if (p.res_mode == 2) res_off = (((int64_t)simg * (p.Hout >> 1) + (oy >> 1)) * (p.Wout >> 1) + (ox >> 1)) * p.ldr;
      else if (p.res_mode == 3) res_off = ((int64_t)oy * p.Wout + ox) * p.ldr;
    } else if (res && m < M) {
      res_off = (int64_t)(p.res_mod > 0 ? m % p.res_mod : m) * p.ldr;
    }
    if (HEAD && tcount % p.group == 0) hacc[0] = hacc[1] = hacc[2] = 0.f;
    // Residual row segments are requested EARLY -- the first chunk's before the wait for the MMAs, the following ones at the end of
    // the previous chunk -- so that their L2 latency is off the epilogue's critical path (ncu, profiles/r02j_*: 20-25 % of the
    // kernel's stall samples sat on the first use of the residual when it was loaded inside its own chunk).
    constexpr int c_first = 0, c_step = NS ? 1 : 2;
    const int c_begin = NS ? half * 4 : half, c_end = NS ? half * 4 + 4 : BN / 32;
    (void)c_first;
    float4 rv[8];
    auto load_res = [&](int c, float4 (&dst)[8]) {
#ifndef MAGE_EXP_NO_RES
      if (res_off >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = __ldg(reinterpret_cast<const float4*>(res + res_off + nt * BN + c * 32) + j);
      }
#endif
    };
    load_res(c_begin, rv);
    mbar_wait(tfull0 + 8u * acc, aph);
    tc_fence_after();
    const uint32_t t_main = tmem_base + ((uint32_t)(quad * 32) << 16) + (NS ? half * 256 : acc * C::ACC_COLS);
    const uint32_t t_corr = t_main + (NS ? 128 : kOneAcc ? 0 : BN);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += c_step) {
      const int n = nt * BN + c * 32;
      const int tc_col = NS ? (c & 3) * 32 : c * 32;
      uint32_t rm[32], rc[32];
      tmem_ld32(t_main + tc_col, rm);
      if (!single) tmem_ld32(t_corr + tc_col, rc);
      tmem_wait_ld();
      if (NS ? (c & 3) == 3 : c + 2 >= BN / 32) {
        // this warp's last chunk is in registers: hand the TMEM accumulator stage back before the arithmetic
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(mapa_u32(tempty0 + 8u * acc, 0));   // the leader's MMA thread waits for both CTAs
          else mbar_arrive(tempty0 + 8u * acc);
        }
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = single ? __uint_as_float(rm[j]) : fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(rm[j]));
#ifdef MAGE_EXP_NO_BIAS
      if (false) {
#else
      if (bias && !HEAD) {   // HEAD: the bias comes out of the shared-memory table together with the head weights
#endif
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n) + j);   // same address in every lane: one broadcast
          v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      }
      if (ACT != MAGE_ACT_NONE && !post) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_fn<ACT>(v[j]);
      }
#ifdef MAGE_EXP_NO_RES
      if (false) {
#else
      if (res_off >= 0) {
#endif
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 a = rv[j];
          if (res_relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
          v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
        }
      }
      if (ACT != MAGE_ACT_NONE && post) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_fn<ACT>(v[j]);
      }
      if (HEAD) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;   // two chains per output channel
        const float4* tab = htab + n;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float4 ta = tab[j], tb = tab[j + 1];   // same address in every lane: broadcast
          const float ra = fmaxf(v[j] + ta.w, 0.f), rb = fmaxf(v[j + 1] + tb.w, 0.f);
          s0 = fmaf(ra, ta.x, s0); s1 = fmaf(ra, ta.y, s1); s2 = fmaf(ra, ta.z, s2);
          u0 = fmaf(rb, tb.x, u0); u1 = fmaf(rb, tb.y, u1); u2 = fmaf(rb, tb.z, u2);
        }
        hacc[0] += s0 + u0; hacc[1] += s1 + u1; hacc[2] += s2 + u2;
      }
#ifdef MAGE_EXP_NO_STORE
      if (has_out && v[0] == 123456.789f) {
#else
      if (has_out) {
#endif
        staging_free();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(row32 + ((j ^ sw32) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                 __float_as_uint(v[4 * j + 3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {   // bulk-group bookkeeping is per thread: always the same lane issues and waits
          tma_store_5d(mapO, stg_s, n, sx0, sy0, simg, 0);
          bulk_commit();
        }
        pending = true;
      }
#ifdef MAGE_EXP_NO_STORE
      if ((has_split || has_relu) && v[0] == 123456.789f) {
#else
      if (has_split || has_relu) {
#endif
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 0 ? !has_split : !has_relu) continue;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = v[2 * j], b = v[2 * j + 1];
            if (pass == 1) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            if (pass == 0 || !has_split) amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
            const __half2 h = __floats2half2_rn(a, b);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn((a - f.x) * kLoScale, (b - f.y) * kLoScale);
            hi[j] = *reinterpret_cast<const uint32_t*>(&h);
            lo[j] = *reinterpret_cast<const uint32_t*>(&l);
          }
          staging_free();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sts128(row16 + ((j ^ sw16) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            sts128(row16 + 2048 + ((j ^ sw16) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(pass == 0 ? mapS : mapR, stg_s, n, sx0, sy0, simg, 0);
            bulk_commit();
          }
          pending = true;
        }
      }
      // next chunk's residual: issued as soon as this chunk's values have left the registers, in flight during the staging
      // hand-off and the next chunk's TMEM loads.  (A register double buffer -- loading it at the TOP of this chunk -- spills at
      // the 168-register cap of a 320-thread CTA and measured slower: profiles/r02k_residual_prefetch_ab.txt.)
      if (c + c_step < c_end) load_res(c + c_step, rv);
    }
    if (HEAD && tcount % p.group == p.group - 1) {
      // every thread holds its row's partial sums over this warp's half of the columns; the partner warp (same quadrant,
      // other half: warp + 4) adds its half through the staging tile, then tanh and the planar pixel stores
      *reinterpret_cast<float4*>(hslot + lane * 16) = make_float4(hacc[0], hacc[1], hacc[2], 0.f);
      __syncwarp();
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      if (half == 0 && m < M) {
        const float4 a = *reinterpret_cast<const float4*>(hslot + lane * 16);
        const float4 b = *reinterpret_cast<const float4*>(hslot + 4 * 512 + lane * 16);   // warp + 4's slot
        const int64_t pix = (int64_t)(oy * p.out_sy + p.out_oy) * p.Wfull + (ox * p.out_sx + p.out_ox);
        const int64_t plane = (int64_t)p.Hfull * p.Wfull;
        float* dst = p.head_out + (int64_t)simg * p.head_img_stride + pix;
        const float sum[3] = {a.x + b.x, a.y + b.y, a.z + b.z};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          if (ch < p.head_cout) dst[ch * plane] = tanhf(sum[ch] + __ldg(p.head_b + ch));
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");   // the partner may reuse its slot
    }
  }
  if (!HEAD && p.ln_count && mt_prev >= 0) ln_finish(mt_prev);
  if (ln_bad && p.flag) atomicOr(p.flag, 1);
#ifdef MAGE_EXP_WAIT_READ   // experiment: only wait until the TMA engine has READ the staging tile (global visibility at grid end)
  if (lane == 0) bulk_wait_read0();
#else
  if (lane == 0) bulk_wait0();   // all of this warp's stores have landed before the CTA may exit
#endif
  if (!(amax <= 65504.f) && p.flag) atomicOr(p.flag, 1);
}

// Epilogue of the fused QKV-projection + axial-attention kernel (AxialAttentionBlock.attention for the H / W blocks,
// mage_model.py:31-33 with the permutes of :36-47 expressed in the tensor maps).  The A tile is 128 rows = 8 lines x 16
// positions of the attended axis (gathered in that order by the A map), the N tile (192 columns) is [q|k|v] x 32 of two heads
// (weights permuted at load).  Warp (quad, half) owns rows quad*32 .. +31 (two whole lines) and head `half` of the tile:
//   tcgen05.ld q, k, v (main + corr*2^-11 + bias) -> the TMEM stage is handed back at once (the MMAs of the next tile overlap
//   the attention math) -> k through the warp's 4 KB staging tile: s[j] = q . k_j over the 16 rows of the thread's own line
//   (the two lines of a warp read two addresses per instruction: broadcasts, no conflicts) -> softmax -> v through the same
//   tile: o = sum_j p[j] v_j -> split (hi/lo) -> one TMA store of the 32x32 output block.  The [rows, 1536] QKV tensor never
//   exists in memory.
template <int BN, int CG>
__device__ __forceinline__ void attn_epilogue_loop(const TcParams& p, const CUtensorMap* mapS, uint32_t tmem_base, uint8_t* staging,
                                                   uint32_t tfull0, uint32_t tempty0) {
  static_assert(BN == 192, "two heads x (q,k,v) x 32 columns");
  using C = Cfg<BN, CG>;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int quad = warp & 3, half = (warp - 2) >> 2;
  uint8_t* const stg = staging + (warp - 2) * 4096;
  const uint32_t stg_s = smem_u32(stg);
  float* const stg_f = reinterpret_cast<float*>(stg);
  const float* const __restrict__ bias = p.bias;
  const int n_tiles = p.n_tiles;
  const int num_tiles = (p.m_tiles / CG) * p.n_tiles;
  const uint32_t row16 = stg_s + lane * 64, sw16 = (lane >> 1) & 3;
  const int line0 = lane & 16;   // first lane of this thread's line
  float amax = 0.f;
  bool pending = false;
  for (int tcount = 0;; ++tcount) {
    const int tile = tile_of(tcount, unit, n_units, 1);
    if (tile >= num_tiles) break;
    const int mt = (tile / n_tiles) * CG + cta_rank, nt = tile % n_tiles;
    const int acc = tcount % C::ACC_STAGES;
    const uint32_t aph = (tcount / C::ACC_STAGES) & 1;
    const int simg = mt / p.tiles_img;
    const int rr = mt - simg * p.tiles_img;
    const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
    const int r0 = quad * 32;
    const int sx0 = tx * p.Wb + (r0 & (p.Wb - 1)), sy0 = ty * p.Hb + (r0 >> p.wb_shift);
    mbar_wait(tfull0 + 8u * acc, aph);
    tc_fence_after();
    const uint32_t t_main = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * C::ACC_COLS + half * 96;
    const uint32_t t_corr = t_main + BN;
    if (pending) {   // the previous tile's TMA store may still be reading the staging tile
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      pending = false;
    }
    // staging tile as a [32 rows][32 floats] exchange buffer: row = lane, 16-byte chunk c of a row sits at chunk c ^ (row & 7)
    // (conflict-free 128-bit stores; a line's 16 lanes read one row per instruction = a broadcast)
    auto load_part = [&](int part, float (&dst)[32]) {
      uint32_t rm[32], rc[32];
      tmem_ld32(t_main + part * 32, rm);
      tmem_ld32(t_corr + part * 32, rc);
      tmem_wait_ld();
      const float* b = bias + nt * BN + half * 96 + part * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + j);
        dst[4 * j] = fmaf(__uint_as_float(rc[4 * j]), kLoInv, __uint_as_float(rm[4 * j])) + bb.x;
        dst[4 * j + 1] = fmaf(__uint_as_float(rc[4 * j + 1]), kLoInv, __uint_as_float(rm[4 * j + 1])) + bb.y;
        dst[4 * j + 2] = fmaf(__uint_as_float(rc[4 * j + 2]), kLoInv, __uint_as_float(rm[4 * j + 2])) + bb.z;
        dst[4 * j + 3] = fmaf(__uint_as_float(rc[4 * j + 3]), kLoInv, __uint_as_float(rm[4 * j + 3])) + bb.w;
      }
    };
    auto to_smem = [&](const float (&src)[32]) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg_f + lane * 32 + 4 * (j ^ (lane & 7))) = make_float4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]);
    };
    float q[32], v[32];
    {
      float k[32];
      load_part(1, k);
      to_smem(k);
    }
    load_part(0, q);
    load_part(2, v);
    // q, k, v have left TMEM: hand the stage back before the attention math (the next tile's MMAs overlap it)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CG == 2) mbar_arrive_cluster(mapa_u32(tempty0 + 8u * acc, 0));
      else mbar_arrive(tempty0 + 8u * acc);
    }
    // ---- scores over the thread's line (the k rows of the warp's two lines are in the exchange buffer)
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(stg_f + (line0 + j) * 32);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 kv = kr[i ^ (j & 7)];
        d = fmaf(q[4 * i], kv.x, d); d = fmaf(q[4 * i + 1], kv.y, d);
        d = fmaf(q[4 * i + 2], kv.z, d); d = fmaf(q[4 * i + 3], kv.w, d);
      }
      s[j] = d * p.attn_scale;
    }
    float mx = s[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) mx = fmaxf(mx, s[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.f / sum;
    __syncwarp();   // everyone is done with k
    to_smem(v);
    __syncwarp();
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4* vr = reinterpret_cast<const float4*>(stg_f + (line0 + j) * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 vv = vr[i ^ (j & 7)];
        o[4 * i] = fmaf(s[j], vv.x, o[4 * i]); o[4 * i + 1] = fmaf(s[j], vv.y, o[4 * i + 1]);
        o[4 * i + 2] = fmaf(s[j], vv.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(s[j], vv.w, o[4 * i + 3]);
      }
    }
    __syncwarp();   // everyone is done with v: the tile becomes the output staging (hi plane 2 KB | lo plane 2 KB, SWIZZLE_64B)
    {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a = o[2 * j] * inv, b = o[2 * j + 1] * inv;
        amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
        const __half2 h = __floats2half2_rn(a, b);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn((a - f.x) * kLoScale, (b - f.y) * kLoScale);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sts128(row16 + ((j ^ sw16) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        sts128(row16 + 2048 + ((j ^ sw16) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_5d(mapS, stg_s, (nt * 2 + half) * 32, sx0, sy0, simg, 0);   // output column = head * 32
        bulk_commit();
      }
      pending = true;
    }
  }
#ifdef MAGE_EXP_WAIT_READ
  if (lane == 0) bulk_wait_read0();
#else
  if (lane == 0) bulk_wait0();
#endif
  if (!(amax <= 65504.f) && p.flag) atomicOr(p.flag, 1);
}

template <int BN, int CG, bool NS = false>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
               const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapS,
               const __grid_constant__ CUtensorMap mapR, const TcParams p) {
  using C = Cfg<BN, CG>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* staging = base_ptr + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES + C::STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + C::STAGES * C::STAGE_BYTES + C::STAGING_BYTES +
                                                                      8 * (2 * C::STAGES + 4));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;
  const int num_tiles = (p.m_tiles / CG) * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    if (p.out) tma_prefetch_desc(&mapO);
    if (p.split) tma_prefetch_desc(&mapS);
    if (p.split_relu) tma_prefetch_desc(&mapR);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), (NS ? 4 : 8) * CG);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::TMEM_COLS);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything is signalled remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // Programmatic dependent launch: this grid may have started while its predecessor is still running.  Everything that does not
  // depend on the predecessor -- barriers, TMEM, descriptor prefetch above, and the WEIGHT tiles of the first pipeline stages
  // below -- happens before griddepcontrol.wait; activations (A tiles, residuals) only after it.

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA loads its A rows + its W rows)
    int pre = 0;   // stages of the first tile whose W tile (and expect_tx) were issued ahead of the dependency wait
    if (!NS) {
      const int tile0 = tile_of(0, unit, n_units, p.group);
      if (tile0 < num_tiles) {
        pre = p.k_iters < C::STAGES ? p.k_iters : C::STAGES;
        const int w_row0 = (tile0 % p.n_tiles) * BN + cta_rank * C::W_ROWS;
        for (int it = 0; it < pre; ++it) {
          const uint32_t sa = base + it * C::STAGE_BYTES;
          if (CG == 2) {
            const uint32_t fb = mapa_u32(full_bar(it), 0);
            if (elect_one()) {
              if (cta_rank == 0) mbar_expect_tx(full_bar(it), 2 * C::STAGE_BYTES);
              tma_load_3d_2sm(sa + C::A_BYTES, &mapW, fb, it * BK, w_row0, 0);
            }
          } else if (elect_one()) {
            mbar_expect_tx(full_bar(it), C::STAGE_BYTES);
            tma_load_3d(sa + C::A_BYTES, &mapW, full_bar(it), it * BK, w_row0, 0);
          }
          __syncwarp();
        }
      }
    }
    pdl_wait();
    {
      int itg = 0;
      for (int ti = 0;; ++ti) {
        const int tile = tile_of(ti, unit, n_units, p.group);
        if (tile >= num_tiles) break;
        const int mt = (tile / p.n_tiles) * CG + cta_rank, nt = tile % p.n_tiles;
        int c1 = mt * BM, c2 = 0, c3 = 0;
        if (p.conv) {
          const int img = mt / p.tiles_img, r = mt - img * p.tiles_img;
          const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
          c1 = tx * p.Wb - p.pad_x;
          c2 = ty * p.Hb - p.pad_y;
          c3 = img;
        }
        const int w_row = NS ? nt * BN + cta_rank * 64 : nt * BN + cta_rank * C::W_ROWS;   // NS: + h*128 per half
        for (int it = 0; it < p.k_iters; ++it, ++itg) {
          const int s = itg % C::STAGES;
          const uint32_t ph = (itg / C::STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = base + s * C::STAGE_BYTES;
          int a0 = it * BK, a1 = c1, a2 = c2;
          if (p.conv) {
            const int tap = it / p.cin_blocks, cb = it - tap * p.cin_blocks;
            const int ky = tap / p.KW, kx = tap - ky * p.KW;
            a0 = cb * BK; a1 = c1 + kx; a2 = c2 + ky;
          }
          const bool w_done = ti == 0 && it < pre;   // expect_tx + W tile of this stage went out before the dependency wait
          if (CG == 2) {
            // both CTAs' bytes land on the LEADER's full barrier: the MMA thread there consumes both halves
            const uint32_t fb = mapa_u32(full_bar(s), 0);
            if (elect_one()) {
              if (cta_rank == 0 && !w_done) mbar_expect_tx(full_bar(s), 2 * C::STAGE_BYTES);
              tma_load_5d_2sm(sa, &mapA, fb, a0, a1, a2, c3, 0);
              if (NS) {   // two 64-row boxes: this CTA's share of each 128-column half
                tma_load_3d_2sm(sa + C::A_BYTES, &mapW, fb, it * BK, w_row, 0);
                tma_load_3d_2sm(sa + C::A_BYTES + C::W_BYTES / 2, &mapW, fb, it * BK, w_row + 128, 0);
              } else if (!w_done) {
                tma_load_3d_2sm(sa + C::A_BYTES, &mapW, fb, it * BK, w_row, 0);
              }
            }
          } else if (elect_one()) {
            if (!w_done) mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
            tma_load_5d(sa, &mapA, full_bar(s), a0, a1, a2, c3, 0);
            if (!w_done) tma_load_3d(sa + C::A_BYTES, &mapW, full_bar(s), it * BK, w_row, 0);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (the leader CTA; one elected lane issues)
    pdl_wait();
    if (cta_rank == 0 && NS) {
      constexpr uint32_t idesc = umma_idesc_f16(128, BM * CG);
      int itg = 0, tcount = 0;
      for (;; ++tcount) {
        if (tile_of(tcount, unit, n_units, p.group) >= num_tiles) break;
        const uint32_t tph = tcount & 1;
        for (int it = 0; it < p.k_iters; ++it, ++itg) {
          const int s = itg % C::STAGES;
          const uint32_t ph = (itg / C::STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = base + s * C::STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + BM * BK * 2);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (it == 0) {   // half h of the previous tile has been drained by both CTAs' epilogue warps
              mbar_wait(tempty_bar(h), tph ^ 1);
              tc_fence_after();
            }
            const uint32_t wb = sa + C::A_BYTES + h * (C::W_BYTES / 2);
            const uint64_t w_hi = umma_desc_sw128(wb), w_lo = umma_desc_sw128(wb + 64 * BK * 2);
            const uint32_t d_main = tmem_base + h * 256, d_corr = d_main + 128;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / UK; ++k) {
                const uint64_t adv = (uint64_t)((k * UK * 2) >> 4);
                const uint32_t accum = (it > 0 || k > 0) ? 1u : 0u;
                umma_f16_2sm(d_corr, a_lo + adv, w_hi + adv, idesc, accum);
                umma_f16_2sm(d_corr, a_hi + adv, w_lo + adv, idesc, 1u);
                umma_f16_2sm(d_main, a_hi + adv, w_hi + adv, idesc, accum);
              }
              if (h == 1) umma_commit_2sm(empty_bar(s), 3);
              if (it + 1 == p.k_iters) umma_commit_2sm(tfull_bar(h), 3);
            }
            __syncwarp();
          }
        }
      }
    } else if (cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN, BM * CG);
      int itg = 0, tcount = 0;
      for (;; ++tcount) {
        if (tile_of(tcount, unit, n_units, p.group) >= num_tiles) break;
        const int acc = tcount % C::ACC_STAGES;
        const uint32_t aph = (tcount / C::ACC_STAGES) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + acc * C::ACC_COLS, d_corr = kOneAcc ? d_main : d_main + BN;
        for (int it = 0; it < p.k_iters; ++it, ++itg) {
          const int s = itg % C::STAGES;
          const uint32_t ph = (itg / C::STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = base + s * C::STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + BM * BK * 2);
          const uint64_t w_hi = umma_desc_sw128(sa + C::A_BYTES), w_lo = umma_desc_sw128(sa + C::A_BYTES + C::W_ROWS * BK * 2);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t adv = (uint64_t)((k * UK * 2) >> 4);  // +32 B per k-step inside the swizzle row
              const uint32_t accum = (it > 0 || k > 0) ? 1u : 0u;
              if (CG == 2) {
                umma_f16_2sm(d_corr, a_lo + adv, w_hi + adv, idesc, accum);
                umma_f16_2sm(d_corr, a_hi + adv, w_lo + adv, idesc, 1u);
                umma_f16_2sm(d_main, a_hi + adv, w_hi + adv, idesc, kOneAcc ? 1u : accum);
              } else {
                umma_f16(d_corr, a_lo + adv, w_hi + adv, idesc, accum);
                umma_f16(d_corr, a_hi + adv, w_lo + adv, idesc, 1u);
                umma_f16(d_main, a_hi + adv, w_hi + adv, idesc, kOneAcc ? 1u : accum);
              }
            }
            if (CG == 2) umma_commit_2sm(empty_bar(s), 3);   // frees the stage in both CTAs
            else umma_commit(empty_bar(s));
            if (it + 1 == p.k_iters) {
              if (CG == 2) umma_commit_2sm(tfull_bar(acc), 3);   // both CTAs' epilogues drain their own TMEM
              else umma_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    pdl_wait();   // residuals are the predecessor's output
    if (BN == 192 && p.attn) {
      if constexpr (BN == 192) attn_epilogue_loop<BN, CG>(p, &mapS, tmem_base, staging, tfull_bar(0), tempty_bar(0));
    } else if ((BN == 256 || BN == 128) && p.head_w) {
      if constexpr (BN == 256 || BN == 128) epilogue_loop<BN, CG, MAGE_ACT_NONE, true>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0));
    } else
    switch (p.act & 0xff) {
      case MAGE_ACT_NONE: epilogue_loop<BN, CG, MAGE_ACT_NONE, false, NS>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0),
                                                                          reinterpret_cast<volatile int*>(tmem_slot + 1)); break;
      case MAGE_ACT_RELU: epilogue_loop<BN, CG, MAGE_ACT_RELU, false, NS>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0)); break;
      case MAGE_ACT_QUICKGELU: epilogue_loop<BN, CG, MAGE_ACT_QUICKGELU, false, NS>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0)); break;
      case MAGE_ACT_GELU: epilogue_loop<BN, CG, MAGE_ACT_GELU, false, NS>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0)); break;
      default: epilogue_loop<BN, CG, MAGE_ACT_TANH, false, NS>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0)); break;
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // no CTA of the pair leaves while the other may still signal its barriers / read its smem
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------ halo-reusing convolution
// KHxKW stride-1 convolution on 16x8-pixel output tiles.  Per 64-channel block ONE TMA box brings the (16+KH-1)x(8+KW-1)
// input patch (both planes); every tap's A operand is that patch read through a shifted descriptor (start + (ky*pitch+kx)
// rows of 128 B, 8-row groups `pitch` rows apart), so the activations are fetched once instead of KH*KW times.  Weight tiles
// stream per (tap, channel block) through their own ring.  Same roles / epilogue / CTA-pair scheme as tc_gemm_kernel.
constexpr int HALO_A_STAGE = 46080;   // 18 x 10 pixels x 128 B x 2 planes (3x3 taps), a multiple of 1024: the largest A stage
constexpr int HALO_SA_MAX = 4, HALO_SW_MAX = 24;   // barrier slots; the ring depths themselves are chosen per launch (TcParams)
constexpr int HALO_BAR_BYTES = 1024;               // 2*4 + 2*24 + 4 mbarriers + the TMEM slot

// Shared memory of the halo kernel, laid out per launch:  [ sa A stages | sw W slots | 32 KB epilogue staging | barriers ]
//   A stage = the halo patch of one 64-channel block (both planes, or the hi plane alone in single-pass mode), 1 KB aligned;
//   W slot  = one (tap, channel block) weight tile: this CTA's W_ROWS rows x 128 B per plane.
// Weight slots work in one of two ways (TcParams::w_resident):
//   ring      sw slots cycle through the (tap, block) tiles of every output tile -- weights are re-fetched per tile;
//   resident  every weight tile this CTA will ever need (taps x channel blocks x the N tiles of its group) has its own slot, is
//             fetched ONCE and stays: a 3x3 64->64 layer re-uses the same nine tiles for each of its ~55 output tiles per CTA,
//             and with a ring shallower than one tile's worth of taps the TMA latency of the refetch was exposed on every tile
//             (ncu, profiles/r02d_*: the MMA warp waiting on w_full, 2.6 us per tile for 0.6 us of MMAs).
template <int BN, int CG>
struct HaloCfg {
  static constexpr int W_ROWS = BN / CG;
  static constexpr int W_PLANE_BYTES = W_ROWS * BK * 2;
  static constexpr int STAGING_BYTES = 8 * 32 * 32 * 4;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_MAX = SMEM_BUDGET;
  static constexpr int FIXED_BYTES = 1024 + STAGING_BYTES + HALO_BAR_BYTES;   // alignment slack + staging + barriers
};

template <int BN, int CG>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                    const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapS,
                    const __grid_constant__ CUtensorMap mapR, const TcParams p) {
  using H = HaloCfg<BN, CG>;
  using C = Cfg<BN, CG>;
  const int SA = p.sa, SW = p.sw;                 // ring depths of this launch
  const uint32_t a_stage_bytes = (uint32_t)p.a_stage_bytes, w_slot_bytes = (uint32_t)p.w_slot_bytes;
  const uint32_t w_plane_off = (uint32_t)H::W_PLANE_BYTES;   // lo plane of a weight slot (3-pass mode)
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_base = base + SA * a_stage_bytes;
  const uint32_t STG_OFF = SA * a_stage_bytes + SW * w_slot_bytes;
  uint8_t* staging = base_ptr + STG_OFF;
  const uint32_t bar_base = base + STG_OFF + H::STAGING_BYTES;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (HALO_SA_MAX + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * HALO_SA_MAX + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * HALO_SA_MAX + HALO_SW_MAX + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * HALO_SA_MAX + 2 * HALO_SW_MAX + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * HALO_SA_MAX + 2 * HALO_SW_MAX + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STG_OFF + H::STAGING_BYTES +
                                                                      8 * (2 * HALO_SA_MAX + 2 * HALO_SW_MAX + 4));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;
  const int num_tiles = (p.m_tiles / CG) * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    if (p.out) tma_prefetch_desc(&mapO);
    if (p.split) tma_prefetch_desc(&mapS);
    if (p.split_relu) tma_prefetch_desc(&mapR);
    for (int s = 0; s < SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < SW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8 * CG); }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(smem_u32(const_cast<uint32_t*>(tmem_slot)), H::TMEM_COLS);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), H::TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // resident weights do not depend on the predecessor grid: all of them go out BEFORE griddepcontrol.wait (PDL), so the weight
    // fetch of this layer overlaps the tail of the previous one; activations only after the wait
    bool w_pre = false;
    if (p.w_resident) {
      for (int g = 0; g < p.group; ++g) {
        const int tile = tile_of(g, unit, n_units, p.group);
        if (tile >= num_tiles) break;
        const int w_row = (tile % p.n_tiles) * BN + cta_rank * H::W_ROWS;
        for (int cb = 0; cb < p.cin_blocks; ++cb)
          for (int tap = 0; tap < p.taps; ++tap) {
            const int s = (g * p.cin_blocks + cb) * p.taps + tap;
            const uint32_t dst = w_base + s * w_slot_bytes;
            const int k0 = (tap * p.cin_blocks + cb) * BK;
            if (CG == 2) {
              const uint32_t fb = mapa_u32(w_full(s), 0);
              if (elect_one()) {
                if (cta_rank == 0) mbar_expect_tx(w_full(s), 2 * p.w_tx_bytes);
                tma_load_3d_2sm(dst, &mapW, fb, k0, w_row, 0);
              }
            } else if (elect_one()) {
              mbar_expect_tx(w_full(s), p.w_tx_bytes);
              tma_load_3d(dst, &mapW, w_full(s), k0, w_row, 0);
            }
            __syncwarp();
          }
      }
      w_pre = true;
    }
    pdl_wait();
    {
      int ia = 0, iw = 0;
      for (int ti = 0;; ++ti) {
        const int tile = tile_of(ti, unit, n_units, p.group);
        if (tile >= num_tiles) break;
        const int mt = (tile / p.n_tiles) * CG + cta_rank, nt = tile % p.n_tiles;
        const int img = mt / p.tiles_img, r = mt - img * p.tiles_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int c1 = tx * p.Wb - p.pad_x, c2 = ty * p.Hb - p.pad_y;
        const int w_row = nt * BN + cta_rank * H::W_ROWS;
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          {
            const int s = ia % SA;
            mbar_wait(a_empty(s), ((ia / SA) & 1) ^ 1);
            const uint32_t dst = base + s * a_stage_bytes;
            if (CG == 2) {
              const uint32_t fb = mapa_u32(a_full(s), 0);
              if (elect_one()) {
                if (cta_rank == 0) mbar_expect_tx(a_full(s), 2 * p.a_tx_bytes);
                tma_load_5d_2sm(dst, &mapA, fb, cb * BK, c1, c2, img, 0);
              }
            } else if (elect_one()) {
              mbar_expect_tx(a_full(s), p.a_tx_bytes);
              tma_load_5d(dst, &mapA, a_full(s), cb * BK, c1, c2, img, 0);
            }
            __syncwarp();
            ++ia;
          }
          for (int tap = 0; tap < p.taps; ++tap, ++iw) {
            int s;
            if (p.w_resident) {
              // every weight tile of this CTA has its own slot and was fetched once, ahead of the dependency wait
              if (w_pre || ti >= p.group) continue;
              s = ((ti % p.group) * p.cin_blocks + cb) * p.taps + tap;
            } else {
              s = iw % SW;
              mbar_wait(w_empty(s), ((iw / SW) & 1) ^ 1);
            }
            const uint32_t dst = w_base + s * w_slot_bytes;
            const int k0 = (tap * p.cin_blocks + cb) * BK;
            if (CG == 2) {
              const uint32_t fb = mapa_u32(w_full(s), 0);
              if (elect_one()) {
                if (cta_rank == 0) mbar_expect_tx(w_full(s), 2 * p.w_tx_bytes);
                tma_load_3d_2sm(dst, &mapW, fb, k0, w_row, 0);
              }
            } else if (elect_one()) {
              mbar_expect_tx(w_full(s), p.w_tx_bytes);
              tma_load_3d(dst, &mapW, w_full(s), k0, w_row, 0);
            }
            __syncwarp();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA; one elected lane issues)
    pdl_wait();
    if (cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN, BM * CG);
      const uint32_t sbo = (uint32_t)p.halo_w * 128u;
      int ia = 0, iw = 0, tcount = 0;
      for (;; ++tcount) {
        if (tile_of(tcount, unit, n_units, p.group) >= num_tiles) break;
        const int acc = tcount % C::ACC_STAGES;
        const uint32_t aph = (tcount / C::ACC_STAGES) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + acc * C::ACC_COLS, d_corr = kOneAcc ? d_main : d_main + BN;
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++ia) {
          const int sa = ia % SA;
          mbar_wait(a_full(sa), (ia / SA) & 1);
          tc_fence_after();
          const uint32_t a_stage = base + sa * a_stage_bytes;
          int ky = 0, kx = 0;
          for (int tap = 0; tap < p.taps; ++tap, ++iw) {
            const int sw = p.w_resident ? ((tcount % p.group) * p.cin_blocks + cb) * p.taps + tap : iw % SW;
            mbar_wait(w_full(sw), p.w_resident ? 0u : (uint32_t)((iw / SW) & 1));   // resident: completes once, stays complete
            tc_fence_after();
            const uint32_t a_addr = a_stage + (uint32_t)(ky * p.halo_w + kx) * 128u;
            const uint64_t a_hi = umma_desc_sw128_sbo(a_addr, sbo), a_lo = umma_desc_sw128_sbo(a_addr + p.a_plane_bytes, sbo);
            const uint32_t w_addr = w_base + sw * w_slot_bytes;
            const uint64_t w_hi = umma_desc_sw128(w_addr), w_lo = umma_desc_sw128(w_addr + w_plane_off);
            const bool last_tap = tap + 1 == p.taps, last = last_tap && cb + 1 == p.cin_blocks;
            if (elect_one()) {
              if (p.passes == 1) {
#pragma unroll
                for (int k = 0; k < BK / UK; ++k) {
                  const uint64_t adv = (uint64_t)((k * UK * 2) >> 4);
                  const uint32_t accum = (cb > 0 || tap > 0 || k > 0) ? 1u : 0u;
                  if (CG == 2) umma_f16_2sm(d_main, a_hi + adv, w_hi + adv, idesc, accum);
                  else umma_f16(d_main, a_hi + adv, w_hi + adv, idesc, accum);
                }
              } else {
#pragma unroll
              for (int k = 0; k < BK / UK; ++k) {
                const uint64_t adv = (uint64_t)((k * UK * 2) >> 4);
                const uint32_t accum = (cb > 0 || tap > 0 || k > 0) ? 1u : 0u;
                if (CG == 2) {
                  umma_f16_2sm(d_corr, a_lo + adv, w_hi + adv, idesc, accum);
                  umma_f16_2sm(d_corr, a_hi + adv, w_lo + adv, idesc, 1u);
                  umma_f16_2sm(d_main, a_hi + adv, w_hi + adv, idesc, kOneAcc ? 1u : accum);
                } else {
                  umma_f16(d_corr, a_lo + adv, w_hi + adv, idesc, accum);
                  umma_f16(d_corr, a_hi + adv, w_lo + adv, idesc, 1u);
                  umma_f16(d_main, a_hi + adv, w_hi + adv, idesc, kOneAcc ? 1u : accum);
                }
              }
              }
              if (CG == 2) {
                if (!p.w_resident) umma_commit_2sm(w_empty(sw), 3);
                if (last_tap) umma_commit_2sm(a_empty(sa), 3);
                if (last) umma_commit_2sm(tfull_bar(acc), 3);
              } else {
                if (!p.w_resident) umma_commit(w_empty(sw));
                if (last_tap) umma_commit(a_empty(sa));
                if (last) umma_commit(tfull_bar(acc));
              }
            }
            __syncwarp();
            if (++kx == p.KW) { kx = 0; ++ky; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    pdl_wait();   // residuals are the predecessor's output
    if ((BN == 256 || BN == 128) && p.head_w) {
      if constexpr (BN == 256 || BN == 128) epilogue_loop<BN, CG, MAGE_ACT_NONE, true>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0));
    } else
    switch (p.act & 0xff) {
      case MAGE_ACT_NONE: epilogue_loop<BN, CG, MAGE_ACT_NONE>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0)); break;
      case MAGE_ACT_RELU: epilogue_loop<BN, CG, MAGE_ACT_RELU>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0)); break;
      default: epilogue_loop<BN, CG, MAGE_ACT_TANH>(p, &mapO, &mapS, &mapR, tmem_base, staging, tfull_bar(0), tempty_bar(0)); break;
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, H::TMEM_COLS);
    else tmem_dealloc(tmem_base, H::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ split conversion
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, int64_t ldx, __half* __restrict__ out,
                                                    int64_t plane, int rows, int C4, int relu, int* flag) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)rows * C4) return;
  const int row = (int)(t / C4), c = (int)(t - (int64_t)row * C4);
  float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)row * ldx) + c);
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  uint2 hi, lo;
  const bool bad = split4(v, hi, lo);
  const int64_t o = ((int64_t)row * C4 + c) * 4;
  *reinterpret_cast<uint2*>(out + o) = hi;
  *reinterpret_cast<uint2*>(out + plane + o) = lo;
  if (bad && flag) atomicOr(flag, 1);
}

// out[r, :] (both planes) = table[idx[r], :] (both planes): nn.Embedding on a pre-split table
__global__ void __launch_bounds__(256) embedding_split_kernel(const int64_t* __restrict__ idx, const __half* __restrict__ table,
                                                              int64_t table_plane, __half* __restrict__ out, int64_t out_plane,
                                                              int rows, int C8) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)rows * C8 * 2) return;
  const int pl = (int)(t / ((int64_t)rows * C8));
  const int64_t u = t - (int64_t)pl * rows * C8;
  const int row = (int)(u / C8), c = (int)(u - (int64_t)row * C8);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(table + pl * table_plane + idx[row] * (int64_t)C8 * 8) + c);
  reinterpret_cast<uint4*>(out + pl * out_plane)[u] = v;
}

// First-layer im2row: out(split)[n, y, x, kx*C + c] = in[n, c, y, x + kx - pad] (0 outside the image and for the padding
// channels up to 64), so that a KHxKW convolution of a C<=9-channel planar image becomes a KHx1 convolution with 64 input
// channels -- a tensor-core (halo) convolution with K = KH*64.  One thread per (pixel, 8-channel chunk).
__global__ void __launch_bounds__(256) patch_rows_split_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t plane,
                                                               int n_img, int C, int H, int W, int KW, int pad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_pix = (int64_t)n_img * H * W;
  if (t >= n_pix * 8) return;
  const int chunk = (int)(t & 7);
  const int64_t pix = t >> 3;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int q = chunk * 8 + j, kx = q / C, c = q - kx * C;
    const int xs = x + kx - pad;
    v[j] = (kx < KW && xs >= 0 && xs < W) ? __ldg(in + (((int64_t)n * C + c) * H + y) * W + xs) : 0.f;
  }
  uint2 h0, l0, h1, l1;
  split4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
  split4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
  const int64_t o = pix * 64 + chunk * 8;
  *reinterpret_cast<uint4*>(out + o) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(out + plane + o) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

// Space-to-depth with a one-pixel top/left pad, emitted in the split format: the input operand of a 4x4 stride-2 pad-1
// convolution rewritten as a 2x2 stride-1 VALID convolution (vqvae_model.py:175).  Output pixel (Y, X), channel block
// q = py*2+px holds input pixel (2Y+py-1, 2X+px-1) (zero outside the image), so output row oy of the strided convolution
// reads rows 2oy-1 .. 2oy+2 = s2d rows oy, oy+1 with both parities -- every weight is used, K = 4 taps x 4C.
//   in fp32 NHWC [n, H, W, C] -> out split [n, H/2+1, W/2+1, 4C].  One thread per (output pixel, parity, 8-channel chunk).
__global__ void __launch_bounds__(256) s2d_pad_split_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t plane,
                                                            int n_img, int H, int W, int C8, int relu, int* flag) {
  const int Ho = H / 2 + 1, Wo = W / 2 + 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n_img * Ho * Wo * 4 * C8;
  if (t >= total) return;
  const int c = (int)(t % C8);
  int64_t r = t / C8;
  const int q = (int)(r & 3); r >>= 2;
  const int X = (int)(r % Wo); r /= Wo;
  const int Y = (int)(r % Ho);
  const int n = (int)(r / Ho);
  const int y = 2 * Y + (q >> 1) - 1, x = 2 * X + (q & 1) - 1;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (y >= 0 && y < H && x >= 0 && x < W) {
    const float4* src = reinterpret_cast<const float4*>(in + (((int64_t)n * H + y) * W + x) * (int64_t)C8 * 8) + 2 * c;
    a = __ldg(src); b = __ldg(src + 1);
    if (relu) {
      a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
      b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
    }
  }
  uint2 h0, l0, h1, l1;
  const bool bad = split4(a, h0, l0) | split4(b, h1, l1);
  const int64_t o = ((((int64_t)n * Ho + Y) * Wo + X) * 4 + q) * (int64_t)C8 * 8 + c * 8;
  *reinterpret_cast<uint4*>(out + o) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(out + plane + o) = make_uint4(l0.x, l0.y, l1.x, l1.y);
  if (bad && flag) atomicOr(flag, 1);
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// rank-R fp16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first; strides in bytes for dims 1..R-1.
int make_map(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
             CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MAGE_ENOTSUP;
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, dtype, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MAGE_EINVAL;
}

// operand maps (A, W) and the epilogue's store maps: O fp32 result, S split(result), R split(relu(result))
struct Maps { CUtensorMap A, W, O, S, R; };

// Store maps over a (possibly strided / scattered) [n_img, Y, X, Ncols] view: one warp stores a 32-row x 32-column box per
// call, the 32 rows being bw consecutive x positions of bh consecutive y rows (GEMM: X = rows, bw = 32, Y = n_img = 1).
// fp32: 128-byte box rows, SWIZZLE_128B.  split: two planes of 64-byte rows, SWIZZLE_64B.  Strides in ELEMENTS.
int make_store_maps(Maps* mp, float* out, void* split, void* split_relu, int64_t plane, int Ncols, int X, int Y, int n_img,
                    int64_t sx, int64_t sy, int64_t simg, int bw, int bh) {
  const cuuint64_t dims[5] = {(cuuint64_t)Ncols, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)n_img, 1};
  if (out) {
    const cuuint64_t st[4] = {(cuuint64_t)sx * 4, (cuuint64_t)sy * 4, (cuuint64_t)simg * 4, (cuuint64_t)simg * 4 * (cuuint64_t)n_img};
    const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    int r = make_map(&mp->O, out, 5, dims, st, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  const cuuint64_t dims2[5] = {(cuuint64_t)Ncols, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)n_img, 2};
  const cuuint64_t st2[4] = {(cuuint64_t)sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)simg * 2, (cuuint64_t)plane * 2};
  const cuuint32_t box2[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1, 2};
  if (split) {
    int r = make_map(&mp->S, split, 5, dims2, st2, box2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (r) return r;
  }
  if (split_relu) {
    int r = make_map(&mp->R, split_relu, 5, dims2, st2, box2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (r) return r;
  }
  return 0;
}

// Per-device launch configuration of a kernel: opt-in shared-memory size set once per handle (= per device), and how many CTAs /
// CTA pairs can be co-resident (1 CTA per SM; GPCs with an odd SM count leave one SM unpaired).
template <typename K>
int configure_kernel(mage_ctx* ctx, K kernel, int cg, int smem_bytes, int* max_units) {
  const void* fn = reinterpret_cast<const void*>(kernel);
  int k = ctx->find(fn);
  if (k < 0) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    int units = ctx->sms / cg;
    if (cg == 2) {
      cudaLaunchConfig_t q{};
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.gridDim = dim3(ctx->sms & ~1); q.blockDim = dim3(NTHREADS); q.dynamicSmemBytes = smem_bytes; q.attrs = qa; q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &q) == cudaSuccess && n > 0) units = n < units ? n : units;
      else (void)cudaGetLastError();
    }
    k = ctx->add(fn, units, (size_t)smem_bytes);
    if (k < 0) return MAGE_EINVAL;
  }
  *max_units = ctx->cfg_units[k];
  // mage_sm_share: this launch sequence was given a share of the machine (another sequence runs beside it on the rest)
  const int share = ctx->eff_sms() / cg;
  if (share < *max_units) *max_units = share > 0 ? share : 1;
  return 0;
}

template <int BN, int CG, bool NS = false>
int launch_tc(mage_ctx* ctx, const Maps& mp, const TcParams& p, cudaStream_t st) {
  using C = Cfg<BN, CG>;
  int max_units = 0;
  if (int r = configure_kernel(ctx, tc_gemm_kernel<BN, CG, NS>, CG, C::SMEM_BYTES, &max_units)) return r;
  const int tiles = (p.m_tiles / CG) * p.n_tiles;
  const int units = tiles < max_units ? tiles : max_units;
  cudaError_t e = mage_launch_pdl(ctx, tc_gemm_kernel<BN, CG, NS>, dim3(units * CG), dim3(NTHREADS), C::SMEM_BYTES, st, CG, mp.A, mp.W, mp.O, mp.S,
                                  mp.R, p);
  if (e != cudaSuccess) return (int)e;
  return mage_post_launch(ctx);
}

struct TileCfg { int bn, cg, ns; };

// Tile selection.  The kernel is bound by L2->SM operand traffic, so the widest tile that still fills the machine wins:
// CTA pair 256x256 (one TMEM accumulator stage), then pair 256x128 (two stages), then the single-CTA tiles.
// The handle's forced_bn / forced_pair (mage_tc_tuning; seeded from MAGE_TC_BN / MAGE_TC_PAIR) force a choice (tuning + tests).
// Small problems (a decode step of a few prompts: M = 2048 rows at 8 prompts) run one or two waves of tiles, so neither the
// double-buffered accumulator nor wave-averaging helps: what counts is the number of waves and what ONE tile costs -- per
// k-block the larger of its MMA time and the time to pull its operand bytes into the SM (measured: ~110 GB/s per SM -- the rate of
// the 3-4 stage operand ring, bytes in flight / slot turnaround, not a port limit: profiles/r02ao_*, r02ap_*; the 128x64 tiles of
// the K = 2048 GEMM are ring-bound at B = 8, profiles/r02a_*).  Pick the tile that minimises
// waves x (fixed + k_iters x t_k).  Candidates: CTA pairs 256 x {64,128,192,256} (even row-tile count), single 128 x {64,128}.
TileCfg pick_small(int N, int64_t m_tiles, int K, int sms, bool pair_ok) {
  const double ingest = 110e9, clk = 1.9e9, flop_clk = 8192.0;
  TileCfg best{0, 0, 0};
  double best_t = 1e30;
  const int k_iters = K / BK;
  for (int cg = 2; cg >= 1; --cg) {
    if (cg == 2 && !pair_ok) continue;
    const int units = sms / cg;
    for (int bn : {256, 192, 128, 64}) {
      if (N % bn != 0 || (cg == 1 && bn > 128)) continue;
      const int64_t tiles = (m_tiles / cg) * (N / bn);
      const int64_t waves = (tiles + units - 1) / units;
      const double bytes = 2.0 * BM * BK * 2 + 2.0 * (bn / cg) * BK * 2;        // per CTA per k-block: A hi+lo, its W share hi+lo
      const double t_mma = 3.0 * (BK / UK) * (2.0 * BM * bn * UK) / flop_clk / clk;
      const double t_k = bytes / ingest > t_mma ? bytes / ingest : t_mma;
      const double t_epi = 0.35e-6 * (bn / 32);                                // epilogue of the last tile: not overlapped
      const double t = waves * (k_iters * t_k + 0.6e-6) + t_epi;
      if (t < best_t) { best_t = t; best = {bn, cg, 0}; }
    }
  }
  return best;
}

TileCfg pick_cfg(const mage_ctx* ctx, int N, int64_t m_tiles, int K, bool gemm = false) {
  const int forced_bn = ctx->forced_bn, forced_pair = ctx->forced_pair, g_ns = ctx->ns, g_small = ctx->small;
  const int sms = ctx->eff_sms();
  const bool pair_ok = forced_pair != 0 && m_tiles % 2 == 0;
  // N-split 256-wide pair tiles (plain GEMMs only): A is fetched once per 256 output columns, the two 128-column halves keep
  // separate accumulators so the epilogue still overlaps the next tile's MMAs.  g_ns: 0 off, 1 automatic, 2 whenever legal.
  const bool ns_ok = gemm && pair_ok && g_ns != 0 && N % 256 == 0 && (forced_bn == 0 || forced_bn == 256);
  // measured: wins only for long k loops (16384x2048x4096: 530 vs 508 TFLOP/s); at K = 512 the 256x128 tile is 13-30 % faster
  if (ns_ok && (g_ns == 2 || (forced_bn == 0 && K >= 4096 && (m_tiles / 2) * (N / 256) >= sms / 2))) return {256, 2, 1};
  if (forced_bn && N % forced_bn == 0 && (forced_bn != 192 || (pair_ok && forced_pair == 1))) return {forced_bn, (pair_ok && forced_pair == 1) ? 2 : 1, 0};
  // fewer than two waves of the default 256x128 / 128x128 tiling: choose by the one-tile cost model
  if (gemm && forced_pair == -1 && g_small && N % 64 == 0 && m_tiles * ((N + 127) / 128) < 2 * sms) {
    const TileCfg c = pick_small(N, m_tiles, K, sms, pair_ok);
    if (c.bn) return c;
  }
  if (pair_ok) {
    // the 256x128 pair tile (two TMEM accumulator stages, each CTA streams half of the W tile) is the fastest shape on every
    // large GEMM / conv of the path; 256x256 (single accumulator stage) only pays off for very long k loops.
    const int64_t pairs = m_tiles / 2;
    if (N % 256 == 0 && ((pairs * (N / 256) >= sms / 2 && K >= 8192) || (forced_pair == 1 && forced_bn == 0))) return {256, 2, 0};
    if (N % 128 == 0 && (pairs * (N / 128) >= sms / 2 || forced_pair == 1)) return {128, 2, 0};
    if (forced_pair == 1 && N % 64 == 0) return {64, 2, 0};
  }
  // BN = 128 keeps two (main + corr) accumulator stages in the 512 TMEM columns, so the epilogue of one tile
  // overlaps the MMAs of the next.  BN = 64 when N is not a multiple of 128 or the 128-wide tiling would leave most SMs idle.
  if (N % 128 == 0 && (m_tiles * (N / 128) >= sms || N % 64 != 0)) return {128, 1, 0};
  if (N % 64 == 0) return {64, 1, 0};
  return {0, 0, 0};
}

int dispatch(mage_ctx* ctx, TileCfg c, const Maps& mp, const TcParams& p, cudaStream_t st) {
  if (c.cg == 2) {
    if (c.ns) return c.bn == 256 ? launch_tc<256, 2, true>(ctx, mp, p, st) : MAGE_ENOTSUP;
    switch (c.bn) {
      case 256: return launch_tc<256, 2>(ctx, mp, p, st);
      case 192: return launch_tc<192, 2>(ctx, mp, p, st);
      case 128: return launch_tc<128, 2>(ctx, mp, p, st);
      case 64: return launch_tc<64, 2>(ctx, mp, p, st);
    }
    return MAGE_ENOTSUP;
  }
  switch (c.bn) {
    case 256: return launch_tc<256, 1>(ctx, mp, p, st);
    case 128: return launch_tc<128, 1>(ctx, mp, p, st);
    case 64: return launch_tc<64, 1>(ctx, mp, p, st);
  }
  return MAGE_ENOTSUP;
}


template <int BN, int CG>
int launch_halo(mage_ctx* ctx, const Maps& mp, const TcParams& p, cudaStream_t st) {
  using H = HaloCfg<BN, CG>;
  int max_units = 0;
  if (int r = configure_kernel(ctx, tc_conv_halo_kernel<BN, CG>, CG, H::SMEM_MAX, &max_units)) return r;
  const int tiles = (p.m_tiles / CG) * p.n_tiles;
  const int units = tiles < max_units ? tiles : max_units;
  // shared-memory plan (see HaloCfg): resident weights when every weight tile a CTA needs fits next to two A stages and the CTA
  // always works on the same N tile(s); otherwise a weight ring as deep as the budget allows.  Spare room deepens the A ring.
  TcParams q = p;
  const int planes = p.passes == 1 ? 1 : 2;
  q.a_stage_bytes = (planes * p.a_plane_bytes + 1023) & ~1023;
  q.w_slot_bytes = planes * H::W_PLANE_BYTES;
  const int budget = H::SMEM_MAX - H::FIXED_BYTES;
  const bool fixed_nt = p.n_tiles == 1 || p.group == p.n_tiles || (p.group == 1 && units % p.n_tiles == 0);
  const int need = p.taps * p.cin_blocks * p.group;
  if (ctx->resident && fixed_nt && need <= HALO_SW_MAX && 2 * q.a_stage_bytes + need * q.w_slot_bytes <= budget) {
    q.w_resident = 1;
    q.sw = need;
  } else {
    q.w_resident = 0;
    q.sw = (budget - 2 * q.a_stage_bytes) / q.w_slot_bytes;
    if (q.sw > 12) q.sw = 12;
    if (q.sw < 2) return MAGE_ENOTSUP;
  }
  q.sa = 2 + (budget - 2 * q.a_stage_bytes - q.sw * q.w_slot_bytes) / q.a_stage_bytes;
  if (q.sa > HALO_SA_MAX) q.sa = HALO_SA_MAX;
  q.smem_bytes = 1024 + q.sa * q.a_stage_bytes + q.sw * q.w_slot_bytes + H::STAGING_BYTES + HALO_BAR_BYTES;
  cudaError_t e = mage_launch_pdl(ctx, tc_conv_halo_kernel<BN, CG>, dim3(units * CG), dim3(NTHREADS), (size_t)q.smem_bytes, st, CG, mp.A, mp.W,
                                  mp.O, mp.S, mp.R, q);
  if (e != cudaSuccess) return (int)e;
  return mage_post_launch(ctx);
}

// tile choice of the halo kernel: CTA pairs whenever the row-tile count is even (each CTA then streams only half of every
// weight tile), the widest N tile the channel count allows; BN = 256 exists only as a pair (weight ring depth).
TileCfg pick_halo_cfg(const mage_ctx* ctx, int Cout, int64_t m_tiles) {
  const int g_forced_bn = ctx->forced_bn, g_forced_pair = ctx->forced_pair;
  const bool pair_ok = g_forced_pair != 0 && m_tiles % 2 == 0;
  if (g_forced_bn && g_forced_bn != 192 && Cout % g_forced_bn == 0 && (g_forced_bn != 256 || pair_ok)) return {g_forced_bn, (pair_ok && (g_forced_pair == 1 || g_forced_bn == 256)) ? 2 : 1, 0};
  if (pair_ok) {
    if (Cout % 128 == 0) return {128, 2, 0};   // measured faster than 256-wide pair tiles (two accumulator stages)
    if (Cout % 64 == 0) return {64, 2, 0};
  }
  if (Cout % 128 == 0) return {128, 1, 0};
  if (Cout % 64 == 0) return {64, 1, 0};
  return {0, 0, 0};
}

int dispatch_halo(mage_ctx* ctx, TileCfg c, const Maps& mp, const TcParams& p, cudaStream_t st) {
  if (c.cg == 2) {
    switch (c.bn) {
      case 256: return launch_halo<256, 2>(ctx, mp, p, st);
      case 128: return launch_halo<128, 2>(ctx, mp, p, st);
      case 64: return launch_halo<64, 2>(ctx, mp, p, st);
    }
    return MAGE_ENOTSUP;
  }
  switch (c.bn) {
    case 128: return launch_halo<128, 1>(ctx, mp, p, st);
    case 64: return launch_halo<64, 1>(ctx, mp, p, st);
  }
  return MAGE_ENOTSUP;
}

int make_w_map(CUtensorMap* map, const void* W, int64_t ldw, int64_t w_plane, int N, int K, int box_rows, int planes = 2) {
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, 2};
  const cuuint64_t strides[2] = {(cuuint64_t)ldw * 2, (cuuint64_t)w_plane * 2};
  const cuuint32_t box[3] = {BK, (cuuint32_t)box_rows, (cuuint32_t)planes};
  return make_map(map, W, 3, dims, strides, box);
}

}  // namespace

extern "C" int mage_tc_tuning(mage_ctx* ctx, int bn, int pair) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG((bn == 0 || bn == 64 || bn == 128 || bn == 192 || bn == 256) && pair >= -1 && pair <= 1);
  ctx->forced_bn = bn;
  ctx->forced_pair = pair;
  return 0;
}

extern "C" int mage_tc_nsplit(mage_ctx* ctx, int mode) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(mode >= 0 && mode <= 2);
  ctx->ns = mode;
  return 0;
}

extern "C" int mage_tc_conv_halo(mage_ctx* ctx, int enable) {
  MAGE_CHECK_CTX(ctx);
  ctx->halo = enable != 0;
  return 0;
}

extern "C" int mage_split_f32(mage_ctx* ctx, const float* x, int64_t ldx, void* out, int64_t plane, int rows, int C, int relu, int* flag,
                              void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(out) & 7) == 0 &&
                 plane % 4 == 0);
  const int64_t total = (int64_t)rows * (C / 4);
  split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, ldx, reinterpret_cast<__half*>(out), plane, rows,
                                                                                C / 4, relu, flag);
  return mage_post_launch(ctx);
}

extern "C" int mage_patch_rows_split_f32(mage_ctx* ctx, const float* in, void* out, int64_t plane, int n_img, int C, int H, int W, int KW, int pad,
                                         void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && C > 0 && H > 0 && W > 0 && KW > 0 && C * KW <= 64 && pad >= 0 && aligned16(out) && plane % 8 == 0);
  const int64_t total = (int64_t)n_img * H * W * 8;
  patch_rows_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(in, reinterpret_cast<__half*>(out), plane, n_img,
                                                                                          C, H, W, KW, pad);
  return mage_post_launch(ctx);
}

extern "C" int mage_s2d_pad_split_f32(mage_ctx* ctx, const float* in, void* out, int64_t plane, int n_img, int H, int W, int C, int relu, int* flag,
                                      void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0 && aligned16(in) && aligned16(out) &&
                 plane % 8 == 0);
  const int64_t total = (int64_t)n_img * (H / 2 + 1) * (W / 2 + 1) * 4 * (C / 8);
  s2d_pad_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(in, reinterpret_cast<__half*>(out), plane, n_img,
                                                                                       H, W, C / 8, relu, flag);
  return mage_post_launch(ctx);
}

extern "C" int mage_embedding_split(mage_ctx* ctx, const int64_t* idx, const void* table, int64_t table_plane, void* out, int64_t out_plane,
                                    int rows, int C, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && aligned16(table) && aligned16(out) && table_plane % 8 == 0 && out_plane % 8 == 0);
  const int64_t total = (int64_t)rows * (C / 8) * 2;
  mage_launch_pdl(ctx, embedding_split_kernel, (unsigned)((total + 255) / 256), 256, 0, as_stream(stream), 1, idx,
                  reinterpret_cast<const __half*>(table), table_plane, reinterpret_cast<__half*>(out), out_plane, rows, C / 8);
  return mage_post_launch(ctx);
}

namespace {
struct LnArgs { const float* gamma; const float* beta; float eps; void* split; int64_t plane; int* count; };

int gemm_tc_impl(mage_ctx* ctx, const void* A, int64_t lda, int64_t a_plane, const void* W, int64_t ldw, int64_t w_plane,
                 const float* bias, const float* residual, int64_t ldr, int res_mod, float* C, void* C_split,
                 void* C_split_relu, int64_t ldc, int64_t c_plane, int M, int N, int K, int act, int* flag,
                 void* stream, const LnArgs* ln) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(M > 0 && N > 0 && K > 0);
  if (K % BK != 0 || N % 64 != 0) return MAGE_ENOTSUP;
  MAGE_CHECK_ARG(aligned16(A) && aligned16(W) && lda % 8 == 0 && ldw % 8 == 0 && a_plane % 8 == 0 && w_plane % 8 == 0);
  MAGE_CHECK_ARG(ldc % 4 == 0 && (!C || aligned16(C)) && (!bias || aligned16(bias)) && (!residual || (aligned16(residual) && ldr % 4 == 0)));
  MAGE_CHECK_ARG(c_plane % 4 == 0 && (C || C_split || C_split_relu));
  const int m_tiles = (M + BM - 1) / BM;
  TileCfg tcfg = pick_cfg(ctx, N, m_tiles, K, true);
  if (!tcfg.bn) return MAGE_ENOTSUP;
  if (ln && tcfg.ns) tcfg.ns = 0;   // the N-split epilogue (two half-tile drains) has no fused LayerNorm: plain 256x256 pair tile
  const int bn = tcfg.bn;
  Maps mp{};
  CUtensorMap& mapA = mp.A;
  CUtensorMap& mapW = mp.W;
  {
    const cuuint64_t dims[5] = {(cuuint64_t)K, (cuuint64_t)M, 1, 1, 2};
    const cuuint64_t strides[4] = {(cuuint64_t)lda * 2, (cuuint64_t)lda * 2 * (cuuint64_t)M, (cuuint64_t)lda * 2 * (cuuint64_t)M,
                                   (cuuint64_t)a_plane * 2};
    const cuuint32_t box[5] = {BK, BM, 1, 1, 2};
    int r = make_map(&mapA, A, 5, dims, strides, box);
    if (r) return r;
    r = make_w_map(&mapW, W, ldw, w_plane, N, K, bn / tcfg.cg / (tcfg.ns ? 2 : 1));
    if (r) return r;
  }
  TcParams p{};
  p.bias = bias; p.res = residual; p.out = C;
  p.split = reinterpret_cast<__half*>(C_split); p.split_relu = reinterpret_cast<__half*>(C_split_relu);
  p.flag = flag; p.ldr = ldr; p.ldc = ldc; p.split_plane = c_plane; p.split_relu_plane = c_plane;
  p.M = M; p.N = N; p.act = act; p.res_mod = res_mod;
  p.m_tiles = m_tiles; p.n_tiles = N / bn; p.k_iters = K / BK; p.group = 1;
  {
    MAGE_CHECK_ARG(ldc % 8 == 0 || !(C_split || C_split_relu));   // TMA strides are multiples of 16 bytes
    int r = make_store_maps(&mp, C, C_split, C_split_relu, c_plane, N, M, 1, 1, ldc, ldc * (int64_t)M, ldc * (int64_t)M, 32, 1);
    if (r) return r;
  }
  if (ln) {
    // the fused LayerNorm reads whole rows of the fp32 result: N = row width = 512, dense rows, no activation in between
    MAGE_CHECK_ARG(C != nullptr && N == 512 && ldc == 512 && (act & 0xff) == MAGE_ACT_NONE && ln->gamma && ln->beta && ln->split && ln->count &&
                   aligned16(ln->gamma) && aligned16(ln->beta) && (reinterpret_cast<uintptr_t>(ln->split) & 7) == 0 && ln->plane % 4 == 0);
    p.ln_gamma = ln->gamma; p.ln_beta = ln->beta; p.ln_eps = ln->eps; p.ln_split = reinterpret_cast<__half*>(ln->split);
    p.ln_plane = ln->plane; p.ln_count = ln->count;
  }
  return dispatch(ctx, tcfg, mp, p, as_stream(stream));
}
}  // namespace

extern "C" int mage_gemm_tc(mage_ctx* ctx, const void* A, int64_t lda, int64_t a_plane, const void* W, int64_t ldw, int64_t w_plane,
                            const float* bias, const float* residual, int64_t ldr, int res_mod, float* C, void* C_split,
                            void* C_split_relu, int64_t ldc, int64_t c_plane, int M, int N, int K, int act, int* flag,
                            void* stream) {
  return gemm_tc_impl(ctx, A, lda, a_plane, W, ldw, w_plane, bias, residual, ldr, res_mod, C, C_split, C_split_relu, ldc, c_plane, M, N, K,
                      act, flag, stream, nullptr);
}

extern "C" int mage_gemm_tc_ln(mage_ctx* ctx, const void* A, int64_t lda, int64_t a_plane, const void* W, int64_t ldw, int64_t w_plane,
                               const float* bias, const float* residual, int64_t ldr, float* C, int M, int K, const float* ln_gamma,
                               const float* ln_beta, float ln_eps, void* ln_split, int64_t ln_plane, int* ln_count, int* flag,
                               void* stream) {
  const LnArgs ln{ln_gamma, ln_beta, ln_eps, ln_split, ln_plane, ln_count};
  return gemm_tc_impl(ctx, A, lda, a_plane, W, ldw, w_plane, bias, residual, ldr, 0, C, nullptr, nullptr, 512, (int64_t)M * 512, M, 512, K,
                      MAGE_ACT_NONE, flag, stream, &ln);
}

// Fused QKV projection + axial (H / W) attention of one temporal position (or of a batch of positions): mage_b200.h.
extern "C" int mage_qkv_axial_attn_tc(mage_ctx* ctx, const void* A, int64_t a_plane, const void* Wp, int64_t w_plane,
                                      const float* bias_p, void* out_split, int64_t out_plane, int n_img, int R, int n_head,
                                      int K, int axis, float scale, int* flag, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && R == 16 && n_head > 0 && n_head % 2 == 0 && K % BK == 0 && (axis == 1 || axis == 2));
  MAGE_CHECK_ARG(aligned16(A) && aligned16(Wp) && aligned16(bias_p) && aligned16(out_split) && a_plane % 8 == 0 && w_plane % 8 == 0 &&
                 out_plane % 8 == 0);
  const int C = n_head * 32, N = 3 * C;
  const int cg = (ctx->forced_pair != 0) ? 2 : 1;   // an image is two 128-row tiles: always pairable
  Maps mp{};
  // rows of A / of the output are (img, h, w); a tile is 8 lines x 16 positions ALONG the attended axis, so the map's "x"
  // dimension is that axis: w (stride 1 row) for axis 2, h (stride R rows) for axis 1
  const int64_t sx = axis == 2 ? 1 : R, sy = axis == 2 ? R : 1;
  {
    const cuuint64_t dims[5] = {(cuuint64_t)K, (cuuint64_t)R, (cuuint64_t)R, (cuuint64_t)n_img, 2};
    const cuuint64_t strides[4] = {(cuuint64_t)(sx * K * 2), (cuuint64_t)(sy * K * 2), (cuuint64_t)((int64_t)R * R * K * 2), (cuuint64_t)a_plane * 2};
    const cuuint32_t box[5] = {BK, 16, 8, 1, 2};
    int r = make_map(&mp.A, A, 5, dims, strides, box);
    if (r) return r;
    r = make_w_map(&mp.W, Wp, K, w_plane, N, K, 192 / cg);
    if (r) return r;
    r = make_store_maps(&mp, nullptr, out_split, nullptr, out_plane, C, R, R, n_img, sx * C, sy * C, (int64_t)R * R * C, 16, 2);
    if (r) return r;
  }
  TcParams p{};
  p.bias = bias_p; p.split = reinterpret_cast<__half*>(out_split); p.flag = flag;
  p.split_plane = out_plane; p.M = n_img * R * R; p.N = N;
  p.m_tiles = n_img * 2; p.n_tiles = N / 192; p.k_iters = K / BK; p.group = 1;
  p.conv = 1; p.Hout = R; p.Wout = R; p.Wb = 16; p.Hb = 8; p.KW = 1; p.cin_blocks = K / BK; p.cin = K;
  p.wb_shift = 4; p.tiles_x = 1; p.tiles_img = 2;
  p.attn = 1; p.attn_scale = scale;
  return cg == 2 ? launch_tc<192, 2>(ctx, mp, p, as_stream(stream)) : launch_tc<192, 1>(ctx, mp, p, as_stream(stream));
}

// Stride-1 NHWC convolution on the tensor cores.  in: split [n_img,Hin,Win,Cin] (Cin % 64 == 0), w: split
// [Cout][KH][KW][Cin]; output geometry / residual modes / scatter as mage_conv2d_nhwc_f32.
namespace {
struct HeadArgs { const float* w; const float* b; float* out; int cout; int64_t img_stride; };

int conv2d_tc_impl(mage_ctx* ctx, const void* in, int64_t in_plane, const void* w, int64_t w_plane, const float* bias,
                   const float* residual, float* out, void* out_split, void* out_split_relu, int64_t out_plane,
                   int n_img, int Hin, int Win, int Cin, int Hout, int Wout, int Cout, int KH, int KW, int pad_y,
                   int pad_x, int res_mode, int act, int out_sy, int out_sx, int out_oy, int out_ox, int Hfull,
                   int Wfull, int64_t out_img_stride, int* flag, void* stream, const HeadArgs* head, int passes) {
  MAGE_CHECK_ARG(n_img > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && Hout > 0 && Wout > 0 && (passes == 1 || passes == 3));
  if (Cin % BK != 0 || Cout % 64 != 0) return MAGE_ENOTSUP;
  // halo mode: 16x8-pixel tiles, the input patch is fetched once per channel block and shared by all taps
  const int act_id = act & 0xff;
  const int g_forced_bn = ctx->forced_bn, g_forced_pair = ctx->forced_pair;
  bool halo = ctx->halo && KH * KW > 1 && Hout % 16 == 0 && Wout % 8 == 0 && (15 + KH) * (7 + KW) * 256 <= HALO_A_STAGE &&
              (act_id == MAGE_ACT_NONE || act_id == MAGE_ACT_RELU || act_id == MAGE_ACT_TANH);
  TileCfg hcfg{0, 0, 0};
  if (halo) {
    const int64_t hm = (int64_t)n_img * (Hout / 16) * (Wout / 8);
    hcfg = pick_halo_cfg(ctx, Cout, hm);
    if (head) {
      // the pixel head needs all 256 channels of a row in one CTA: two consecutive 128-wide tiles of the same rows (the
      // partial sums stay in registers across the group; TMEM double buffering is kept), or one 256-wide tile when forced
      if (hm % 2 == 0 && g_forced_pair != 0 && Cout == 256) hcfg = {g_forced_bn == 256 ? 256 : 128, 2, 0};
      else halo = false;
    }
    if (!hcfg.bn) halo = false;
  }
  const int Wb = halo ? 8 : (Wout < BM ? Wout : BM);
  if (BM % Wb != 0) return MAGE_ENOTSUP;
  const int Hb = BM / Wb;
  if (Wout % Wb != 0 || Hout % Hb != 0) return MAGE_ENOTSUP;
  MAGE_CHECK_ARG(aligned16(in) && aligned16(w) && in_plane % 8 == 0 && w_plane % 8 == 0 && out_plane % 4 == 0);
  MAGE_CHECK_ARG(res_mode >= 0 && res_mode <= 3 && (res_mode == 0 || residual != nullptr) && (out || out_split || out_split_relu || head));
  MAGE_CHECK_ARG(Cout % 4 == 0 && out_img_stride % 4 == 0 && (!out || aligned16(out)) && (!bias || aligned16(bias)) &&
                 (!residual || aligned16(residual)));
  const int64_t m_tiles = (int64_t)n_img * (Hout / Hb) * (Wout / Wb);
  MAGE_CHECK_ARG(m_tiles < ((int64_t)1 << 24));
  const int K = KH * KW * Cin;
  TileCfg tcfg = halo ? hcfg : pick_cfg(ctx, Cout, m_tiles, K);
  if (head && !halo) {
    // the pixel head needs every output channel of a row in one CTA: one 256-wide N tile
    if (Cout != 256 || (act & 0xff) != MAGE_ACT_NONE) return MAGE_ENOTSUP;
    MAGE_CHECK_ARG(head->w && head->b && head->out && head->cout >= 1 && head->cout <= 3 && aligned16(head->w));
    tcfg.bn = 256;
    if (tcfg.cg == 2 && m_tiles % 2 != 0) tcfg.cg = 1;
  }
  if (!tcfg.bn) return MAGE_ENOTSUP;
  const int bn = tcfg.bn;
  Maps mp{};
  CUtensorMap& mapA = mp.A;
  CUtensorMap& mapW = mp.W;
  {
    const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)n_img, 2};
    const cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)Win * Cin * 2, (cuuint64_t)Hin * Win * Cin * 2,
                                   (cuuint64_t)in_plane * 2};
    // single-pass (hi*hi only) exists in the halo kernel; elsewhere the request falls back to the fp32-grade product
    const int planes = (halo && passes == 1) ? 1 : 2;
    const cuuint32_t box[5] = {BK, (cuuint32_t)(halo ? Wb + KW - 1 : Wb), (cuuint32_t)(halo ? Hb + KH - 1 : Hb), 1, (cuuint32_t)planes};
    int r = make_map(&mapA, in, 5, dims, strides, box);
    if (r) return r;
    r = make_w_map(&mapW, w, K, w_plane, Cout, K, bn / tcfg.cg, planes);
    if (r) return r;
  }
  TcParams p{};
  p.bias = bias; p.res = residual; p.out = out;
  p.split = reinterpret_cast<__half*>(out_split); p.split_relu = reinterpret_cast<__half*>(out_split_relu);
  p.flag = flag; p.ldr = Cout; p.ldc = Cout; p.split_plane = out_plane; p.split_relu_plane = out_plane;
  p.out_img_stride = out_img_stride;
  p.M = (int)(m_tiles * BM); p.N = Cout; p.act = act; p.res_mod = 0;
  p.m_tiles = (int)m_tiles; p.n_tiles = Cout / bn; p.k_iters = KH * KW * (Cin / BK);
  p.conv = 1; p.Hout = Hout; p.Wout = Wout; p.Wb = Wb; p.Hb = Hb; p.KW = KW; p.cin_blocks = Cin / BK; p.cin = Cin;
  p.wb_shift = 0;
  while ((1 << p.wb_shift) < Wb) ++p.wb_shift;
  p.tiles_x = Wout / Wb; p.tiles_img = (Wout / Wb) * (Hout / Hb);
  p.pad_y = pad_y; p.pad_x = pad_x; p.res_mode = res_mode;
  p.out_sy = out_sy; p.out_sx = out_sx; p.out_oy = out_oy; p.out_ox = out_ox; p.Hfull = Hfull; p.Wfull = Wfull;
  p.group = 1;
  if (head) {
    p.head_w = head->w; p.head_b = head->b; p.head_out = head->out; p.head_cout = head->cout; p.head_img_stride = head->img_stride;
    p.group = p.n_tiles;   // all N tiles of a row tile go to the same CTA, back to back
  }
  {
    // store maps over the scattered view out[img, oy*sy + out_oy, ox*sx + out_ox, :] (sub-pixel phases of an upsample / ConvTranspose)
    MAGE_CHECK_ARG(Cout % 8 == 0 && out_img_stride % 8 == 0 && out_plane % 8 == 0);
    const int64_t o0 = ((int64_t)out_oy * Wfull + out_ox) * Cout;
    const int bw = Wb < 32 ? Wb : 32;
    int r = make_store_maps(&mp, out ? out + o0 : nullptr, out_split ? reinterpret_cast<__half*>(out_split) + o0 : nullptr,
                            out_split_relu ? reinterpret_cast<__half*>(out_split_relu) + o0 : nullptr, out_plane, Cout, Wout, Hout, n_img,
                            (int64_t)out_sx * Cout, (int64_t)out_sy * Wfull * Cout, out_img_stride, bw, 32 / bw);
    if (r) return r;
  }
  if (halo) {
    p.taps = KH * KW; p.halo_w = Wb + KW - 1;
    p.passes = passes;
    p.a_plane_bytes = (Hb + KH - 1) * (Wb + KW - 1) * 128; p.a_tx_bytes = (passes == 1 ? 1 : 2) * p.a_plane_bytes;
    p.w_tx_bytes = (passes == 1 ? 1 : 2) * (bn / tcfg.cg) * BK * 2;
    return dispatch_halo(ctx, tcfg, mp, p, as_stream(stream));
  }
  return dispatch(ctx, tcfg, mp, p, as_stream(stream));
}
}  // namespace

extern "C" int mage_conv2d_tc(mage_ctx* ctx, const void* in, int64_t in_plane, const void* w, int64_t w_plane, const float* bias,
                              const float* residual, float* out, void* out_split, void* out_split_relu, int64_t out_plane,
                              int n_img, int Hin, int Win, int Cin, int Hout, int Wout, int Cout, int KH, int KW, int pad_y,
                              int pad_x, int res_mode, int act, int out_sy, int out_sx, int out_oy, int out_ox, int Hfull,
                              int Wfull, int64_t out_img_stride, int passes, int* flag, void* stream) {
  MAGE_CHECK_CTX(ctx);
  return conv2d_tc_impl(ctx, in, in_plane, w, w_plane, bias, residual, out, out_split, out_split_relu, out_plane, n_img, Hin, Win, Cin,
                        Hout, Wout, Cout, KH, KW, pad_y, pad_x, res_mode, act, out_sy, out_sx, out_oy, out_ox, Hfull, Wfull,
                        out_img_stride, flag, stream, nullptr, passes);
}

extern "C" int mage_conv2d_tc_pixel_head(mage_ctx* ctx, const void* in, int64_t in_plane, const void* w, int64_t w_plane, const float* bias,
                                         const float* residual, int n_img, int Hin, int Win, int Cin, int Hout, int Wout,
                                         int Cout, int KH, int KW, int pad_y, int pad_x, int res_mode, const float* head_w,
                                         const float* head_b, int head_cout, float* head_out, int64_t head_img_stride,
                                         int passes, int* flag, void* stream) {
  MAGE_CHECK_CTX(ctx);
  const HeadArgs h{head_w, head_b, head_out, head_cout, head_img_stride};
  return conv2d_tc_impl(ctx, in, in_plane, w, w_plane, bias, residual, nullptr, nullptr, nullptr, 0, n_img, Hin, Win, Cin, Hout, Wout,
                        Cout, KH, KW, pad_y, pad_x, res_mode, MAGE_ACT_NONE, 1, 1, 0, 0, Hout, Wout,
                        (int64_t)Hout * Wout * Cout, flag, stream, &h, passes);
}
