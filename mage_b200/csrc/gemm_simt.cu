// fp32 FFMA register-tiled GEMM / implicit-GEMM convolution (the fp32-exact back end).
//
//   C[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual
//
// Both operands are K-contiguous ("TN"): nn.Linear weights are [out,in] and conv weights are
// packed [Cout][KH][KW][Cin], so one kernel serves nn.Linear, 1x1 convs and -- with the A tile
// gathered on the fly from an NHWC image -- every 3x3 / 4x4 / sub-pixel convolution on the path.
// It serves every shape the tensor-core back end (gemm_tc.cu) does not take and is its in-GPU reference.
#include "common.cuh"

namespace {

struct GemmArgs {
  const float* A;
  const float* W;
  const float* bias;
  const float* res;
  float* C;
  int M, N, K;
  int64_t lda, ldw, ldr, ldc;
  int res_mod, act, relu_a, vec_ok;
  // convolution geometry (A = NHWC image)
  int Hin, Win, Cin, Hout, Wout, KW, stride, pad_y, pad_x, in_up;
  int res_mode, out_sy, out_sx, out_oy, out_ox, Hfull, Wfull;
  int64_t out_img_stride;
};

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN, bool CONV>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(const GemmArgs p) {
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int LA = BM * BK / 4 / 256, LB = BN * BK / 4 / 256;
  constexpr int TCOLS = BN / TN;           // threads along N
  constexpr int GM = TM / 4, GN = TN / 4;  // 4-wide groups per thread
  constexpr int GROUP_M = BM / GM, GROUP_N = BN / GN;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  static_assert(LA >= 1 && LB >= 1, "tile too small");

  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int tid = threadIdx.x;
  const int tx = tid % TCOLS, ty = tid / TCOLS;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- per-thread load slots ----
  const float* a_ptr[LA];
  int a_oy[LA], a_ox[LA];
  bool a_ok[LA];
#pragma unroll
  for (int l = 0; l < LA; ++l) {
    const int row = (tid + l * 256) >> 2;
    const int gm = m0 + row;
    a_ok[l] = gm < p.M;
    a_oy[l] = a_ox[l] = 0;
    if (CONV) {
      const int hw = p.Hout * p.Wout;
      const int img = gm / hw, r = gm - img * hw;
      a_oy[l] = r / p.Wout;
      a_ox[l] = r - a_oy[l] * p.Wout;
      a_ptr[l] = p.A + (int64_t)img * p.Hin * p.Win * p.Cin;
    } else {
      a_ptr[l] = p.A + (int64_t)gm * p.lda;
    }
  }
  const float* b_ptr[LB];
  bool b_ok[LB];
#pragma unroll
  for (int l = 0; l < LB; ++l) {
    const int row = (tid + l * 256) >> 2;
    b_ok[l] = (n0 + row) < p.N;
    b_ptr[l] = p.W + (int64_t)(n0 + row) * p.ldw;
  }
  const int kv = (tid & 3) * 4;

  float4 ra[LA], rb[LB];
  auto load_global = [&](int kb) {
    const int k0 = kb * BK + kv;
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[l] && k0 < p.K) {
        if (CONV) {
          const int tap = k0 / p.Cin, c = k0 - tap * p.Cin;
          const int ky = tap / p.KW, kx = tap - ky * p.KW;
          const int iy = a_oy[l] * p.stride - p.pad_y + ky;
          const int ix = a_ox[l] * p.stride - p.pad_x + kx;
          if (iy >= 0 && ix >= 0 && iy < (p.Hin << p.in_up) && ix < (p.Win << p.in_up))
            v = __ldg(reinterpret_cast<const float4*>(
                a_ptr[l] + ((int64_t)(iy >> p.in_up) * p.Win + (ix >> p.in_up)) * p.Cin + c));
        } else {
          v = __ldg(reinterpret_cast<const float4*>(a_ptr[l] + k0));
        }
        if (p.relu_a) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
      }
      ra[l] = v;
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      rb[l] = (b_ok[l] && k0 < p.K) ? __ldg(reinterpret_cast<const float4*>(b_ptr[l] + k0))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      const int row = (tid + l * 256) >> 2;
      As[buf][kv + 0][row] = ra[l].x; As[buf][kv + 1][row] = ra[l].y;
      As[buf][kv + 2][row] = ra[l].z; As[buf][kv + 3][row] = ra[l].w;
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      const int row = (tid + l * 256) >> 2;
      Bs[buf][kv + 0][row] = rb[l].x; Bs[buf][kv + 1][row] = rb[l].y;
      Bs[buf][kv + 2][row] = rb[l].z; Bs[buf][kv + 3][row] = rb[l].w;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  load_global(0);
  store_smem(0);
  __syncthreads();
  int cur = 0;
  for (int kb = 0; kb < nk; ++kb) {
    if (kb + 1 < nk) load_global(kb + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < GM; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[cur][k][g * GROUP_M + ty * 4]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < GN; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[cur][k][g * GROUP_N + tx * 4]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) store_smem(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

  // ---- epilogue ----
  const int act = p.act & 0xff;
  const bool post = (p.act & MAGE_ACT_POST_RES) != 0, res_relu = (p.act & MAGE_RES_RELU) != 0;
#pragma unroll
  for (int gi = 0; gi < GM; ++gi) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + gi * GROUP_M + ty * 4 + i;
      if (m >= p.M) continue;
      float* crow;
      const float* rrow = nullptr;
      if (CONV) {
        const int hw = p.Hout * p.Wout;
        const int img = m / hw, r = m - img * hw;
        const int oy = r / p.Wout, ox = r - oy * p.Wout;
        crow = p.C + (int64_t)img * p.out_img_stride +
               ((int64_t)(oy * p.out_sy + p.out_oy) * p.Wfull + (ox * p.out_sx + p.out_ox)) * p.N;
        if (p.res_mode == 1) rrow = p.res + (int64_t)m * p.N;
        else if (p.res_mode == 2)
          rrow = p.res + (((int64_t)img * (p.Hout >> 1) + (oy >> 1)) * (p.Wout >> 1) + (ox >> 1)) * p.N;
        else if (p.res_mode == 3) rrow = p.res + (int64_t)r * p.N;
      } else {
        crow = p.C + (int64_t)m * p.ldc;
        if (p.res) rrow = p.res + (int64_t)(p.res_mod > 0 ? m % p.res_mod : m) * p.ldr;
      }
#pragma unroll
      for (int gj = 0; gj < GN; ++gj) {
        const int n = n0 + gj * GROUP_N + tx * 4;
        if (n >= p.N) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[gi * 4 + i][gj * 4 + j];
        if (p.vec_ok && n + 3 < p.N) {
          if (p.bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
          }
          if (!post) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = mage_act(v[j], act);
          }
          if (rrow) {
            float4 rv = *reinterpret_cast<const float4*>(rrow + n);
            if (res_relu) { rv.x = fmaxf(rv.x, 0.f); rv.y = fmaxf(rv.y, 0.f); rv.z = fmaxf(rv.z, 0.f); rv.w = fmaxf(rv.w, 0.f); }
            v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
          }
          if (post) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = mage_act(v[j], act);
          }
          *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j >= p.N) break;
            float o = v[j] + (p.bias ? p.bias[n + j] : 0.f);
            if (!post) o = mage_act(o, act);
            if (rrow) o += res_relu ? fmaxf(rrow[n + j], 0.f) : rrow[n + j];
            if (post) o = mage_act(o, act);
            crow[n + j] = o;
          }
        }
      }
    }
  }
}

template <bool CONV>
int launch(mage_ctx* ctx, const GemmArgs& a, cudaStream_t st) {
  const long t128 = (long)((a.M + 127) / 128) * ((a.N + 127) / 128);
  if (a.N <= 64 || (a.N % 128 != 0 && a.N % 128 <= 64 && a.N < 256)) {
    if ((long)((a.M + 127) / 128) * ((a.N + 63) / 64) >= 148) {
      dim3 g((a.M + 127) / 128, (a.N + 63) / 64);
      sgemm_kernel<128, 64, 8, 4, CONV><<<g, 256, 0, st>>>(a);
    } else {
      dim3 g((a.M + 63) / 64, (a.N + 63) / 64);
      sgemm_kernel<64, 64, 4, 4, CONV><<<g, 256, 0, st>>>(a);
    }
  } else if (t128 >= 148) {
    dim3 g((a.M + 127) / 128, (a.N + 127) / 128);
    sgemm_kernel<128, 128, 8, 8, CONV><<<g, 256, 0, st>>>(a);
  } else {
    dim3 g((a.M + 63) / 64, (a.N + 63) / 64);
    sgemm_kernel<64, 64, 4, 4, CONV><<<g, 256, 0, st>>>(a);
  }
  return mage_post_launch(ctx);
}

}  // namespace

extern "C" int mage_gemm_f32(mage_ctx* ctx, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                             const float* residual, int64_t ldr, int res_mod, float* C, int64_t ldc,
                             int M, int N, int K, int act, int relu_a, void* stream) {
  MAGE_CHECK_CTX(ctx);
  cudaStream_t st = as_stream(stream);
  MAGE_CHECK_ARG(M > 0 && N > 0 && K > 0 && (K % 4) == 0 && (lda % 4) == 0 && (ldw % 4) == 0);
  MAGE_CHECK_ARG(aligned16(A) && aligned16(W));
  GemmArgs a{};
  a.A = A; a.W = W; a.bias = bias; a.res = residual; a.C = C;
  a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldw = ldw; a.ldr = ldr; a.ldc = ldc;
  a.res_mod = res_mod; a.act = act; a.relu_a = relu_a;
  a.vec_ok = (N % 4 == 0) && (ldc % 4 == 0) && aligned16(C) && (!bias || aligned16(bias)) &&
             (!residual || (aligned16(residual) && ldr % 4 == 0));
  return launch<false>(ctx, a, st);
}

extern "C" int mage_conv2d_nhwc_f32(mage_ctx* ctx, const float* in, const float* w, const float* bias, const float* residual,
                                    float* out, int n_img, int Hin, int Win, int Cin, int Hout, int Wout, int Cout,
                                    int KH, int KW, int stride, int pad_y, int pad_x, int in_up, int res_mode,
                                    int relu_in, int act, int out_sy, int out_sx, int out_oy, int out_ox,
                                    int Hfull, int Wfull, int64_t out_img_stride, void* stream) {
  MAGE_CHECK_CTX(ctx);
  MAGE_CHECK_ARG(n_img > 0 && Cin > 0 && (Cin % 4) == 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0);
  MAGE_CHECK_ARG(aligned16(in) && aligned16(w) && (in_up == 0 || in_up == 1));
  MAGE_CHECK_ARG(res_mode >= 0 && res_mode <= 3 && (res_mode == 0 || residual != nullptr));
  MAGE_CHECK_ARG((int64_t)n_img * Hout * Wout < (int64_t)1 << 31);
  GemmArgs a{};
  a.A = in; a.W = w; a.bias = bias; a.res = residual; a.C = out;
  a.M = n_img * Hout * Wout; a.N = Cout; a.K = KH * KW * Cin;
  a.lda = 0; a.ldw = a.K; a.ldr = Cout; a.ldc = Cout;
  a.res_mod = 0; a.act = act; a.relu_a = relu_in;
  a.Hin = Hin; a.Win = Win; a.Cin = Cin; a.Hout = Hout; a.Wout = Wout; a.KW = KW;
  a.stride = stride; a.pad_y = pad_y; a.pad_x = pad_x; a.in_up = in_up;
  a.res_mode = res_mode; a.out_sy = out_sy; a.out_sx = out_sx; a.out_oy = out_oy; a.out_ox = out_ox;
  a.Hfull = Hfull; a.Wfull = Wfull; a.out_img_stride = out_img_stride;
  a.vec_ok = (Cout % 4 == 0) && aligned16(out) && (out_img_stride % 4 == 0) && (!bias || aligned16(bias)) &&
             (!residual || aligned16(residual));
  return launch<true>(ctx, a, as_stream(stream));
}
