/*
 * mage_b200 -- C ABI of libmage_sm100.so: hand-written sm_100a kernels for the MAGE
 * autoregressive video-token sampling path.
 *
 * The reference (Youncy-Hu/MAGE) has no FFI: its hot path is Python calling torch.nn
 * modules.  Each entry point below replaces the torch call(s) cited beside
 * it (file:line in /root/reference); INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - the first argument of every call is the opaque handle made by mage_ctx_create (one per GPU / process); the library has
 *     no other mutable state;
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocator); nothing is
 *     allocated, freed or cached by the library, so it is re-entrant per handle;
 *   - activations are channels-last fp32 (`[rows, C]`, images `[N,H,W,C]`); token / code
 *     indices are int64 (what the reference's `torch.max`/`torch.min` return);
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and return at once;
 *   - return value: 0 on success, a positive cudaError_t from the launch, or a negative
 *     MAGE_E* code for an unsupported argument.  Nothing throws across the ABI;
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef MAGE_B200_H
#define MAGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAGE_EINVAL (-1)   /* shape / alignment not supported */
#define MAGE_ENOTSUP (-2)  /* variant not built */

/* epilogue activations */
#define MAGE_ACT_NONE 0
#define MAGE_ACT_RELU 1
#define MAGE_ACT_QUICKGELU 2 /* x*sigmoid(1.702x): mage_model.py:11-13 */
#define MAGE_ACT_GELU 3      /* exact erf GELU: nn.TransformerEncoderLayer(activation="gelu"), mage_model.py:192-199 */
#define MAGE_ACT_TANH 4      /* vqvae_model.py:188,213 */

/* flags OR-ed into `act` of mage_gemm_f32 / mage_conv2d_nhwc_f32 */
#define MAGE_ACT_POST_RES 0x100 /* apply the activation after the residual add: act(x + bias + residual) */
#define MAGE_RES_RELU 0x200     /* read the residual through a ReLU (in-place ReLU skip of ResBlock, vqvae_model.py:114-124) */

/* Library info */
int mage_abi_version(void); /* 5 */

/* The opaque handle.  Everything the library remembers between calls lives in it: the device it was
 * created for, that device's SM count and per-kernel launch configuration (opt-in shared-memory size, number of co-resident
 * CTA pairs), the tile-selection / launch switches below, the launch counter.  There is no other mutable state, so the library
 * is re-entrant per handle: one handle per (process, GPU) -- or per thread -- and two handles never see each other.  Every entry
 * point takes the handle first; the caller keeps that device current (cudaSetDevice) and passes its own stream.
 * mage_ctx_create fails with MAGE_ENOTSUP on anything but an sm_100 device (no other code path exists). */
typedef struct mage_ctx mage_ctx;
int mage_ctx_create(int device, mage_ctx** out);
int mage_ctx_destroy(mage_ctx* ctx);
int mage_ctx_device(mage_ctx* ctx);
/* Number of kernels launched through this handle so far. */
int64_t mage_launch_count(mage_ctx* ctx);

/* enable != 0: the kernels of the per-step path are launched with the programmatic-stream-serialization attribute (they call
 * griddepcontrol.launch_dependents / .wait themselves).  Default on (MAGE_PDL=0 turns it off): the next kernel's launch latency and
 * prologue overlap the previous kernel's tail -- 3 % per generate at 8 prompts per GPU, neutral at 64; results are bit-identical. */
int mage_pdl(mage_ctx* ctx, int enable);
/* Give the launches that follow a SHARE of the machine: the persistent tensor-core kernels (GEMM, convolution, fused attention)
 * size their grids for at most `sms` SMs (0 = all; they hold one CTA per SM).  Two launch sequences on two streams whose shares
 * add up to the SM count then run side by side without ever waiting for each other's CTAs to retire -- the VQ-VAE decoder next
 * to the latency-bound decode steps of a small batch.  Tiles are independent, so the share cannot change a bit of any result. */
int mage_sm_share(mage_ctx* ctx, int sms);
/* mage_temporal_attn_step_f32 as one short-lived CTA per (location, head half) (default) or as a persistent kernel with a ring of
 * staging slots (MAGE_TATTN_RING=1; measured 1.2-1.4 % slower per generate): same arithmetic per unit, same bits. */
int mage_temporal_attn_ring(mage_ctx* ctx, int enable);

/* C[M,N] = act(relu_a?(A)[M,K] . W[N,K]^T + bias[N]) + residual
 * residual row for output row m is (res_mod > 0 ? m % res_mod : m), leading dim ldr; may alias C.
 * Replaces nn.Linear / MHA in-proj / out-proj / MLP (mage_model.py:20-26,33,50-51,375-376,385),
 * and 1x1 convolutions on NHWC data (vqvae_model.py:131,142,150,153).  K % 4 == 0, lda/ldw % 4 == 0. */
int mage_gemm_f32(mage_ctx* ctx, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                  const float* residual, int64_t ldr, int res_mod, float* C, int64_t ldc,
                  int M, int N, int K, int act, int relu_a, void* stream);

/* Implicit-GEMM 2-D convolution, NHWC fp32, weights packed [Cout][KH][KW][Cin] (Cin % 4 == 0).
 *   out[n, oy*out_sy+out_oy, ox*out_sx+out_ox, :] =
 *        act( sum_{ky,kx,c} relu_in?(in_up(in))[n, oy*stride-pad_y+ky, ox*stride-pad_x+kx, c] * w[:,ky,kx,c] + bias )
 *        + residual
 *   in_up = 1 reads the stored [Hin,Win] input through a nearest x2 upsample (nn.Upsample,
 *   vqvae_model.py:205-209) without materialising it; res_mode: 0 none, 1 same shape as out,
 *   2 stored at half resolution and read through nearest x2, 3 one [Hout,Wout,Cout] map shared
 *   by all images (H/W positional embeddings, mage_model.py:649,676).
 *   The out_s / out_o scatter writes one sub-pixel phase of a ConvTranspose2d(4,2,1)
 *   (vqvae_model.py:184,187) into the full [Hfull,Wfull] output; out_img_stride is in elements.
 * Replaces nn.Conv2d / nn.ConvTranspose2d calls of vqvae_model.py:111-166,172-214 and
 * mage_model.py:485-488,304-305,504. */
int mage_conv2d_nhwc_f32(mage_ctx* ctx, const float* in, const float* w, const float* bias, const float* residual, float* out,
                         int n_img, int Hin, int Win, int Cin, int Hout, int Wout, int Cout,
                         int KH, int KW, int stride, int pad_y, int pad_x,
                         int in_up, int res_mode, int relu_in, int act,
                         int out_sy, int out_sx, int out_oy, int out_ox, int Hfull, int Wfull,
                         int64_t out_img_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core back end (tcgen05.mma + TMEM accumulators + TMA operand tiles), fp32-grade.
 *
 * "split" tensors: an fp32 tensor [rows, C] carried as two fp16 planes (hi at `ptr`, lo at
 * `ptr + plane` elements): x ~= hi + lo * 2^-11 to ~2^-24 relative for |x| < 65504.  Every product is
 * three fp16 MMAs with fp32 accumulation (hi*hi, lo*hi, hi*lo), i.e. fp32-grade results at a third
 * of the fp16 tensor rate.  `flag` (device int, may be NULL) is OR-ed with 1 when a value outside
 * the fp16 range is split -- callers check it instead of trusting a saturated result.
 * --------------------------------------------------------------------------------------------- */

/* Tuning / test hook for the tile selection of mage_gemm_tc / mage_conv2d_tc (per handle):
 * bn in {0 = automatic, 64, 128, 256} forces the N tile; pair in {-1 automatic, 0 single-CTA tiles only,
 * 1 CTA-pair (tcgen05 cta_group::2, 256-row tiles) whenever the row-tile count is even}.  Results do not depend on it
 * beyond fp32 summation order (identical here: the k order is the same for every tile shape). */
int mage_tc_tuning(mage_ctx* ctx, int bn, int pair);
/* enable != 0 (default): KHxKW convolutions whose output is a multiple of 16x8 pixels run in halo mode (the input patch of a
 * tile is fetched once per 64-channel block and shared by all taps through shifted shared-memory descriptors); 0: every tap
 * re-fetches its own box.  Same results either way (same k order). */
int mage_tc_conv_halo(mage_ctx* ctx, int enable);
/* N-split 256-wide CTA-pair tiles of mage_gemm_tc (two 128-column halves with separate TMEM accumulators and barriers):
 * 0 never, 1 automatic (default), 2 whenever the shape allows (N % 256 == 0, even row-tile count). */
int mage_tc_nsplit(mage_ctx* ctx, int mode);

/* out(split)[r, :] = split(relu?(x[r, :])); x row stride ldx (elements), C % 4 == 0. */
int mage_split_f32(mage_ctx* ctx, const float* x, int64_t ldx, void* out, int64_t plane, int rows, int C, int relu, int* flag,
                   void* stream);

/* First-layer im2row for the tensor cores: in planar [n,C,H,W] fp32 -> out split [n,H,W,64] with
 *   out[n,y,x, kx*C + c] = in[n,c,y, x+kx-pad]   (zero outside the image; channels >= KW*C are zero),  C*KW <= 64.
 * A KHxKW convolution of the image (vqvae_model.py:193, 7x7 pad 3) is then mage_conv2d_tc with a KHx1 kernel over these 64
 * channels and weights w2[co, ky, 0, kx*C + c] = w[co, c, ky, kx]. */
int mage_patch_rows_split_f32(mage_ctx* ctx, const float* in, void* out, int64_t plane, int n_img, int C, int H, int W, int KW, int pad,
                              void* stream);

/* Space-to-depth with a one-pixel top/left pad, in the split format: in fp32 NHWC [n,H,W,C] (H, W even, C % 8 == 0) -> out split
 * [n, H/2+1, W/2+1, 4C] with out[n, Y, X, (py*2+px)*C + c] = relu?(in[n, 2Y+py-1, 2X+px-1, c]) (zero outside the image).
 * A 4x4 stride-2 pad-1 convolution (vqvae_model.py:175) over `in` is then the 2x2 stride-1 VALID mage_conv2d_tc over `out`
 * with w2[co, ty, tx, (py*2+px)*C + c] = w[co, c, 2ty+py, 2tx+px]: same products, every weight used once. */
int mage_s2d_pad_split_f32(mage_ctx* ctx, const float* in, void* out, int64_t plane, int n_img, int H, int W, int C, int relu, int* flag,
                           void* stream);

/* out(split)[r, :] = table(split)[idx[r], :]   (nn.Embedding on a pre-split table: mage_model.py:644,682;
 * vqvae_model.py:240).  C % 8 == 0. */
int mage_embedding_split(mage_ctx* ctx, const int64_t* idx, const void* table, int64_t table_plane, void* out, int64_t out_plane,
                         int rows, int C, void* stream);

/* C[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) (+ residual), A and W split tensors (row strides lda/ldw in
 * elements), K % 64 == 0, N % 64 == 0 (else MAGE_ENOTSUP: use mage_gemm_f32).  Any of the three outputs
 * may be NULL: C fp32 [M,N] (ldc), C_split = split(result), C_split_relu = split(relu(result)) -- the
 * operand format of the next tensor-core op -- all with row stride ldc and plane stride c_plane.
 * act / residual semantics as mage_gemm_f32.  Same reference call sites as mage_gemm_f32. */
int mage_gemm_tc(mage_ctx* ctx, const void* A, int64_t lda, int64_t a_plane, const void* W, int64_t ldw, int64_t w_plane,
                 const float* bias, const float* residual, int64_t ldr, int res_mod, float* C, void* C_split,
                 void* C_split_relu, int64_t ldc, int64_t c_plane, int M, int N, int K, int act, int* flag,
                 void* stream);

/* mage_gemm_tc for the two GEMMs that write the residual stream (attention out-projection and c_proj, N = 512, fp32 result with
 * residual, no activation: mage_model.py:50-51) with the FOLLOWING LayerNorm (ln_2 of the block / ln_1 of the next, :49-51) fused:
 *   C[M,512] = A . W^T + bias + residual        and        ln_split = split(LayerNorm_{gamma,beta,eps}(C))   (the next GEMM's operand)
 * A LayerNorm needs whole rows while the GEMM is tiled over N, so every CTA, once the stores of a tile are complete and visible
 * device-wide, bumps ln_count[128-row block]; the CTA that brings it to N / tile_width normalises those 128 rows (read back with
 * ld.global.cg) -- same arithmetic as mage_layernorm_f32, row by row, so the result does not depend on which CTA does it.
 * ln_count: int32 [ceil(M / 128)], zero before the first call; the kernel leaves it zero.  residual may alias C. */
int mage_gemm_tc_ln(mage_ctx* ctx, const void* A, int64_t lda, int64_t a_plane, const void* W, int64_t ldw, int64_t w_plane,
                    const float* bias, const float* residual, int64_t ldr, float* C, int M, int K, const float* ln_gamma,
                    const float* ln_beta, float ln_eps, void* ln_split, int64_t ln_plane, int* ln_count, int* flag, void* stream);

/* Fused QKV projection + axial attention of the H / W blocks (AxialAttentionBlock.attention, mage_model.py:31-33, with the
 * permutes of :36-47 expressed in tensor maps): for rows ordered (img, h, w) of A = split(ln_1(x)) [n_img*R*R, K],
 *   out = softmax(q k^T * scale) v   per (img, line, head) over the R = 16 positions of the attended axis
 *         (axis 1: along h for fixed w;  axis 2: along w for fixed h),   [q|k|v] = A . W_in^T + b_in,
 * written as split(out) [n_img*R*R, n_head*32] -- the operand of the out-projection.  One tcgen05 GEMM whose epilogue does the
 * attention: the [rows, 3C] QKV tensor never reaches memory.  Wp / bias_p are W_in / b_in with rows PERMUTED so that every
 * 192-row tile holds [q|k|v] x 32 of two heads: new row t*192 + s*96 + part*32 + d  <-  old row part*C + (2t+s)*32 + d.
 * R = 16, head_dim 32, n_head even, K % 64 == 0. */
int mage_qkv_axial_attn_tc(mage_ctx* ctx, const void* A, int64_t a_plane, const void* Wp, int64_t w_plane, const float* bias_p,
                           void* out_split, int64_t out_plane, int n_img, int R, int n_head, int K, int axis, float scale,
                           int* flag, void* stream);

/* Stride-1 NHWC convolution as a tcgen05 implicit GEMM: the A tile of tap (ky,kx) is a TMA box of the
 * split input [n_img,Hin,Win,Cin] shifted by (ky-pad_y, kx-pad_x); out-of-bounds zero fill is the padding.
 * w split [Cout][KH][KW][Cin]; Cin % 64 == 0, Cout % 64 == 0, 128 % min(Wout,128) == 0 (else MAGE_ENOTSUP).
 * Output scatter / residual modes as mage_conv2d_nhwc_f32; outputs as mage_gemm_tc.  Same reference call
 * sites as mage_conv2d_nhwc_f32, plus the f4 stack's stride-2 / transposed convolutions (vqvae_model.py:175,184) once the
 * caller has rewritten them as stride-1 convolutions (mage_s2d_pad_split_f32; 2x2 sub-pixel phases).
 * passes: 3 = fp32-grade products (hi*hi + lo*hi + hi*lo); 1 = hi*hi only (plain fp16 operands, fp32 accumulation; only the
 * hi planes are fetched) -- allowed ONLY where the result feeds no token: the last block of the VQ-VAE decoder, whose pixel
 * error stays inside the 1e-3 bar (tests/test_gpu_parity.py::test_decoder_precision_budget).  A shape the single-pass kernel
 * does not cover silently runs with 3 passes. */
int mage_conv2d_tc(mage_ctx* ctx, const void* in, int64_t in_plane, const void* w, int64_t w_plane, const float* bias,
                   const float* residual, float* out, void* out_split, void* out_split_relu, int64_t out_plane,
                   int n_img, int Hin, int Win, int Cin, int Hout, int Wout, int Cout, int KH, int KW, int pad_y,
                   int pad_x, int res_mode, int act, int out_sy, int out_sx, int out_oy, int out_ox, int Hfull,
                   int Wfull, int64_t out_img_stride, int passes, int* flag, void* stream);

/* mage_conv2d_tc fused with the decoder's pixel head (vqvae_model.py:210-213: DecoderBlock conv -> ReLU -> Conv2d(dim,C,1) ->
 * Tanh): the conv result x[row, 0..255] (+ bias + residual, never written to memory) is reduced in the epilogue to
 *   head_out[img, c, y, x] = tanh(head_b[c] + sum_n relu(x[row, n]) * head_w[c, n]),   c < head_cout <= 3,
 * planar output, image stride head_img_stride elements.  Cout must be 256 (one N tile holds a whole row). */
int mage_conv2d_tc_pixel_head(mage_ctx* ctx, const void* in, int64_t in_plane, const void* w, int64_t w_plane, const float* bias,
                              const float* residual, int n_img, int Hin, int Win, int Cin, int Hout, int Wout, int Cout,
                              int KH, int KW, int pad_y, int pad_x, int res_mode, const float* head_w, const float* head_b,
                              int head_cout, float* head_out, int64_t head_img_stride, int passes, int* flag, void* stream);

/* First-layer convolution from a planar NCHW image with a tiny channel count (Cin <= 4):
 * in [N,Cin,H,W], w_t [Cin*KH*KW][Cout] (transposed), out NHWC [N,Hout,Wout,Cout], optional ReLU.
 * Replaces vqvae_model.py:172-173 (4x4 s2, BN folded by the caller) and :193 (7x7 pad 3). */
int mage_conv2d_first_f32(mage_ctx* ctx, const float* in, const float* w_t, const float* bias, float* out,
                          int n_img, int Cin, int H, int W, int Hout, int Wout, int Cout,
                          int KH, int KW, int stride, int pad, int act, void* stream);

/* Last f8 decoder layer: out[n,c,y,x] = tanh(sum_k relu(in[n,y,x,k]) * w[c,k] + bias[c]), planar
 * output with image stride out_img_stride elements (vqvae_model.py:211-213). Cin % 128 == 0, Cout <= 4. */
int mage_conv1x1_tanh_nchw_f32(mage_ctx* ctx, const float* in, const float* w, const float* bias, float* out,
                               int n_img, int HW, int Cin, int Cout, int64_t out_img_stride, void* stream);

/* 2x2 max pooling, NHWC (vqvae_model.py:195,197,199). */
int mage_maxpool2x2_nhwc_f32(mage_ctx* ctx, const float* in, float* out, int n_img, int Hin, int Win, int C, void* stream);

/* Row LayerNorm over the last dim C (C % 128 == 0, C <= 1024); in may alias out.  out (fp32) and/or
 * out_split (split copy for a following tensor-core GEMM, plane stride split_plane) may be NULL.
 * Replaces nn.LayerNorm (mage_model.py:21,27,84,204,206). */
int mage_layernorm_f32(mage_ctx* ctx, const float* in, const float* gamma, const float* beta, float* out, void* out_split,
                       int64_t split_plane, int* flag, int rows, int C, float eps, void* stream);

/* Multi-head attention core, head_dim 32: for every (outer, inner, head, query)
 *   out = softmax(scale * q . K^T [keys >= key_len[outer] masked]) . V        (Sk <= 64)
 * Element strides address q/k/v/out as base + outer*X_outer + inner*X_inner + s*X_seq + head*32.
 * Covers the SDPA inside every nn.MultiheadAttention on the path (mage_model.py:33,89,193-199):
 * temporal attention over the K/V cache, H-/W-axial attention, text self-attention (key padding),
 * motion-anchor cross-attention.  out (fp32) and/or out_split (split copy, same element offsets) may be NULL. */
int mage_mha_f32(mage_ctx* ctx, const float* q, const float* k, const float* v, float* out,
                 int n_outer, int n_inner, int n_head, int Sq, int Sk,
                 int64_t q_outer, int64_t q_inner, int64_t q_seq,
                 int64_t k_outer, int64_t k_inner, int64_t k_seq,
                 int64_t v_outer, int64_t v_inner, int64_t v_seq,
                 int64_t o_outer, int64_t o_inner, int64_t o_seq,
                 const int32_t* key_len, float scale, void* out_split, int64_t split_plane, int* flag, void* stream);

/* H-/W-axial attention of one temporal position (the non-causal blocks i % 3 == 1, 2 of FlatAxialDecoder,
 * mage_model.py:35-53,340-345): qkv [B*R*R, 3C] rows ordered (b,h,w), R = 16, head_dim 32, C = 32*n_head.
 * axis 1 attends along h (fixed b,w), axis 2 along w (fixed b,h).  One warp per (line, head), tiles staged
 * once in shared memory; out fp32 [B*R*R, C] and/or out_split may be NULL. */
int mage_axial_attn_f32(mage_ctx* ctx, const float* qkv, float* out, void* out_split, int64_t split_plane, int* flag, int B, int R,
                        int n_head, int axis, float scale, void* stream);

/* Temporal attention for one decode step with a TMA-staged K/V cache (bulk async copies into
 * shared memory).  qkv [M, 3C] holds this position's q|k|v; k,v are appended to the caches at `pos` and q attends
 * positions 0..pos.  out [M, C].  C = 512, 16 heads x 32.  The caches belong to this entry point and are laid out
 * [M][2 head-halves][Lmax][256] (each CTA's live prefix is one contiguous block = one bulk copy); they hold M*Lmax*C floats. */
int mage_temporal_attn_step_f32(mage_ctx* ctx, const float* qkv, float* kcache, float* vcache, float* out, void* out_split,
                                int64_t split_plane, int* flag, int M, int pos, int Lmax, float scale, void* stream);

/* Append this step's K and V (columns C..3C of qkv [M,3C]) at position `pos` of caches [M,Lmax,C]. */
int mage_kv_append_f32(mage_ctx* ctx, const float* qkv, float* kcache, float* vcache, int M, int C, int pos, int Lmax, void* stream);

/* L2 nearest-code search of the VectorQuantizer (vqvae_model.py:8-25):
 *   idx[n] = argmin_k ( (|c_k|^2 + |z_n|^2) - 2 z_n.c_k ),   z [N,D], codebook [K,D], D % 16 == 0, K % 128 == 0.
 * csq_scratch: K floats of scratch.  Ties resolve to the lowest index. */
int mage_vq_argmin_f32(mage_ctx* ctx, const float* z, const float* codebook, float* csq_scratch, int64_t* idx,
                       int N, int D, int K, void* stream);

/* idx[r] = argmax_n x[r, n] (lowest index on ties); greedy decode, mage_model.py:681,687. */
int mage_argmax_rows_f32(mage_ctx* ctx, const float* x, int64_t ldx, int64_t* idx, int rows, int N, void* stream);

/* out[r, :] = table[idx[r], :]  (nn.Embedding: mage_model.py:644,682; vqvae_model.py:240). C % 4 == 0. */
int mage_embedding_f32(mage_ctx* ctx, const int64_t* idx, const float* table, float* out, int rows, int C, void* stream);

/* Convolution of a codebook-embedded token map followed by a linear layer, as table lookups.  Because the conv input is one of K
 * embedding rows per pixel, in_linear(conv3x3(E[tok]) + pos) (mage_model.py:674-676,375) is exactly
 *   out[b,y,x,:] = sum_{ky,kx} table[ky*KW+kx][tok[b, y+ky-KH/2, x+kx-KW/2]][:] + pos_bias[y*R+x][:] + bias[:]
 * with table[tap][code] = W_in . Wc[:, :, tap] . E[code] precomputed at load ([KH*KW, K, C] fp32), pos_bias = W_in . pos.
 * tok int64 [n_img, R, R]; zero padding outside the map; C = 512; out fp32 [n_img*R*R, C]. */
int mage_token_taps_f32(mage_ctx* ctx, const int64_t* tok, const float* table, const float* pos_bias, const float* bias, float* out,
                        int n_img, int R, int K, int C, int KH, int KW, void* stream);

/* mage_token_taps_f32 that also applies the FIRST block's ln_1 (mage_model.py:49, eps 1e-5) to each finished row and writes it as
 * the split operand of that block's QKV projection: ln_split = split(LayerNorm(out)), plane stride ln_plane elements. */
int mage_token_taps_ln_f32(mage_ctx* ctx, const int64_t* tok, const float* table, const float* pos_bias, const float* bias, float* out,
                           int n_img, int R, int K, int C, int KH, int KW, const float* ln_gamma, const float* ln_beta, float ln_eps,
                           void* ln_split, int64_t ln_plane, int* flag, void* stream);

/* Text-encoder front end (mage_model.py:224-237): x[b,t,:] = LN_eps(tok_emb[text[b,t]] + pos_emb[t]) * (text[b,t] != pad);
 * key_len[b] = #non-pad tokens.  C = 512.  tok_emb has `vocab` rows: an id outside [0, vocab) -- on which the reference's
 * nn.Embedding raises (mage_model.py:228) -- is never dereferenced; bit 1 of *flag is set instead (the host raises after the call). */
int mage_text_embed_f32(mage_ctx* ctx, const int64_t* text, const float* tok_emb, const float* pos_emb,
                        const float* gamma, const float* beta, float* x, int32_t* key_len,
                        int B, int T, int C, int pad_idx, float eps, int vocab, int* flag, void* stream);

/* AdaIN (mage_model.py:309-314): out = gamma * InstanceNorm(x) + beta over the HW positions of each
 * (image, channel); x/gamma/beta/out NHWC [N,HW,C]; out may alias x. */
int mage_adain_nhwc_f32(mage_ctx* ctx, const float* x, const float* gamma, const float* beta, float* out,
                        int n_img, int HW, int C, float eps, void* stream);

/* x[n, p, :] += s[n] * vec[:]  (speed embedding, mage_model.py:666-668). */
int mage_add_scaled_vec_f32(mage_ctx* ctx, float* x, const float* s, const float* vec, int n_img, int HW, int C, void* stream);

/* out[n,h,w,c] = in[n,c,h,w]  (noise [B,64,16,16] -> NHWC). */
int mage_nchw_to_nhwc_f32(mage_ctx* ctx, const float* in, float* out, int n_img, int C, int HW, void* stream);

/* Temporal attention of n_pos CONSECUTIVE positions pos0 .. pos0+n_pos-1 in one launch (full-sequence form, used by the MAGE+
 * suffix re-evaluation; causal mask of mage_model.py:367-372 = "query at position p reads keys 0..p"): qkv rows of position s,
 * location m at row s*M + m ([n_pos*M, 3*512] fp32); the new k,v rows are appended to the caches ([M][2][Lmax][256], as
 * mage_temporal_attn_step_f32); out_split = split(attention) with the same row order.  Equivalent to n_pos calls of
 * mage_temporal_attn_step_f32 (same math order), with the K/V prefix of a location staged once. */
int mage_temporal_attn_seq_f32(mage_ctx* ctx, const float* qkv, float* kcache, float* vcache, void* out_split, int64_t split_plane,
                               int* flag, int M, int pos0, int n_pos, int Lmax, float scale, void* stream);

/* MAGE+ continuous head (use_cids=False), mage_model.py:349-354 + :386-388: GroupNorm(groups) over (C/groups channels x all
 * temporal slots x H x W) of a sample -> SiLU -> 1x1x1 Conv3d to `cout` latent channels.
 * mage_gn_partial_f32: x fp32 [n_slots, B, HW, C] -> part double [n_slots, B, groups, 2] = (sum, sum of squares) of each
 *   (slot, sample, group); a slot is re-reduced only when its hidden state changes.
 * mage_gn_silu_head_f32: rows of x [rows = k*B*HW, C] (row -> sample (row / HW) % B), statistics combined over the n_slots
 *   slots of `part`; out fp32 [rows, cout] = bias + w[cout, C] . silu(GN(x)).  C = 512, groups = 32, cout <= 8. */
int mage_gn_partial_f32(mage_ctx* ctx, const float* x, double* part, int n_slots, int B, int HW, int C, int groups, void* stream);
int mage_gn_silu_head_f32(mage_ctx* ctx, const float* x, const double* part, const float* gamma, const float* beta, const float* w,
                          const float* bias, float* out, int rows, int B, int HW, int n_slots, int C, int groups, int cout,
                          float eps, void* stream);

/* ---- Stage-2 objective, forward half (MAGE.forward in eval mode, mage_model.py:575-639; the "next" row N2) ----------------------
 * The teacher-forced pass itself runs on the sampling kernels above (same blocks, the given tokens fed back); these are the pieces
 * only the objective needs.  The 3x3x3 convolutions of the video posterior (BasicBlock, mage_model.py:264-297) are mage_conv2d_tc
 * calls with the three temporal taps folded into the channel axis (Cin = 3*512).
 *
 * mage_gn_apply_f32: GroupNorm(groups) over (C/groups channels x n_slots frames x HW positions) of each sample, from the partial
 *   sums of mage_gn_partial_f32 (part double [n_slots, B, groups, 2]); x fp32 [n_slots, B, HW, C=512] (frame-major);
 *   y = GN(x) * gamma + beta (+ residual) (ReLU if relu) -> out fp32 and/or out_split.  stat: scratch, 2*B*groups floats.
 * mage_cross_entropy_rows_f32: loss[row] = logsumexp(logits[row, 0..K)) - logits[row, target[row]] (F.cross_entropy, :619);
 *   a target outside [0, K) sets bit 2 of *flag (the reference raises).
 * mage_reparam_kl_f32: mu_logvar fp32 [B, HW, 2*Cz] (conv_mu2 | conv_var2 columns), eps fp32 [B, Cz, HW] (the stored draw):
 *   z[b,c,p] = eps * exp(0.5 * logvar) + mu (NCHW, :569-573; z may be NULL), kl_rows[b] = sum(1 + logvar - mu^2 - exp(logvar)) (:625).
 * mage_scaled_sum_f32: out[0] = scale * sum(x[0..n)), one block, fixed order, double accumulation (the means of :619 and :625). */
int mage_gn_apply_f32(mage_ctx* ctx, const float* x, const double* part, float* stat, const float* gamma, const float* beta,
                      const float* residual, float* out, void* out_split, int64_t split_plane, int* flag, int n_slots, int B, int HW, int C,
                      int groups, int relu, float eps, void* stream);
int mage_cross_entropy_rows_f32(mage_ctx* ctx, const float* logits, int64_t ld, const int64_t* target, float* loss, int rows, int K,
                                int* flag, void* stream);
int mage_reparam_kl_f32(mage_ctx* ctx, const float* mu_logvar, const float* eps, float* z, float* kl_rows, int B, int HW, int Cz,
                        void* stream);
int mage_scaled_sum_f32(mage_ctx* ctx, const float* x, float* out, int64_t n, double scale, void* stream);
/* out[0] = scale * sum((a[i] - b[i])^2): F.mse_loss of the MAGE+ objective (mage_model.py:621), same reduction scheme. */
int mage_scaled_sqdiff_sum_f32(mage_ctx* ctx, const float* a, const float* b, float* out, int64_t n, double scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGE_B200_H */
