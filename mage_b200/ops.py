"""Thin Python wrappers over the C ABI (include/mage_b200.h): torch tensors in, raw device
pointers + the current CUDA stream out.  torch is used only for memory and streams; every
arithmetic operation on the sampling path is a kernel of libmage_sm100.so."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check

ACT_NONE, ACT_RELU, ACT_QUICKGELU, ACT_GELU, ACT_TANH = 0, 1, 2, 3, 4


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda, "libmage_sm100 has no CPU path: tensor must live on a CUDA device"
    return t.data_ptr()


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ctx() -> int:
    """Library handle of the CURRENT CUDA device (one mage_ctx per device, created on first use)."""
    return _lib.ctx(torch.cuda.current_device())


# Optional per-launch CUDA-event instrumentation (bench.py's roofline / breakdown leg): when PROFILE is a
# list, every wrapper appends (kind, algorithmic flops or bytes, start_event, end_event) around its launch.
PROFILE = None


class _Prof:
    def __init__(self, kind: str, flops: float):
        self.kind, self.flops = kind, flops

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(torch.cuda.current_stream())
            PROFILE.append((self.kind, self.flops, self.e0, self.e1))


def launch_count() -> int:
    return int(_lib.lib().mage_launch_count(_ctx()))


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, out: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, res_mod: int = 0, act: int = ACT_NONE, relu_a: bool = False) -> torch.Tensor:
    """out[M,N] = act(a[M,K] @ w[N,K].T + bias) + residual.  `a` may be a row-strided 2-D view."""
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.dtype == torch.float32 and a.stride(1) == 1 and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    assert out.shape == (M, N) and out.stride(1) == 1
    if residual is not None:
        assert residual.dim() == 2 and residual.shape[1] == N and residual.stride(1) == 1
    with _Prof("gemm", 2.0 * M * N * K):
        check(_lib.lib().mage_gemm_f32(_ctx(), _p(a), a.stride(0), _p(_f32(w)), w.stride(0), _p(bias), _p(residual),
                                       residual.stride(0) if residual is not None else 0, res_mod, _p(out), out.stride(0),
                                       M, N, K, act, int(relu_a), _stream()), "mage_gemm_f32")
    return out


# ------------------------------------------------------------------ tensor-core back end (split operands)
_FLAGS = {}


def flag(device) -> torch.Tensor:
    """Per-device int32 that the tensor-core kernels OR with 1 when a value leaves the fp16 split range."""
    key = torch.device(device).index or 0
    f = _FLAGS.get(key)
    if f is None:
        f = _FLAGS[key] = torch.zeros(1, device=device, dtype=torch.int32)
    return f


def check_flag(device) -> None:
    """Raise (loudly, after a sync) if a kernel since the last check reported a condition on which the reference would have
    raised or which invalidates the results: bit 0 = a split conversion saw |x| > 65504 or NaN, bit 1 = a caption token id
    outside the vocabulary table (nn.Embedding raises IndexError, mage_model.py:228)."""
    f = flag(device)
    v = int(f.item())
    if v != 0:
        f.zero_()
        if v & 2:
            raise IndexError("caption token id outside the text encoder's vocabulary (index out of range in self)")
        if v & 4:
            raise IndexError("cross-entropy target outside [0, n_classes) (Target is out of bounds)")
        raise _lib.MageSplitRangeError("a tensor-core operand left the fp16 hi/lo split range (|x| > 65504 or NaN); the results of "
                                       "this call are invalid (MAGE.autoregressive_generate repeats it on the fp32 SIMT kernels)")


def tc_tuning(bn: int = 0, pair: int = -1) -> None:
    """Tile-selection override of the tensor-core kernels (tests / tuning): bn 0|64|128|256, pair -1 auto | 0 | 1."""
    check(_lib.lib().mage_tc_tuning(_ctx(), bn, pair), "mage_tc_tuning")


def pdl(enable: bool) -> None:
    """Programmatic dependent launch for the per-step kernels on/off (see mage_b200.h)."""
    check(_lib.lib().mage_pdl(_ctx(), int(enable)), "mage_pdl")


def temporal_attn_ring(enable: bool) -> None:
    """temporal_attn_step as the persistent ring kernel, or (default) one CTA per unit (see mage_b200.h); same bits."""
    check(_lib.lib().mage_temporal_attn_ring(_ctx(), int(enable)), "mage_temporal_attn_ring")


def sm_share(sms: int = 0) -> None:
    """The launches that follow size their persistent kernels for at most `sms` SMs (0 = the whole GPU): mage_b200.h."""
    check(_lib.lib().mage_sm_share(_ctx(), int(sms)), "mage_sm_share")


def tc_nsplit(mode: int = 1) -> None:
    """N-split 256-wide pair tiles of gemm_tc: 0 off, 1 automatic, 2 whenever legal (tests / tuning)."""
    check(_lib.lib().mage_tc_nsplit(_ctx(), mode), "mage_tc_nsplit")


def tc_conv_halo(enable: bool = True) -> None:
    """Halo mode of the tensor-core convolutions on/off (tests / tuning)."""
    check(_lib.lib().mage_tc_conv_halo(_ctx(), int(enable)), "mage_tc_conv_halo")


def _f16(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float16 and t.is_contiguous() and t.shape[0] == 2, (t.dtype, t.shape, t.stride())
    return t


def split(x: torch.Tensor, relu: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [..., C] (rows may be strided for a 2-D view) -> split fp16 [2, ..., C] (hi plane, lo plane)."""
    C = x.shape[-1]
    if x.dim() == 2:
        rows, ldx = x.shape[0], x.stride(0)
        assert x.stride(1) == 1
    else:
        assert x.is_contiguous()
        rows, ldx = x.numel() // C, C
    if out is None:
        out = torch.empty(2, *x.shape, device=x.device, dtype=torch.float16)
    with _Prof("split", 8.0 * rows * C):
        check(_lib.lib().mage_split_f32(_ctx(), _p(x), ldx, _p(_f16(out)), rows * C, rows, C, int(relu), _p(flag(x.device)), _stream()),
              "mage_split_f32")
    return out


def patch_rows_split(x_nchw: torch.Tensor, kw: int, pad: int) -> torch.Tensor:
    """planar [n,C,H,W] fp32 -> split [2,n,H,W,64] im2row over the kw horizontal taps (channel kx*C + c), see mage_b200.h."""
    n, C, H, W = x_nchw.shape
    out = torch.empty(2, n, H, W, 64, device=x_nchw.device, dtype=torch.float16)
    with _Prof("conv_first", 16.0 * n * H * W * 64):
        check(_lib.lib().mage_patch_rows_split_f32(_ctx(), _p(_f32(x_nchw)), _p(out), n * H * W * 64, n, C, H, W, kw, pad, _stream()),
              "mage_patch_rows_split_f32")
    return out


def s2d_pad_split(x: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """fp32 NHWC [n,H,W,C] -> split [2, n, H/2+1, W/2+1, 4C]: padded space-to-depth (operand of a 4x4 stride-2 pad-1 convolution
    rewritten as a 2x2 stride-1 valid convolution, see mage_b200.h)."""
    n, H, W, C = x.shape
    out = torch.empty(2, n, H // 2 + 1, W // 2 + 1, 4 * C, device=x.device, dtype=torch.float16)
    with _Prof("split", 8.0 * out.numel()):
        check(_lib.lib().mage_s2d_pad_split_f32(_ctx(), _p(_f32(x)), _p(out), out.numel() // 2, n, H, W, C, int(relu), _p(flag(x.device)),
                                                _stream()), "mage_s2d_pad_split_f32")
    return out


def embedding_split(idx: torch.Tensor, table: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """idx int64 [...], table split [2, K, C] -> split [2, ..., C]."""
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    K, C = table.shape[1], table.shape[2]
    rows = idx.numel()
    if out is None:
        out = torch.empty(2, *idx.shape, C, device=table.device, dtype=torch.float16)
    with _Prof("embed", 8.0 * rows * C):
        check(_lib.lib().mage_embedding_split(_ctx(), _p(idx), _p(_f16(table)), K * C, _p(_f16(out)), rows * C, rows, C, _stream()),
              "mage_embedding_split")
    return out


def gemm_tc(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, out: Optional[torch.Tensor] = None,
            out_split: Optional[torch.Tensor] = None, out_split_relu: Optional[torch.Tensor] = None, want=("f32",),
            residual: Optional[torch.Tensor] = None, res_mod: int = 0, act: int = ACT_NONE):
    """Tensor-core GEMM on split operands: a [2,M,K], w [2,N,K] -> act(a @ w.T + bias) + residual.
    `want` lists the outputs to allocate when not passed: "f32", "split", "split_relu".
    Returns (out_f32, out_split, out_split_relu) with None for the ones not produced."""
    _f16(w)
    # `a` may be a row range of a larger split tensor: each plane [M,K] dense, the lo plane a.stride(0) elements after the hi plane
    assert a.dtype == torch.float16 and a.dim() == 3 and a.shape[0] == 2 and a[0].is_contiguous() and a.stride(0) % 8 == 0, (a.shape, a.stride())
    M, K = a.shape[1], a.shape[2]
    N = w.shape[1]
    assert w.shape[2] == K
    dev = a.device
    if out is None and "f32" in want:
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
    if out_split is None and "split" in want:
        out_split = torch.empty(2, M, N, device=dev, dtype=torch.float16)
    if out_split_relu is None and "split_relu" in want:
        out_split_relu = torch.empty(2, M, N, device=dev, dtype=torch.float16)
    if out is not None:
        assert out.shape == (M, N) and out.is_contiguous()
    if residual is not None:
        assert residual.dim() == 2 and residual.shape[1] == N and residual.stride(1) == 1
    with _Prof("gemm", 2.0 * M * N * K):
        check(_lib.lib().mage_gemm_tc(_ctx(), _p(a), K, a.stride(0), _p(w), K, N * K, _p(bias), _p(residual),
                                      residual.stride(0) if residual is not None else 0, res_mod, _p(out), _p(out_split),
                                      _p(out_split_relu), N, M * N, M, N, K, act, _p(flag(dev)), _stream()), "mage_gemm_tc")
    return out, out_split, out_split_relu


def gemm_tc_ln(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, *, residual: torch.Tensor, out: torch.Tensor, gamma: torch.Tensor,
               beta: torch.Tensor, ln_out: torch.Tensor, counters: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """out[M,512] = a @ w.T + bias + residual (may alias out) and ln_out = split(LayerNorm(out)) in ONE launch (mage_b200.h:
    mage_gemm_tc_ln).  counters: int32 [>= ceil(M/128)], zero (the kernel leaves them zero)."""
    _f16(a), _f16(w), _f16(ln_out)
    M, K = a.shape[1], a.shape[2]
    assert w.shape[1] == 512 and w.shape[2] == K and tuple(out.shape) == (M, 512) and out.is_contiguous() and tuple(ln_out.shape[1:]) == (M, 512)
    assert residual.shape == out.shape and residual.stride(1) == 1 and counters.dtype == torch.int32 and counters.numel() >= (M + 127) // 128
    with _Prof("gemm", 2.0 * M * 512 * K):
        check(_lib.lib().mage_gemm_tc_ln(_ctx(), _p(a), K, M * K, _p(w), K, 512 * K, _p(bias), _p(residual), residual.stride(0), _p(out), M, K,
                                         _p(_f32(gamma)), _p(_f32(beta)), eps, _p(ln_out), M * 512, _p(counters), _p(flag(a.device)),
                                         _stream()), "mage_gemm_tc_ln")
    return out


def permute_qkv_for_axial(w_in: torch.Tensor, b_in: torch.Tensor, n_head: int):
    """nn.MultiheadAttention's packed in-projection (rows [q(C) | k(C) | v(C)], mage_model.py:20) re-ordered for
    mage_qkv_axial_attn_tc: every 192-row tile = [q|k|v] x 32 of two heads.  Returns (weight [3C, K], bias [3C])."""
    C = w_in.shape[0] // 3
    idx = torch.arange(3 * C, device=w_in.device).view(3, n_head, 32)          # [part, head, d] -> old row
    perm = idx.permute(1, 0, 2).reshape(-1)                                     # new order: head, part, d
    return w_in[perm].contiguous(), b_in[perm].contiguous()


def qkv_axial_attn_tc(a: torch.Tensor, w_perm: torch.Tensor, bias_perm: torch.Tensor, out_split: torch.Tensor, *, n_img: int, R: int,
                      n_head: int, axis: int, scale: float) -> torch.Tensor:
    """Fused QKV projection + H (axis=1) / W (axis=2) axial attention: a split [2, n_img*R*R, K] -> out_split [2, n_img*R*R, C]
    (see mage_b200.h).  `out_split` must not alias `a`."""
    _f16(a), _f16(w_perm), _f16(out_split)
    M, K = a.shape[1], a.shape[2]
    C = n_head * 32
    assert M == n_img * R * R and tuple(w_perm.shape[1:]) == (3 * C, K) and tuple(out_split.shape[1:]) == (M, C)
    assert out_split.data_ptr() != a.data_ptr()
    with _Prof("gemm", 2.0 * M * 3 * C * K):
        check(_lib.lib().mage_qkv_axial_attn_tc(_ctx(), _p(a), M * K, _p(w_perm), 3 * C * K, _p(_f32(bias_perm)), _p(out_split), M * C,
                                                n_img, R, n_head, K, axis, scale, _p(flag(a.device)), _stream()),
              "mage_qkv_axial_attn_tc")
    return out_split


def conv2d_tc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, pad=(0, 0),
              residual: Optional[torch.Tensor] = None, res_mode: int = 0, act: int = ACT_NONE, want=("f32",),
              out: Optional[torch.Tensor] = None, out_split: Optional[torch.Tensor] = None,
              out_split_relu: Optional[torch.Tensor] = None, out_hw=None, scatter=(1, 1, 0, 0), full_hw=None, passes: int = 3):
    """Tensor-core stride-1 NHWC convolution on split operands: x [2,n,Hin,Win,Cin], w [2,Cout,KH,KW,Cin].
    passes=1: hi*hi products only (see mage_b200.h) -- decoder layers that feed no token."""
    _f16(x), _f16(w)
    _, n, Hin, Win, Cin = x.shape
    _, Cout, KH, KW, Cin2 = w.shape
    assert Cin == Cin2
    if out_hw is None:
        out_hw = (Hin + 2 * pad[0] - KH + 1, Win + 2 * pad[1] - KW + 1)
    Hout, Wout = out_hw
    sy, sx, oy, ox = scatter
    if full_hw is None:
        full_hw = (Hout * sy, Wout * sx)
    Hfull, Wfull = full_hw
    dev = x.device
    if out is None and "f32" in want:
        out = torch.empty(n, Hfull, Wfull, Cout, device=dev, dtype=torch.float32)
    if out_split is None and "split" in want:
        out_split = torch.empty(2, n, Hfull, Wfull, Cout, device=dev, dtype=torch.float16)
    if out_split_relu is None and "split_relu" in want:
        out_split_relu = torch.empty(2, n, Hfull, Wfull, Cout, device=dev, dtype=torch.float16)
    if residual is not None and res_mode == 0:
        res_mode = 1
    img = Hfull * Wfull * Cout
    with _Prof("conv", 2.0 * n * Hout * Wout * Cout * KH * KW * Cin):
        check(_lib.lib().mage_conv2d_tc(_ctx(), _p(x), n * Hin * Win * Cin, _p(w), Cout * KH * KW * Cin, _p(bias), _p(residual), _p(out),
                                        _p(out_split), _p(out_split_relu), n * img, n, Hin, Win, Cin, Hout, Wout, Cout, KH, KW,
                                        pad[0], pad[1], res_mode, act, sy, sx, oy, ox, Hfull, Wfull, img, passes, _p(flag(dev)),
                                        _stream()), "mage_conv2d_tc")
    return out, out_split, out_split_relu


def conv2d_tc_pixel_head(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, pad, residual: Optional[torch.Tensor],
                         res_mode: int, head_w: torch.Tensor, head_b: torch.Tensor, out: torch.Tensor, out_img_stride: int,
                         passes: int = 3) -> None:
    """conv2d_tc whose 256-channel result never leaves the SM: tanh(head_b + head_w . relu(conv + bias + residual)) is written
    planar at out.data_ptr() + img*out_img_stride (vqvae_model.py:210-213)."""
    _f16(x), _f16(w)
    _, n, Hin, Win, Cin = x.shape
    _, Cout, KH, KW, Cin2 = w.shape
    assert Cin == Cin2 and head_w.shape == (head_b.numel(), Cout) and head_w.is_contiguous()
    Hout, Wout = Hin + 2 * pad[0] - KH + 1, Win + 2 * pad[1] - KW + 1
    with _Prof("conv", 2.0 * n * Hout * Wout * Cout * (KH * KW * Cin + head_b.numel())):
        check(_lib.lib().mage_conv2d_tc_pixel_head(_ctx(), _p(x), n * Hin * Win * Cin, _p(w), Cout * KH * KW * Cin, _p(bias), _p(residual),
                                                   n, Hin, Win, Cin, Hout, Wout, Cout, KH, KW, pad[0], pad[1], res_mode,
                                                   _p(head_w), _p(head_b), head_b.numel(), _p(out), out_img_stride, passes,
                                                   _p(flag(x.device)), _stream()), "mage_conv2d_tc_pixel_head")


def conv2d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, stride: int = 1, pad=(0, 0),
           in_up: bool = False, residual: Optional[torch.Tensor] = None, res_mode: int = 0, relu_in: bool = False,
           act: int = ACT_NONE, out: Optional[torch.Tensor] = None, out_hw=None, scatter=(1, 1, 0, 0),
           full_hw=None, out_img_stride: Optional[int] = None) -> torch.Tensor:
    """NHWC implicit-GEMM convolution.  x [N,Hin,Win,Cin], w [Cout,KH,KW,Cin] -> [N,Hout,Wout,Cout]
    (or a scatter into `out` [N,Hfull,Wfull,Cout] for the sub-pixel phases of a transposed conv)."""
    n, Hin, Win, Cin = x.shape
    Cout, KH, KW, Cin2 = w.shape
    assert Cin == Cin2
    Hl, Wl = (Hin * 2, Win * 2) if in_up else (Hin, Win)
    if out_hw is None:
        out_hw = ((Hl + 2 * pad[0] - KH) // stride + 1, (Wl + 2 * pad[1] - KW) // stride + 1)
    Hout, Wout = out_hw
    sy, sx, oy, ox = scatter
    if full_hw is None:
        full_hw = (Hout * sy, Wout * sx)
    Hfull, Wfull = full_hw
    if out is None:
        out = torch.empty(n, Hfull, Wfull, Cout, device=x.device, dtype=torch.float32)
    if out_img_stride is None:
        out_img_stride = Hfull * Wfull * Cout
    if residual is not None and res_mode == 0:
        res_mode = 1
    with _Prof("conv", 2.0 * n * Hout * Wout * Cout * KH * KW * Cin):
        check(_lib.lib().mage_conv2d_nhwc_f32(_ctx(), _p(_f32(x)), _p(_f32(w)), _p(bias), _p(residual), _p(out), n, Hin, Win, Cin,
                                              Hout, Wout, Cout, KH, KW, stride, pad[0], pad[1], int(in_up), res_mode,
                                              int(relu_in), act, sy, sx, oy, ox, Hfull, Wfull, out_img_stride, _stream()),
              "mage_conv2d_nhwc_f32")
    return out


def conv2d_first(x_nchw: torch.Tensor, w_t: torch.Tensor, bias: Optional[torch.Tensor], *, cout: int, kh: int, kw: int,
                 stride: int, pad: int, act: int = ACT_NONE) -> torch.Tensor:
    n, Cin, H, W = x_nchw.shape
    Hout = (H + 2 * pad - kh) // stride + 1
    Wout = (W + 2 * pad - kw) // stride + 1
    out = torch.empty(n, Hout, Wout, cout, device=x_nchw.device, dtype=torch.float32)
    with _Prof("conv_first", 4.0 * out.numel()):
        check(_lib.lib().mage_conv2d_first_f32(_ctx(), _p(_f32(x_nchw)), _p(_f32(w_t)), _p(bias), _p(out), n, Cin, H, W, Hout, Wout,
                                               cout, kh, kw, stride, pad, act, _stream()), "mage_conv2d_first_f32")
    return out


def conv1x1_tanh_nchw(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, out_img_stride: int) -> None:
    """x NHWC [N,H,W,Cin] -> tanh(conv1x1(relu(x))) written planar into `out` (image stride in elements)."""
    n, H, W, Cin = x.shape
    with _Prof("conv1x1_tanh", 4.0 * x.numel()):
        check(_lib.lib().mage_conv1x1_tanh_nchw_f32(_ctx(), _p(_f32(x)), _p(_f32(w)), _p(bias), _p(out), n, H * W, Cin, w.shape[0],
                                                    out_img_stride, _stream()), "mage_conv1x1_tanh_nchw_f32")


def maxpool2x2(x: torch.Tensor) -> torch.Tensor:
    n, H, W, C = x.shape
    out = torch.empty(n, H // 2, W // 2, C, device=x.device, dtype=torch.float32)
    with _Prof("maxpool", 5.0 * x.numel()):
        check(_lib.lib().mage_maxpool2x2_nhwc_f32(_ctx(), _p(_f32(x)), _p(out), n, H, W, C, _stream()), "mage_maxpool2x2_nhwc_f32")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None, out_split: Optional[torch.Tensor] = None):
    """LayerNorm over the last dim.  With `out_split` (fp16 [2, rows, C]) the result is written in the split
    operand format of the tensor-core GEMMs (and the fp32 copy only if `out` is given as well)."""
    C = x.shape[-1]
    rows = x.numel() // C
    if out is None and out_split is None:
        out = torch.empty_like(x)
    with _Prof("layernorm", 8.0 * x.numel()):
        check(_lib.lib().mage_layernorm_f32(_ctx(), _p(_f32(x)), _p(gamma), _p(beta), _p(out), _p(out_split), rows * C, _p(flag(x.device)),
                                            rows, C, eps, _stream()), "mage_layernorm_f32")
    return out if out_split is None else out_split


def mha(q, k, v, out, *, n_outer, n_inner, n_head, Sq, Sk, q_strides, k_strides, v_strides, o_strides,
        key_len: Optional[torch.Tensor] = None, scale: float, out_split: Optional[torch.Tensor] = None) -> None:
    """Strided SDPA core (head_dim 32).  *_strides = (outer, inner, seq) in elements; q/k/v/out may be
    views into one packed qkv buffer (pass the view: its data_ptr carries the column offset)."""
    with _Prof("mha", 0.0):
        check(_lib.lib().mage_mha_f32(_ctx(), _p(q), _p(k), _p(v), _p(out), n_outer, n_inner, n_head, Sq, Sk, *q_strides, *k_strides,
                                      *v_strides, *o_strides, _p(key_len), scale, _p(out_split),
                                      out_split.numel() // 2 if out_split is not None else 0, _p(flag(q.device)), _stream()),
              "mage_mha_f32")


def axial_attn(qkv: torch.Tensor, out: Optional[torch.Tensor], *, B: int, R: int, n_head: int, axis: int, scale: float,
               out_split: Optional[torch.Tensor] = None) -> None:
    """H (axis=1) / W (axis=2) axial attention of one temporal position: qkv [B*R*R, 3C] -> out [B*R*R, C] and/or split."""
    with _Prof("axial_attn", 4.0 * qkv.numel() + 4.0 * qkv.numel() / 3):
        check(_lib.lib().mage_axial_attn_f32(_ctx(), _p(_f32(qkv)), _p(out), _p(out_split),
                                             out_split.numel() // 2 if out_split is not None else 0, _p(flag(qkv.device)),
                                             B, R, n_head, axis, scale, _stream()), "mage_axial_attn_f32")


def temporal_attn_step(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, out: Optional[torch.Tensor], pos: int,
                       scale: float, out_split: Optional[torch.Tensor] = None) -> None:
    M = qkv.shape[0]
    Lmax = kcache.shape[1]
    if out_split is not None:   # may be a row range of a larger split tensor: the lo plane sits stride(0) elements after the hi plane
        assert out_split.dtype == torch.float16 and out_split.shape[0] == 2 and out_split[0].is_contiguous()
    with _Prof("temporal_attn", 8.0 * M * (pos + 1) * kcache.shape[2]):
        check(_lib.lib().mage_temporal_attn_step_f32(_ctx(), _p(_f32(qkv)), _p(kcache), _p(vcache), _p(out), _p(out_split),
                                                     out_split.stride(0) if out_split is not None else 0, _p(flag(qkv.device)),
                                                     M, pos, Lmax, scale, _stream()), "mage_temporal_attn_step_f32")


def temporal_attn_seq(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, pos0: int, n_pos: int, scale: float,
                      out_split: torch.Tensor) -> None:
    """n_pos consecutive temporal positions in one launch: qkv [n_pos*M, 3C] (position-major), out_split [2, n_pos*M, C]."""
    M = kcache.shape[0]
    assert qkv.shape[0] == n_pos * M and _f16(out_split).shape[1] == n_pos * M
    with _Prof("temporal_attn", 8.0 * M * kcache.shape[2] * sum(pos0 + s + 1 for s in range(n_pos)) / max(n_pos, 1) * 1.0):
        check(_lib.lib().mage_temporal_attn_seq_f32(_ctx(), _p(_f32(qkv)), _p(kcache), _p(vcache), _p(out_split), out_split.stride(0),
                                                    _p(flag(qkv.device)), M, pos0, n_pos, kcache.shape[1], scale, _stream()),
              "mage_temporal_attn_seq_f32")


def kv_append(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, pos: int) -> None:
    M, C3 = qkv.shape
    with _Prof("kv_append", 0.0):
        check(_lib.lib().mage_kv_append_f32(_ctx(), _p(_f32(qkv)), _p(kcache), _p(vcache), M, C3 // 3, pos, kcache.shape[1], _stream()),
              "mage_kv_append_f32")


def vq_argmin(z: torch.Tensor, codebook: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """z [N,D] fp32, codebook [K,D] -> int64 [N] nearest-code indices (vqvae_model.py:8-25)."""
    N, D = z.shape
    K = codebook.shape[0]
    if out is None:
        out = torch.empty(N, device=z.device, dtype=torch.int64)
    scratch = torch.empty(K, device=z.device, dtype=torch.float32)
    with _Prof("vq_argmin", 4.0 * N * D):
        check(_lib.lib().mage_vq_argmin_f32(_ctx(), _p(_f32(z)), _p(_f32(codebook)), _p(scratch), _p(out), N, D, K, _stream()),
              "mage_vq_argmin_f32")
    return out


def argmax_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows, N = x.shape
    if out is None:
        out = torch.empty(rows, device=x.device, dtype=torch.int64)
    with _Prof("argmax", 4.0 * rows * N):
        check(_lib.lib().mage_argmax_rows_f32(_ctx(), _p(x), x.stride(0), _p(out), rows, N, _stream()), "mage_argmax_rows_f32")
    return out


def embedding(idx: torch.Tensor, table: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    rows, C = idx.numel(), table.shape[1]
    if out is None:
        out = torch.empty(*idx.shape, C, device=table.device, dtype=torch.float32)
    with _Prof("embed", 8.0 * rows * C):
        check(_lib.lib().mage_embedding_f32(_ctx(), _p(idx), _p(_f32(table)), _p(out), rows, C, _stream()), "mage_embedding_f32")
    return out


def token_taps_ln(tok: torch.Tensor, table: torch.Tensor, pos_bias: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, gamma: torch.Tensor,
                  beta: torch.Tensor, ln_out: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """token_taps + the first block's ln_1 in one launch: out fp32 [n*R*R, C] and ln_out = split(LayerNorm(out)) [2, n*R*R, C]."""
    n, R, _ = tok.shape
    taps, K, C = table.shape
    kh = int(round(taps ** 0.5))
    assert kh * kh == taps and tok.dtype == torch.int64 and tok.is_contiguous() and out.is_contiguous() and _f16(ln_out).shape[1] == n * R * R
    with _Prof("embed", 4.0 * n * R * R * C * (taps + 2)):
        check(_lib.lib().mage_token_taps_ln_f32(_ctx(), _p(tok), _p(_f32(table)), _p(_f32(pos_bias)), _p(_f32(bias)), _p(out), n, R, K, C, kh,
                                                kh, _p(_f32(gamma)), _p(_f32(beta)), eps, _p(ln_out), n * R * R * C, _p(flag(out.device)),
                                                _stream()), "mage_token_taps_ln_f32")
    return out


def token_taps(tok: torch.Tensor, table: torch.Tensor, pos_bias: torch.Tensor, bias: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[b,y,x,:] = sum_taps table[tap][tok[b, y+dy, x+dx]] + pos_bias[y*R+x] + bias (see mage_b200.h).
    tok int64 [n,R,R]; table [KH*KW, K, C]; out fp32 [n*R*R, C]."""
    n, R, _ = tok.shape
    taps, K, C = table.shape
    kh = int(round(taps ** 0.5))
    assert kh * kh == taps and tok.dtype == torch.int64 and tok.is_contiguous() and out.is_contiguous()
    with _Prof("embed", 4.0 * n * R * R * C * (taps + 2)):
        check(_lib.lib().mage_token_taps_f32(_ctx(), _p(tok), _p(_f32(table)), _p(_f32(pos_bias)), _p(_f32(bias)), _p(out), n, R, K, C, kh, kh,
                                             _stream()), "mage_token_taps_f32")
    return out


def text_embed(text: torch.Tensor, tok_emb, pos_emb, gamma, beta, pad_idx: int, eps: float):
    B, T = text.shape
    C = tok_emb.shape[1]
    if T > pos_emb.shape[0]:
        # nn.Embedding would raise on position indices past the table (mage_model.py:228-231); never read out of bounds
        raise IndexError(f"caption length {T} exceeds the text encoder's context_length {pos_emb.shape[0]}")
    x = torch.empty(B, T, C, device=tok_emb.device, dtype=torch.float32)
    key_len = torch.empty(B, device=tok_emb.device, dtype=torch.int32)
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_text_embed_f32(_ctx(), _p(text), _p(tok_emb), _p(pos_emb), _p(gamma), _p(beta), _p(x), _p(key_len), B, T, C,
                                             pad_idx, eps, tok_emb.shape[0], _p(flag(tok_emb.device)), _stream()), "mage_text_embed_f32")
    return x, key_len


def adain(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    n, H, W, C = x.shape
    out = torch.empty_like(x)
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_adain_nhwc_f32(_ctx(), _p(_f32(x)), _p(_f32(gamma)), _p(_f32(beta)), _p(out), n, H * W, C, eps, _stream()),
              "mage_adain_nhwc_f32")
    return out


def add_scaled_vec(x: torch.Tensor, s: torch.Tensor, vec: torch.Tensor) -> None:
    n, C = x.shape[0], x.shape[-1]
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_add_scaled_vec_f32(_ctx(), _p(_f32(x)), _p(s), _p(vec), n, x.numel() // (n * C), C, _stream()),
              "mage_add_scaled_vec_f32")


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    n, C, H, W = x.shape
    out = torch.empty(n, H, W, C, device=x.device, dtype=torch.float32)
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_nchw_to_nhwc_f32(_ctx(), _p(_f32(x)), _p(out), n, C, H * W, _stream()), "mage_nchw_to_nhwc_f32")
    return out


def gn_partial(x: torch.Tensor, part: torch.Tensor, B: int, HW: int, groups: int = 32) -> None:
    """x fp32 [n_slots*B*HW, C] -> part f64 [n_slots, B, groups, 2] (sum, sum of squares per slot / sample / group)."""
    C = x.shape[-1]
    n_slots = x.numel() // (B * HW * C)
    assert part.dtype == torch.float64 and part.is_contiguous() and part.numel() == n_slots * B * groups * 2
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_gn_partial_f32(_ctx(), _p(_f32(x)), _p(part), n_slots, B, HW, C, groups, _stream()), "mage_gn_partial_f32")


def gn_apply(x: torch.Tensor, part: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, B: int, HW: int, *, relu: bool,
             residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_split: Optional[torch.Tensor] = None,
             eps: float = 1e-5) -> None:
    """GroupNorm over (channel group x all slots x HW) of each sample from gn_partial's sums (part f64 [n_slots,B,groups,2]):
    x fp32 [n_slots*B*HW, 512] -> (+ residual) (ReLU) -> out fp32 (may alias x) and/or out_split fp16 [2, rows, 512]."""
    rows, C = x.shape
    n_slots, groups = part.shape[0], part.shape[2]
    assert rows == n_slots * B * HW and part.dtype == torch.float64 and part.is_contiguous() and (out is not None or out_split is not None)
    stat = torch.empty(B * groups * 2, device=x.device, dtype=torch.float32)
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_gn_apply_f32(_ctx(), _p(_f32(x)), _p(part), _p(stat), _p(_f32(gamma)), _p(_f32(beta)),
                                           _p(_f32(residual)) if residual is not None else None, _p(out), _p(out_split), rows * C,
                                           _p(flag(x.device)), n_slots, B, HW, C, groups, int(relu), eps, _stream()), "mage_gn_apply_f32")


def cross_entropy_rows(logits: torch.Tensor, target: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[row] = logsumexp(logits[row]) - logits[row, target[row]]  (F.cross_entropy(reduction='none'))."""
    rows, K = logits.shape
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and target.dtype == torch.int64 and target.is_contiguous()
    assert target.numel() == rows and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == rows
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_cross_entropy_rows_f32(_ctx(), _p(logits), logits.stride(0), _p(target), _p(out), rows, K,
                                                     _p(flag(logits.device)), _stream()), "mage_cross_entropy_rows_f32")
    return out


def reparam_kl(mu_logvar: torch.Tensor, eps: Optional[torch.Tensor], B: int, HW: int, want_z: bool = True):
    """mu_logvar fp32 [B*HW, 2*Cz] (conv_mu2 | conv_var2) and the stored draw eps [B, Cz, h, w] ->
    (z = eps * exp(0.5 logvar) + mu, NCHW like eps; kl_rows [B] = sum(1 + logvar - mu^2 - exp(logvar)))."""
    Cz = mu_logvar.shape[-1] // 2
    assert mu_logvar.is_contiguous() and mu_logvar.numel() == B * HW * 2 * Cz
    z = torch.empty(B, Cz, HW, device=mu_logvar.device, dtype=torch.float32) if want_z else None
    kl_rows = torch.empty(B, device=mu_logvar.device, dtype=torch.float32)
    if want_z:
        assert eps is not None and eps.is_contiguous() and eps.numel() == B * Cz * HW
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_reparam_kl_f32(_ctx(), _p(_f32(mu_logvar)), _p(_f32(eps)) if want_z else None, _p(z), _p(kl_rows), B, HW, Cz,
                                             _stream()), "mage_reparam_kl_f32")
    return z, kl_rows


def scaled_sum(x: torch.Tensor, scale: float) -> torch.Tensor:
    """scale * sum(x) as a device scalar: one block, fixed summation order, double accumulation."""
    assert x.is_contiguous()
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_scaled_sum_f32(_ctx(), _p(_f32(x)), _p(out), x.numel(), float(scale), _stream()), "mage_scaled_sum_f32")
    return out


def scaled_sqdiff_sum(a: torch.Tensor, b: torch.Tensor, scale: float) -> torch.Tensor:
    """scale * sum((a - b)^2) as a device scalar (F.mse_loss with scale = 1/numel): one block, fixed order, double accumulation."""
    assert a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()
    out = torch.empty(1, device=a.device, dtype=torch.float32)
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_scaled_sqdiff_sum_f32(_ctx(), _p(_f32(a)), _p(_f32(b)), _p(out), a.numel(), float(scale), _stream()),
              "mage_scaled_sqdiff_sum_f32")
    return out


def gn_silu_head(x: torch.Tensor, part: torch.Tensor, gamma, beta, w: torch.Tensor, bias: torch.Tensor, B: int, HW: int,
                 eps: float = 1e-5) -> torch.Tensor:
    """GroupNorm(32) (statistics over all slots of `part` [n_slots,B,32,2]) -> SiLU -> linear to w.shape[0] channels, rows of x."""
    rows, C = x.shape
    cout = w.shape[0]
    out = torch.empty(rows, cout, device=x.device, dtype=torch.float32)
    with _Prof("misc", 0.0):
        check(_lib.lib().mage_gn_silu_head_f32(_ctx(), _p(_f32(x)), _p(part), _p(_f32(gamma)), _p(_f32(beta)), _p(_f32(w)), _p(_f32(bias)),
                                               _p(out), rows, B, HW, part.shape[0], C, 32, cout, eps, _stream()), "mage_gn_silu_head_f32")
    return out
