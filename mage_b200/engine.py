"""Host-side orchestration of the MAGE sampling path on one B200.

Two engines, both driven by the reference's state-dict layout:

* `VQVAEEngine`  -- VectorQuantizedVAE.encode / .decode (vqvae_model.py:233-242) on NHWC tensors:
  conv stacks as implicit GEMMs, nearest-upsample folded into the consumers' address math,
  eval-mode BatchNorm folded into the neighbouring conv at load, fused VQ argmin.
* `SamplerEngine` -- MAGE.autoregressive_generate (mage_model.py:641-693) re-designed as the
  incremental algorithm of SURVEY.md App. D: per step only the newest temporal position runs
  through the six axial blocks; the two temporal blocks keep a K/V cache `[B*256, L, 512]`.

Python only sequences kernels of libmage_sm100.so (mage_b200.ops) and owns buffers; the one-time
weight re-layout at load (permute / BatchNorm folding) uses torch tensor ops and is not on the
timed path.  A whole `generate` call is captured into a CUDA graph per input signature.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_QUICKGELU, ACT_RELU, ACT_TANH

ACT_POST = 0x100
RES_RELU = 0x200


def default_backend() -> str:
    """"tc": dense contractions on tcgen05 (fp16 hi/lo split operands, fp32-grade); "simt": fp32 FFMA kernels."""
    b = os.environ.get("MAGE_BACKEND", "tc")
    assert b in ("tc", "simt"), b
    return b


def _upsample_phase_weights(w: torch.Tensor):
    """3x3 conv over a nearest-x2-upsampled map == four 2x2 sub-pixel phase convs over the stored map
    (vqvae_model.py:205-209 without materialising the upsample).  w [Cout,Cin,3,3] -> {(py,px): [Cout,2,2,Cin]}:
    output row 2y+py reads stored rows (y-1, y) with weights (w0, w1+w2) for py=0 and rows (y, y+1) with
    (w0+w1, w2) for py=1 (same along x); sums are formed in fp64 and rounded once to fp32."""
    wd = w.double()
    rows = {0: [wd[:, :, 0], wd[:, :, 1] + wd[:, :, 2]], 1: [wd[:, :, 0] + wd[:, :, 1], wd[:, :, 2]]}  # each [Cout,Cin,3(kx)]
    out = {}
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for r in rows[py]:
                cols = [r[:, :, 0], r[:, :, 1] + r[:, :, 2]] if px == 0 else [r[:, :, 0] + r[:, :, 1], r[:, :, 2]]
                taps.append(torch.stack(cols, dim=1))  # [Cout,2(kx),Cin]
            out[(py, px)] = torch.stack(taps, dim=1).float().contiguous()  # [Cout,2(ky),2(kx),Cin]
    return out


def _pack_conv(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,KH,KW] -> [Cout,KH,KW,Cin] (K-contiguous rows for the implicit GEMM)."""
    return w.permute(0, 2, 3, 1).contiguous()


def _bn_fold(sd, name, eps=1e-5):
    s = sd[name + ".weight"] / torch.sqrt(sd[name + ".running_var"] + eps)
    return s, sd[name + ".bias"] - sd[name + ".running_mean"] * s


class VQVAEEngine:
    def __init__(self, sd: Dict[str, torch.Tensor], backend: Optional[str] = None):
        """sd: VectorQuantizedVAE.state_dict() tensors (fp32, on the CUDA device)."""
        self.device = sd["codebook.embedding.weight"].device
        assert self.device.type == "cuda", "VQVAEEngine needs CUDA tensors (no CPU path)"
        self.down_ratio = 4 if "encoder.1.running_mean" in sd else 8
        self.backend = backend or default_backend()
        ops.flag(self.device)
        # Decoder precision budget: the decoder feeds no token and its bar is 1e-3 relative on pixels (north_star),
        # so the two convolutions that hold 63 % of its FLOPs -- block.5 and block.7 (+ fused pixel head) of the last, 128x128
        # block -- run single-pass (fp16 operands, fp32 accumulation) instead of the 3-MMA fp32-grade product: measured pixel
        # rel-L2 3.6e-4 .. 5.8e-4 against fp32 (tools/experiments/decoder_precision_emulation.py, tests::test_decoder_precision_budget).
        # "fp32" keeps every layer fp32-grade (pixels ~1e-6).  The encoder and everything that feeds a token is always fp32-grade.
        self.decoder_precision = os.environ.get("MAGE_DECODER_PRECISION", "budget")
        assert self.decoder_precision in ("budget", "fp32"), self.decoder_precision
        self.codebook = sd["codebook.embedding.weight"].contiguous()
        self.K, self.D = self.codebook.shape
        w = {}
        if self.down_ratio == 8:
            w["enc0_wt"] = sd["encoder.0.weight"].permute(1, 2, 3, 0).reshape(-1, sd["encoder.0.weight"].shape[0]).contiguous()
            w["enc0_b"] = sd["encoder.0.bias"].contiguous()
            self.in_ch = sd["encoder.0.weight"].shape[1]
            for k, v in sd.items():
                if k.endswith(".weight") and v.dim() == 4 and k != "encoder.0.weight":
                    w[k] = _pack_conv(v) if v.shape[-1] > 1 else v.reshape(v.shape[0], v.shape[1]).contiguous()
                elif k.endswith(".bias"):
                    w[k] = v.contiguous()
        else:
            self.in_ch = sd["encoder.0.weight"].shape[1]
            s, b = _bn_fold(sd, "encoder.1")
            w0 = sd["encoder.0.weight"] * s.view(-1, 1, 1, 1)
            w["enc0_wt"] = w0.permute(1, 2, 3, 0).reshape(-1, w0.shape[0]).contiguous()
            w["enc0_b"] = (sd["encoder.0.bias"] * s + b).contiguous()
            w["encoder.3.weight"] = _pack_conv(sd["encoder.3.weight"])
            w["encoder.3.bias"] = sd["encoder.3.bias"].contiguous()
            for blk in ("encoder.4", "encoder.5", "decoder.0", "decoder.1"):
                s1, b1 = _bn_fold(sd, blk + ".block.2")
                w[blk + ".c3.w"] = _pack_conv(sd[blk + ".block.1.weight"] * s1.view(-1, 1, 1, 1))
                w[blk + ".c3.b"] = (sd[blk + ".block.1.bias"] * s1 + b1).contiguous()
                s2, b2 = _bn_fold(sd, blk + ".block.5")
                w[blk + ".c1.w"] = (sd[blk + ".block.4.weight"][:, :, 0, 0] * s2.view(-1, 1)).contiguous()
                w[blk + ".c1.b"] = (sd[blk + ".block.4.bias"] * s2 + b2).contiguous()
            # ConvTranspose2d(4,2,1): out[2y+py] gets taps (input offset, k): py=0 -> (-1,3),(0,1); py=1 -> (0,2),(+1,0)
            s3, b3 = _bn_fold(sd, "decoder.4")
            w["decoder.3.b"] = (sd["decoder.3.bias"] * s3 + b3).contiguous()
            w["decoder.6.b"] = sd["decoder.6.bias"].contiguous()
            taps = {0: (3, 1), 1: (2, 0)}
            for name, scale in (("decoder.3", s3), ("decoder.6", None)):
                wt = sd[name + ".weight"]  # [Cin, Cout, 4, 4]
                if scale is not None:
                    wt = wt * scale.view(1, -1, 1, 1)
                for py in (0, 1):
                    for px in (0, 1):
                        sub = wt[:, :, list(taps[py]), :][:, :, :, list(taps[px])]  # [Cin,Cout,2,2]
                        w[f"{name}.p{py}{px}"] = sub.permute(1, 2, 3, 0).contiguous()  # [Cout,2,2,Cin]
        self.w = w
        if self.backend == "tc":
            if self.down_ratio == 8:
                self._prepare_tc(sd)
            else:
                self._prepare_tc_f4(sd)

    def _prepare_tc_f4(self, sd):
        """Split copies of the f4 (MNIST) stack's tensor-core operands (vqvae_model.py:172-189): the 4x4 stride-2 convolution
        encoder.3 as a 2x2 valid convolution over the padded space-to-depth map (mage_s2d_pad_split_f32), the ResBlocks' BN-folded
        3x3 / 1x1 weights, the four 2x2 sub-pixel phases of ConvTranspose decoder.3, the (ReLU-ed) codebook.  The two single-channel
        ends -- encoder.0 (1 -> 256, K = 16) and decoder.6 (256 -> 1) -- are 8 MFLOP per image each and stay on the FFMA kernels."""
        w, ws = self.w, {}
        w3 = sd["encoder.3.weight"]                                   # [Cout, Cin, 4, 4], ky = 2*ty + py, kx = 2*tx + px
        co, ci = w3.shape[:2]
        w2 = w3.view(co, ci, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(co, 2, 2, 4 * ci).contiguous()   # [co, ty, tx, (py,px,c)]
        ws["encoder.3.s2d"] = ops.split(w2)
        for blk in ("encoder.4", "encoder.5", "decoder.0", "decoder.1"):
            ws[blk + ".c3.w"] = ops.split(w[blk + ".c3.w"])
            ws[blk + ".c1.w"] = ops.split(w[blk + ".c1.w"])
        for py in (0, 1):
            for px in (0, 1):
                ws[f"decoder.3.p{py}{px}"] = ops.split(w[f"decoder.3.p{py}{px}"])
        self.ws = ws
        self.cb_relu = torch.relu(self.codebook).contiguous()
        self.cb_relu_split = ops.split(self.codebook, relu=True)

    def _prepare_tc(self, sd):
        """Split (fp16 hi/lo) copies of every tensor-core operand: conv / 1x1 weights, the codebook (raw and
        ReLU-ed: DecoderBlock reads both, vqvae_model.py:150-156), phase kernels of the upsample-reading convs."""
        ws = {}
        for k, v in self.w.items():
            if k.endswith(".weight") and not k.startswith("enc0"):
                ws[k] = ops.split(v)
        for name in ("decoder.2", "decoder.4", "decoder.6"):
            for ph, wp in _upsample_phase_weights(sd[name + ".block.3.weight"]).items():
                ws[f"{name}.block.3.p{ph[0]}{ph[1]}"] = ops.split(wp)
        # encoder.0 (7x7, pad 3, 3 -> dim) as a 7x1 tensor-core convolution over the im2row'ed image: w2[co,ky,0,kx*C+c] = w[co,c,ky,kx]
        w0 = sd["encoder.0.weight"]
        co, ci, kh, kw = w0.shape
        if ci * kw <= 64:
            w2 = torch.zeros(co, kh, 1, 64, device=w0.device, dtype=torch.float32)
            w2[:, :, 0, : kw * ci] = w0.permute(0, 2, 3, 1).reshape(co, kh, kw * ci)
            ws["enc0_rows"] = ops.split(w2)
            self.enc0_kw, self.enc0_pad = kw, kw // 2
        self.ws = ws
        self.cb_split = ops.split(self.codebook)
        self.cb_relu_split = ops.split(self.codebook, relu=True)
        # decoder.0 reads codebook rows, so its two 1x1 convolutions (id_path 4dim->2dim, block.1 4dim->hid) are functions of the
        # code index alone: evaluate them once per code with the same kernels (rows of a GEMM are independent, so the per-pixel
        # results are bit-identical) and gather at decode time
        self.dec0_tables = None
        if ("decoder.0.id_path.weight") in self.w:
            idp_tab, _, _ = ops.gemm_tc(self.cb_split, ws["decoder.0.id_path.weight"], self.w["decoder.0.id_path.bias"])
            _, h_tab, _ = ops.gemm_tc(self.cb_relu_split, ws["decoder.0.block.1.weight"], self.w["decoder.0.block.1.bias"], act=ACT_RELU,
                                      want=("split",))
            self.dec0_tables = (idp_tab.contiguous(), h_tab.contiguous())

    # ------------------------------------------------------------------ encoder
    def _enc_block(self, name: str, x: torch.Tensor, final_relu: bool = False) -> torch.Tensor:
        """EncoderBlock (vqvae_model.py:126-145): id(x) + 1x1(relu 3x3(relu 3x3(relu 3x3(relu x))))."""
        w = self.w
        n, H, W, C = x.shape
        x2 = x.view(-1, C)
        idp = ops.gemm(x2, w[name + ".id_path.weight"], w[name + ".id_path.bias"]) if (name + ".id_path.weight") in w else x2
        h = ops.conv2d(x, w[name + ".block.1.weight"], w[name + ".block.1.bias"], pad=(1, 1), relu_in=True, act=ACT_RELU)
        h = ops.conv2d(h, w[name + ".block.3.weight"], w[name + ".block.3.bias"], pad=(1, 1), act=ACT_RELU)
        h = ops.conv2d(h, w[name + ".block.5.weight"], w[name + ".block.5.bias"], pad=(1, 1), act=ACT_RELU)
        out = ops.gemm(h.view(-1, h.shape[-1]), w[name + ".block.7.weight"], w[name + ".block.7.bias"], residual=idp,
                       act=(ACT_RELU | ACT_POST) if final_relu else ACT_NONE)
        return out.view(n, H, W, -1)

    def _res_block(self, name: str, xr: torch.Tensor, post_relu: bool, res_relu: bool = False) -> torch.Tensor:
        """ResBlock (vqvae_model.py:111-124) on an input whose in-place ReLU is already applied
        (or applied on the fly with res_relu): xr + BN(1x1(relu(BN(3x3(xr)))))."""
        w = self.w
        n, H, W, C = xr.shape
        h = ops.conv2d(xr, w[name + ".c3.w"], w[name + ".c3.b"], pad=(1, 1), relu_in=res_relu, act=ACT_RELU)
        act = (ACT_RELU | ACT_POST) if post_relu else ACT_NONE
        if res_relu:
            act |= RES_RELU
        out = ops.gemm(h.view(-1, C), w[name + ".c1.w"], w[name + ".c1.b"], residual=xr.view(-1, C), act=act)
        return out.view(n, H, W, C)

    def _enc_block_tc(self, name: str, x: torch.Tensor, final_relu: bool = False, xr: Optional[torch.Tensor] = None) -> torch.Tensor:
        """EncoderBlock on the tensor cores: x fp32 NHWC in, fp32 NHWC out (feeds max-pool / VQ argmin).
        xr = split(relu(x)) when the producer already emitted it."""
        w, ws = self.w, self.ws
        n, H, W, C = x.shape
        if xr is None:
            xr = ops.split(x, relu=True)
        if (name + ".id_path.weight") in w:
            idp, _, _ = ops.gemm_tc(ops.split(x).view(2, -1, C), ws[name + ".id_path.weight"], w[name + ".id_path.bias"])
        else:
            idp = x.view(-1, C)
        _, h, _ = ops.conv2d_tc(xr, ws[name + ".block.1.weight"], w[name + ".block.1.bias"], pad=(1, 1), act=ACT_RELU, want=("split",))
        _, h, _ = ops.conv2d_tc(h, ws[name + ".block.3.weight"], w[name + ".block.3.bias"], pad=(1, 1), act=ACT_RELU, want=("split",))
        _, h, _ = ops.conv2d_tc(h, ws[name + ".block.5.weight"], w[name + ".block.5.bias"], pad=(1, 1), act=ACT_RELU, want=("split",))
        out, _, _ = ops.gemm_tc(h.view(2, -1, h.shape[-1]), ws[name + ".block.7.weight"], w[name + ".block.7.bias"], residual=idp,
                                act=(ACT_RELU | ACT_POST) if final_relu else ACT_NONE)
        return out.view(n, H, W, -1)

    def _res_block_tc(self, name: str, xr: torch.Tensor, xr_split: torch.Tensor, post_relu: bool, want):
        """ResBlock (vqvae_model.py:111-124) on the tensor cores; xr / xr_split = relu(x) as fp32 (the skip: the block's leading
        ReLU is in place) and as split operand.  Returns (fp32, split) of xr + BN(1x1(relu(BN(3x3(xr))))), ReLU-ed when the next
        consumer's in-place ReLU is folded in (`post_relu`)."""
        w, ws = self.w, self.ws
        n, H, W, C = xr.shape
        _, h, _ = ops.conv2d_tc(xr_split, ws[name + ".c3.w"], w[name + ".c3.b"], pad=(1, 1), act=ACT_RELU, want=("split",))
        out, out_split, _ = ops.gemm_tc(h.view(2, -1, C), ws[name + ".c1.w"], w[name + ".c1.b"], residual=xr.view(-1, C),
                                        act=(ACT_RELU | ACT_POST) if post_relu else ACT_NONE, want=want)
        return (out.view(n, H, W, C) if out is not None else None,
                out_split.view(2, n, H, W, C) if out_split is not None else None)

    def encode_features(self, x: torch.Tensor) -> torch.Tensor:
        """x [N,C,H,W] planar fp32 -> z_e NHWC [N,h,w,D] (vqvae_model.py:172-179 / :192-202)."""
        w = self.w
        x = x.contiguous()
        if self.backend == "tc" and self.down_ratio == 4:
            h = ops.conv2d_first(x, w["enc0_wt"], w["enc0_b"], cout=w["enc0_b"].numel(), kh=4, kw=4, stride=2, pad=1, act=ACT_RELU)
            xr, xr_split, _ = ops.conv2d_tc(ops.s2d_pad_split(h), self.ws["encoder.3.s2d"], w["encoder.3.bias"], pad=(0, 0),
                                            act=ACT_RELU, want=("f32", "split"))          # relu = encoder.4's in-place one
            xr, xr_split = self._res_block_tc("encoder.4", xr, xr_split, post_relu=True, want=("f32", "split"))
            return self._res_block_tc("encoder.5", xr, xr_split, post_relu=False, want=("f32",))[0]
        if self.backend == "tc":
            if "enc0_rows" in self.ws and x.shape[2] % 16 == 0 and x.shape[3] % 8 == 0:
                rows = ops.patch_rows_split(x, self.enc0_kw, self.enc0_pad)
                h, _, hr = ops.conv2d_tc(rows, self.ws["enc0_rows"], w["enc0_b"], pad=(self.enc0_pad, 0), want=("f32", "split_relu"))
                h = ops.maxpool2x2(self._enc_block_tc("encoder.1", h, xr=hr))
            else:
                h = ops.conv2d_first(x, w["enc0_wt"], w["enc0_b"], cout=w["enc0_b"].numel(), kh=7, kw=7, stride=1, pad=3)
                h = ops.maxpool2x2(self._enc_block_tc("encoder.1", h))
            h = ops.maxpool2x2(self._enc_block_tc("encoder.3", h))
            h = ops.maxpool2x2(self._enc_block_tc("encoder.5", h))
            return self._enc_block_tc("encoder.7", h, final_relu=True)
        if self.down_ratio == 8:
            h = ops.conv2d_first(x, w["enc0_wt"], w["enc0_b"], cout=w["enc0_b"].numel(), kh=7, kw=7, stride=1, pad=3)
            h = ops.maxpool2x2(self._enc_block("encoder.1", h))
            h = ops.maxpool2x2(self._enc_block("encoder.3", h))
            h = ops.maxpool2x2(self._enc_block("encoder.5", h))
            return self._enc_block("encoder.7", h, final_relu=True)
        h = ops.conv2d_first(x, w["enc0_wt"], w["enc0_b"], cout=w["enc0_b"].numel(), kh=4, kw=4, stride=2, pad=1, act=ACT_RELU)
        h = ops.conv2d(h, w["encoder.3.weight"], w["encoder.3.bias"], stride=2, pad=(1, 1), act=ACT_RELU)  # relu = ResBlock's in-place one
        h = self._res_block("encoder.4", h, post_relu=True)
        return self._res_block("encoder.5", h, post_relu=False)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """VectorQuantizedVAE.encode: [N,C,H,W] -> int64 [N,h,w]."""
        z = self.encode_features(x)
        n, h, w_, D = z.shape
        return ops.vq_argmin(z.view(-1, D), self.codebook).view(n, h, w_)

    # ------------------------------------------------------------------ decoder
    def _dec_block(self, name: str, x: torch.Tensor, up: bool) -> torch.Tensor:
        """DecoderBlock (vqvae_model.py:147-166) applied to nearest-x2-upsampled x when `up`.
        The two 1x1 convs commute with nearest upsampling, so they run at the stored (low)
        resolution and the 3x3 convs / the skip read them through the upsample."""
        w = self.w
        n, H, W, C = x.shape
        x2 = x.view(-1, C)
        has_id = (name + ".id_path.weight") in w
        idp = ops.gemm(x2, w[name + ".id_path.weight"], w[name + ".id_path.bias"]).view(n, H, W, -1) if has_id else x
        h = ops.gemm(x2, w[name + ".block.1.weight"], w[name + ".block.1.bias"], relu_a=True, act=ACT_RELU).view(n, H, W, -1)
        h = ops.conv2d(h, w[name + ".block.3.weight"], w[name + ".block.3.bias"], pad=(1, 1), in_up=up, act=ACT_RELU)
        h = ops.conv2d(h, w[name + ".block.5.weight"], w[name + ".block.5.bias"], pad=(1, 1), act=ACT_RELU)
        return ops.conv2d(h, w[name + ".block.7.weight"], w[name + ".block.7.bias"], pad=(1, 1), residual=idp,
                          res_mode=2 if up else 1)

    def _decode_into_tc(self, idx: torch.Tensor, out: torch.Tensor, out_img_stride: int) -> None:
        """f8 decoder on the tensor cores.  Activations travel between layers in the split operand format;
        a block's output is produced in exactly the forms its consumer reads: split(raw) for an id_path conv,
        fp32 for an identity skip (read through the nearest x2 upsample), split(relu) for the next 1x1."""
        w, ws = self.w, self.ws
        x_f32 = x_split = xr_split = None
        if self.dec0_tables is None:
            x_split = ops.embedding_split(idx, self.cb_split)        # [2,n,h,w,D]
            xr_split = ops.embedding_split(idx, self.cb_relu_split)
        blocks = [("decoder.0", False), ("decoder.2", True), ("decoder.4", True), ("decoder.6", True)]
        for bi, (name, up) in enumerate(blocks):
            if bi == 0 and self.dec0_tables is not None:
                n, H, W = idx.shape
                idp = ops.embedding(idx.reshape(-1), self.dec0_tables[0]).view(n, H, W, -1)
                h = ops.embedding_split(idx, self.dec0_tables[1])          # [2,n,h,w,hid]
            else:
                _, n, H, W, C = xr_split.shape
                if (name + ".id_path.weight") in w:
                    idp, _, _ = ops.gemm_tc(x_split.view(2, -1, C), ws[name + ".id_path.weight"], w[name + ".id_path.bias"])
                    idp = idp.view(n, H, W, -1)
                else:
                    idp = x_f32
                _, h, _ = ops.gemm_tc(xr_split.view(2, -1, C), ws[name + ".block.1.weight"], w[name + ".block.1.bias"], act=ACT_RELU,
                                      want=("split",))
            hid = h.shape[-1]
            h = h.view(2, n, H, W, hid)
            if up:
                h2 = torch.empty(2, n, 2 * H, 2 * W, hid, device=h.device, dtype=torch.float16)
                for py in (0, 1):
                    for px in (0, 1):
                        ops.conv2d_tc(h, ws[f"{name}.block.3.p{py}{px}"], w[name + ".block.3.bias"], pad=(1 - py, 1 - px), act=ACT_RELU,
                                      want=(), out_split=h2, out_hw=(H, W), scatter=(2, 2, py, px), full_hw=(2 * H, 2 * W))
            else:
                _, h2, _ = ops.conv2d_tc(h, ws[name + ".block.3.weight"], w[name + ".block.3.bias"], pad=(1, 1), act=ACT_RELU,
                                         want=("split",))
            last_passes = 1 if (bi + 1 == len(blocks) and self.decoder_precision == "budget") else 3
            _, h3, _ = ops.conv2d_tc(h2, ws[name + ".block.5.weight"], w[name + ".block.5.bias"], pad=(1, 1), act=ACT_RELU,
                                     want=("split",), passes=last_passes)
            if bi + 1 == len(blocks) and h3.shape[-1] % 64 == 0 and w[name + ".block.7.weight"].shape[0] == 256 \
                    and w["decoder.8.bias"].numel() <= 3:
                # last block: block.7 conv + skip, ReLU, 1x1 conv to pixels and tanh in ONE kernel (the [n,128,128,256] map stays on chip)
                ops.conv2d_tc_pixel_head(h3, ws[name + ".block.7.weight"], w[name + ".block.7.bias"], pad=(1, 1), residual=idp,
                                         res_mode=2 if up else 1, head_w=w["decoder.8.weight"], head_b=w["decoder.8.bias"],
                                         out=out, out_img_stride=out_img_stride, passes=last_passes)
                return
            if bi + 1 == len(blocks):
                want = ("f32",)
            elif (blocks[bi + 1][0] + ".id_path.weight") in w:
                want = ("split", "split_relu")
            else:
                want = ("f32", "split_relu")
            x_f32, x_split, xr_split = ops.conv2d_tc(h3, ws[name + ".block.7.weight"], w[name + ".block.7.bias"], pad=(1, 1),
                                                     residual=idp, res_mode=2 if up else 1, want=want, passes=last_passes)
        ops.conv1x1_tanh_nchw(x_f32, w["decoder.8.weight"], w["decoder.8.bias"], out, out_img_stride)

    def decode_into(self, idx: torch.Tensor, out: torch.Tensor, out_img_stride: int) -> None:
        """VectorQuantizedVAE.decode: idx int64 [N,h,w] -> tanh pixels written planar at
        out.data_ptr() + n*out_img_stride (elements), each image [C,H,W] contiguous."""
        if self.backend == "tc" and self.down_ratio == 8:
            return self._decode_into_tc(idx, out, out_img_stride)
        w = self.w
        n = idx.shape[0]
        if self.backend == "tc":
            # f4 decoder (vqvae_model.py:180-189): ResBlocks and ConvTranspose 256 -> 256 (four 2x2 sub-pixel phases) on tcgen05
            H, W = idx.shape[1], idx.shape[2]
            zr = ops.embedding(idx.reshape(-1), self.cb_relu).view(n, H, W, self.D)       # relu(z): decoder.0's in-place ReLU
            zr_split = ops.embedding_split(idx, self.cb_relu_split)
            h, h_split = self._res_block_tc("decoder.0", zr, zr_split, post_relu=True, want=("f32", "split"))
            _, h_split = self._res_block_tc("decoder.1", h, h_split, post_relu=True, want=("split",))   # + decoder.2 ReLU
            C = self.D
            up1 = torch.empty(n, 2 * H, 2 * W, C, device=idx.device, dtype=torch.float32)
            for py in (0, 1):
                for px in (0, 1):
                    ops.conv2d_tc(h_split, self.ws[f"decoder.3.p{py}{px}"], w["decoder.3.b"], pad=(1 - py, 1 - px), act=ACT_RELU,
                                  want=(), out=up1, out_hw=(H, W), scatter=(2, 2, py, px), full_hw=(2 * H, 2 * W))
            for py in (0, 1):
                for px in (0, 1):
                    ops.conv2d(up1, w[f"decoder.6.p{py}{px}"], w["decoder.6.b"], pad=(1 - py, 1 - px), act=ACT_TANH, out=out,
                               out_hw=(2 * H, 2 * W), scatter=(2, 2, py, px), full_hw=(4 * H, 4 * W), out_img_stride=out_img_stride)
            return
        z = ops.embedding(idx.reshape(-1), self.codebook).view(n, idx.shape[1], idx.shape[2], self.D)
        if self.down_ratio == 8:
            h = self._dec_block("decoder.0", z, up=False)
            h = self._dec_block("decoder.2", h, up=True)
            h = self._dec_block("decoder.4", h, up=True)
            h = self._dec_block("decoder.6", h, up=True)
            ops.conv1x1_tanh_nchw(h, w["decoder.8.weight"], w["decoder.8.bias"], out, out_img_stride)
            return
        h = self._res_block("decoder.0", z, post_relu=True, res_relu=True)
        h = self._res_block("decoder.1", h, post_relu=True)  # + decoder.2 ReLU
        _, H, W, C = h.shape
        up1 = torch.empty(n, 2 * H, 2 * W, C, device=h.device, dtype=torch.float32)
        for py in (0, 1):
            for px in (0, 1):
                ops.conv2d(h, w[f"decoder.3.p{py}{px}"], w["decoder.3.b"], pad=(1 - py, 1 - px), act=ACT_RELU, out=up1,
                           out_hw=(H, W), scatter=(2, 2, py, px), full_hw=(2 * H, 2 * W))
        for py in (0, 1):
            for px in (0, 1):
                ops.conv2d(up1, w[f"decoder.6.p{py}{px}"], w["decoder.6.b"], pad=(1 - py, 1 - px), act=ACT_TANH, out=out,
                           out_hw=(2 * H, 2 * W), scatter=(2, 2, py, px), full_hw=(4 * H, 4 * W), out_img_stride=out_img_stride)

    def decode(self, idx: torch.Tensor) -> torch.Tensor:
        n, h, w_ = idx.shape
        R = h * self.down_ratio
        out = torch.empty(n, self.in_ch, R, R, device=idx.device, dtype=torch.float32)
        self.decode_into(idx.contiguous(), out, self.in_ch * R * R)
        return out


class SamplerEngine:
    """Incremental greedy sampler for one device.  `sd` is MAGE.state_dict() (CUDA fp32)."""

    def __init__(self, sd: Dict[str, torch.Tensor], frames_length: int, randomness: bool, padding_idx: int = 0,
                 temporal_attn: Optional[str] = None, use_cuda_graph: Optional[bool] = None, backend: Optional[str] = None,
                 use_cids: bool = True, ma_ln: bool = False):
        self.device = sd["visual_token_embedding.weight"].device
        assert self.device.type == "cuda", "SamplerEngine needs CUDA tensors (no CPU path)"
        self.backend = backend or default_backend()
        ops.flag(self.device)
        self.L = frames_length
        self.randomness = randomness
        self.padding_idx = padding_idx
        self.use_cids = use_cids      # False: MAGE+ continuous-latent branch (generate_continuous)
        self.ma_ln = ma_ln            # TransformerBlock line 93 (ln_q / ln_kv) instead of the shipped line 92
        self.temporal_attn = temporal_attn or os.environ.get("MAGE_TEMPORAL_ATTN", "tma")
        if use_cuda_graph is None:
            use_cuda_graph = os.environ.get("MAGE_CUDA_GRAPH", "1") != "0"
        self.use_cuda_graph = use_cuda_graph
        self.vq = VQVAEEngine({k[len("first_stage_model."):]: v for k, v in sd.items() if k.startswith("first_stage_model.")},
                              backend=self.backend) if use_cids else None
        g = lambda k: sd[k].contiguous()
        self.sd = sd
        self.E = g("visual_token_embedding.weight")      # [K, C] table (use_cids) or Linear weight [C, embed_dim] (MAGE+)
        self.C = sd["conv.0.weight"].shape[0]
        self.R = sd["H_positional_embedding"].shape[1]
        self.Wc = _pack_conv(sd["conv.0.weight"])
        self.posHW = (sd["H_positional_embedding"] + sd["W_positional_embedding"]).reshape(self.R * self.R, self.C).contiguous()
        p = "generate_model."
        Tp = sd[p + "T_positional_embedding"].reshape(-1, self.C)
        assert Tp.shape[0] >= frames_length, "checkpoint has fewer temporal positions than frames_length"
        self.bias_in_T = (sd[p + "in_linear.bias"].unsqueeze(0) + Tp).contiguous()      # row p: in_linear bias + T_pos[p]
        self.bias_ctx0 = (sd[p + "context_linear.bias"] + Tp[0]).contiguous()
        self.n_blocks = 0
        while (p + f"blocks.{self.n_blocks}.ln_1.weight") in sd:
            self.n_blocks += 1
        self.n_text_layers = 0
        while f"text_encoder.transformer.layers.{self.n_text_layers}.linear1.weight" in sd:
            self.n_text_layers += 1
        self.n_ma_layers = 0
        while f"ma_encoder.blocks.{self.n_ma_layers}.attn.in_proj_weight" in sd:
            self.n_ma_layers += 1
        if randomness:
            self.Wd2 = _pack_conv(sd["conv_d2.weight"])
            self.adain_w = {f"{br}.{i}": (_pack_conv(sd[f"adain.{br}.{i}.weight"]), g(f"adain.{br}.{i}.bias"))
                            for br in ("conv_mu", "conv_var") for i in (0, 1)}
        self.scale = 1.0 / math.sqrt(32.0)
        self.n_head = self.C // 32
        self._graphs = {}
        self.tok_table = None
        self.kernels_per_generate = None
        # The VQ-VAE decode of frame j depends only on that frame's tokens and nothing downstream depends on it, so it runs on a
        # side stream next to transformer step j+1: the tails / launch gaps of one kernel sequence are filled by the other.
        self.overlap_decode = os.environ.get("MAGE_OVERLAP_DECODE", "0") != "0"   # measured at B=64: no gain (154.2 vs 154.4 ms), off there
        self._side = None
        self._copy = None
        self._chunk_streams = []
        self._pool = None
        self.max_graphs = max(1, int(os.environ.get("MAGE_MAX_GRAPHS", "8")))
        # frames are decoded in groups of `decode_group` steps (one VQ-VAE decoder pass over group*B images): the low-resolution
        # decoder layers are too small to fill 148 SMs at B images, and a frame's pixels are not needed before the call returns.
        # 0 = chosen from the batch size (`_plan`); likewise the number of chunk streams.
        self.decode_group = max(0, int(os.environ.get("MAGE_DECODE_GROUP", "0")))
        self.n_streams = max(0, int(os.environ.get("MAGE_STREAMS", "0")))
        # small-batch schedule: decoder beside the steps on a share of the SMs (`_side_plan`)
        self.n_sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.side_sms = int(os.environ.get("MAGE_SIDE_SMS", "0"))
        self.side_frames = int(os.environ.get("MAGE_SIDE_FRAMES", "0"))
        self.side_group = int(os.environ.get("MAGE_SIDE_GROUP", "0"))
        self.fused_axial = self.fused_ln = self.fused_ln_taps = False   # tensor-core back end only (set below)
        self._post = None           # packed video posterior (MAGE.forward only), built on first use
        if self.backend == "tc":
            # split (fp16 hi/lo) copies of the per-step tensor-core operands
            ws = {"Wc": ops.split(self.Wc)}
            if use_cids:
                ws["E"] = ops.split(self.E)
            for k in ("in_linear.weight", "out.weight") if use_cids else ("in_linear.weight",):
                ws[p + k] = ops.split(g(p + k))
            # H / W blocks: QKV projection and the 16x16 axial attention run as ONE kernel (mage_qkv_axial_attn_tc) on a
            # head-permuted copy of the packed in-projection; MAGE_FUSED_AXIAL=0 keeps the two-kernel form (tests compare them)
            self.fused_axial = os.environ.get("MAGE_FUSED_AXIAL", "1") != "0" and self.R == 16 and (self.C // 32) % 2 == 0
            # LayerNorm fused into the kernel that produces its input; every form gives the same bits (tests compare them).
            #   fused_ln_taps (MAGE_FUSED_LN != 0, default): mage_token_taps_ln_f32 applies the first block's ln_1 to the row its
            #     warp has just finished -- one launch less per step, no synchronisation involved.
            #   fused_ln (MAGE_FUSED_LN=all, OFF by default): mage_gemm_tc_ln -- the out-projection applies the block's ln_2, c_proj
            #     the next block's ln_1 (11 more launches less per step).  A row spans several CTAs' tiles, so the CTA that
            #     completes a 128-row block normalises it; that completion hand-off costs more than the launch it saves: measured
            #     7 % SLOWER at 64 prompts, 17 % slower at 8 (profiles/r02aa_fused_layernorm_ab.txt).  Kept as a tested option.
            mode = os.environ.get("MAGE_FUSED_LN", "taps")
            self.fused_ln_taps = mode != "0" and self.C == 512
            self.fused_ln = mode in ("all", "2") and self.C == 512
            self.axial_bias = {}
            for i in range(self.n_blocks):
                bp = p + f"blocks.{i}."
                for k in ("attn.in_proj_weight", "attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight"):
                    ws[bp + k] = ops.split(g(bp + k))
                if i % 3 != 0 and self.R == 16 and (self.C // 32) % 2 == 0:
                    wp, bp_ = ops.permute_qkv_for_axial(g(bp + "attn.in_proj_weight"), g(bp + "attn.in_proj_bias"), self.C // 32)
                    ws[bp + "attn.in_proj_weight.axial"] = ops.split(wp)
                    self.axial_bias[i] = bp_
            # prelude operands that see the whole batch (M = B*256 rows): motion-anchor block, context_linear, AdaIN convs
            ws[p + "context_linear.weight"] = ops.split(g(p + "context_linear.weight"))
            for i in range(self.n_ma_layers):
                mp = f"ma_encoder.blocks.{i}."
                ws[mp + "q_weight"] = ops.split(sd[mp + "attn.in_proj_weight"][:self.C].contiguous())
                for k in ("attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight"):
                    ws[mp + k] = ops.split(g(mp + k))
            # in_linear(conv3x3(E[tok]) + pos) as lookups: the conv input is one of K embedding rows per pixel, so
            # table[tap][code] = W_in . Wc[:,:,tap] . E[code] (formed in fp64, stored fp32) replaces a 3x3 512->512 convolution and a
            # 512x512 GEMM per step (86 GFLOP at B=64) by nine 2 KB gathers per pixel; fixed summation order (taps 0..8).
            if use_cids and os.environ.get("MAGE_TOKEN_TABLE", "1") != "0":
                Win = sd[p + "in_linear.weight"].double()
                Wc = sd["conv.0.weight"].double()                                # [C, C, 3, 3]
                comp = torch.einsum("oc,cikl->klio", Win, Wc)                      # [3, 3, C_in, C_out]
                table = torch.einsum("ei,klio->kleo", self.E.double(), comp)       # [3, 3, K, C_out]
                self.tok_table = table.reshape(9, self.E.shape[0], self.C).float().contiguous()
                self.tok_posW = (self.posHW.double() @ Win.t()).float().contiguous()  # [R*R, C]
            else:
                self.tok_table = None
            if randomness:
                ws["Wd2"] = ops.split(self.Wd2)
                for k, (wk, _) in self.adain_w.items():
                    ws["adain." + k] = ops.split(wk)
            self.ws = ws

    # ------------------------------------------------------------------ prelude
    def _token_features(self, tok: torch.Tensor, B: int) -> torch.Tensor:
        """f(tok) = conv3x3(E[tok]) + Hpos + Wpos  -> [B*R*R, C]   (mage_model.py:644-649,674-676)."""
        emb = ops.embedding(tok.reshape(-1), self.E).view(B, self.R, self.R, self.C)
        return ops.conv2d(emb, self.Wc, None, pad=(1, 1), residual=self.posHW, res_mode=3).view(-1, self.C)

    def _text_encoder(self, text: torch.Tensor):
        """TransformerTextEncoder.forward (mage_model.py:223-250) -> ([B*T, C], key_len)."""
        sd, p = self.sd, "text_encoder."
        B, T = text.shape
        x, key_len = ops.text_embed(text, sd[p + "token_embedding.weight"], sd[p + "positions.weight"],
                                    sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], self.padding_idx, 1e-8)
        x = x.view(B * T, -1)
        W = x.shape[1]
        for i in range(self.n_text_layers):
            lp = p + f"transformer.layers.{i}"
            qkv = ops.gemm(x, sd[lp + ".self_attn.in_proj_weight"], sd[lp + ".self_attn.in_proj_bias"])
            a = torch.empty(B * T, W, device=x.device, dtype=torch.float32)
            ops.mha(qkv, qkv[:, W:], qkv[:, 2 * W:], a, n_outer=B, n_inner=1, n_head=W // 32, Sq=T, Sk=T,
                    q_strides=(T * 3 * W, 0, 3 * W), k_strides=(T * 3 * W, 0, 3 * W), v_strides=(T * 3 * W, 0, 3 * W),
                    o_strides=(T * W, 0, W), key_len=key_len, scale=self.scale)
            y = ops.gemm(a, sd[lp + ".self_attn.out_proj.weight"], sd[lp + ".self_attn.out_proj.bias"], residual=x)
            x = ops.layernorm(y, sd[lp + ".norm1.weight"], sd[lp + ".norm1.bias"])
            h = ops.gemm(x, sd[lp + ".linear1.weight"], sd[lp + ".linear1.bias"], act=ACT_GELU)
            y = ops.gemm(h, sd[lp + ".linear2.weight"], sd[lp + ".linear2.bias"], residual=x)
            x = ops.layernorm(y, sd[lp + ".norm2.weight"], sd[lp + ".norm2.bias"])
        x = ops.layernorm(x, sd[p + "ln_text_final.weight"], sd[p + "ln_text_final.bias"])
        return ops.gemm(x, sd[p + "text_projection.weight"], sd[p + "text_projection.bias"]), key_len

    def _ma_encoder(self, q: torch.Tensor, temb: torch.Tensor, B: int, T: int) -> torch.Tensor:
        """MAEncoder (mage_model.py:114-117, TransformerBlock line 92: no LN on q/kv, no key mask).
        q [B*HW, C] batch-major, temb [B*T, C]."""
        sd, C = self.sd, self.C
        HW = self.R * self.R
        x = q
        for i in range(self.n_ma_layers):
            p = f"ma_encoder.blocks.{i}"
            Win, bin_ = sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"]
            qp = ops.gemm(x, Win[:C], bin_[:C])
            kv = ops.gemm(temb, Win[C:], bin_[C:])  # [B*T, 2C]
            a = torch.empty(B * HW, C, device=x.device, dtype=torch.float32)
            ops.mha(qp, kv, kv[:, C:], a, n_outer=B, n_inner=1, n_head=self.n_head, Sq=HW, Sk=T,
                    q_strides=(HW * C, 0, C), k_strides=(T * 2 * C, 0, 2 * C), v_strides=(T * 2 * C, 0, 2 * C),
                    o_strides=(HW * C, 0, C), key_len=None, scale=self.scale)
            x = ops.gemm(a, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"], residual=x)
            u = ops.layernorm(x, sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"])
            h = ops.gemm(u, sd[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"], act=ACT_QUICKGELU)
            x = ops.gemm(h, sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"], residual=x)
        return x

    def _adain(self, anchor: torch.Tensor, noise_nchw: torch.Tensor, B: int) -> torch.Tensor:
        """conv_d2 + ADAIN2D (mage_model.py:660-664, 309-314); anchor NHWC [B,R,R,C]."""
        y = ops.conv2d(ops.nchw_to_nhwc(noise_nchw), self.Wd2, None, pad=(1, 1))
        mods = []
        for br in ("conv_mu", "conv_var"):
            w0, b0 = self.adain_w[br + ".0"]
            w1, b1 = self.adain_w[br + ".1"]
            mods.append(ops.conv2d(ops.conv2d(y, w0, b0, pad=(1, 1)), w1, b1, pad=(1, 1)))
        return ops.adain(anchor, mods[0], mods[1], 1e-5)

    def _ma_encoder_tc(self, q: torch.Tensor, q_split: torch.Tensor, temb: torch.Tensor, B: int, T: int) -> torch.Tensor:
        """MAEncoder on the tensor cores (same math as _ma_encoder): q fp32 [B*HW, C] and its split copy."""
        sd, ws, C = self.sd, self.ws, self.C
        HW = self.R * self.R
        M = B * HW
        x, x_split = q, q_split
        for i in range(self.n_ma_layers):
            p = f"ma_encoder.blocks.{i}"
            Win, bin_ = sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"]
            if self.ma_ln:   # TransformerBlock line 93 (MAGE+): attention reads ln_q(q), ln_kv(kv); the residual stays q
                x_split = ops.layernorm(x, sd[p + ".ln_q.weight"], sd[p + ".ln_q.bias"],
                                        out_split=torch.empty(2, M, C, device=x.device, dtype=torch.float16))
                temb_i = ops.layernorm(temb, sd[p + ".ln_kv.weight"], sd[p + ".ln_kv.bias"])
            else:
                temb_i = temb
                if x_split is None:
                    x_split = ops.split(x)
            qp, _, _ = ops.gemm_tc(x_split, ws[p + ".q_weight"], bin_[:C].contiguous())
            kv = ops.gemm(temb_i, Win[C:], bin_[C:])  # [B*T, 2C]: a few hundred rows, fp32 FFMA kernel
            a = torch.empty(2, M, C, device=x.device, dtype=torch.float16)
            ops.mha(qp, kv, kv[:, C:], None, n_outer=B, n_inner=1, n_head=self.n_head, Sq=HW, Sk=T,
                    q_strides=(HW * C, 0, C), k_strides=(T * 2 * C, 0, 2 * C), v_strides=(T * 2 * C, 0, 2 * C),
                    o_strides=(HW * C, 0, C), key_len=None, scale=self.scale, out_split=a)
            x, _, _ = ops.gemm_tc(a, ws[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"], residual=x)
            ops.layernorm(x, sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], out_split=a)
            _, h, _ = ops.gemm_tc(a, ws[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"], act=ACT_QUICKGELU, want=("split",))
            ops.gemm_tc(h, ws[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"], residual=x, out=x)
            x_split = None
        return x

    def _adain_tc(self, anchor: torch.Tensor, noise_nchw: torch.Tensor, B: int) -> torch.Tensor:
        """conv_d2 + ADAIN2D with the five 3x3 convs as tcgen05 implicit GEMMs (same math as _adain)."""
        ws = self.ws
        _, y, _ = ops.conv2d_tc(ops.split(ops.nchw_to_nhwc(noise_nchw)), ws["Wd2"], None, pad=(1, 1), want=("split",))
        mods = []
        for br in ("conv_mu", "conv_var"):
            _, h, _ = ops.conv2d_tc(y, ws[f"adain.{br}.0"], self.adain_w[br + ".0"][1], pad=(1, 1), want=("split",))
            m, _, _ = ops.conv2d_tc(h, ws[f"adain.{br}.1"], self.adain_w[br + ".1"][1], pad=(1, 1))
            mods.append(m)
        return ops.adain(anchor, mods[0], mods[1], 1e-5)

    # ------------------------------------------------------------------ decoder step
    def _block_step(self, i: int, x: torch.Tensor, pos: int, B: int, caches) -> torch.Tensor:
        """AxialAttentionBlock (mage_model.py:35-53) on one temporal position; x [B*R*R, C] updated in place."""
        sd, C, R = self.sd, self.C, self.R
        p = f"generate_model.blocks.{i}"
        M = x.shape[0]
        u = ops.layernorm(x, sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"])
        qkv = ops.gemm(u, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"])
        a = u  # reuse the LN buffer for the attention output
        kind = i % 3
        if kind == 0:
            kc, vc = caches[i]
            if self.temporal_attn == "tma":
                ops.temporal_attn_step(qkv, kc, vc, a, pos, self.scale)
            else:
                ops.kv_append(qkv, kc, vc, pos)
                Lmax = kc.shape[1]
                ops.mha(qkv, kc, vc, a, n_outer=M, n_inner=1, n_head=self.n_head, Sq=1, Sk=pos + 1,
                        q_strides=(3 * C, 0, 0), k_strides=(Lmax * C, 0, C), v_strides=(Lmax * C, 0, C),
                        o_strides=(C, 0, 0), key_len=None, scale=self.scale)
        else:
            # rows are (b, h, w); H-block: sequences run over h (stride R rows) for fixed (b, w); W-block over w
            if R == 16:
                ops.axial_attn(qkv, a, B=B, R=R, n_head=self.n_head, axis=kind, scale=self.scale)
            else:
                inner, seq = (1, R) if kind == 1 else (R, 1)
                ops.mha(qkv, qkv[:, C:], qkv[:, 2 * C:], a, n_outer=B, n_inner=R, n_head=self.n_head, Sq=R, Sk=R,
                        q_strides=(R * R * 3 * C, inner * 3 * C, seq * 3 * C), k_strides=(R * R * 3 * C, inner * 3 * C, seq * 3 * C),
                        v_strides=(R * R * 3 * C, inner * 3 * C, seq * 3 * C), o_strides=(R * R * C, inner * C, seq * C),
                        key_len=None, scale=self.scale)
        ops.gemm(a, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"], residual=x, out=x)
        ops.layernorm(x, sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], out=u)
        h = ops.gemm(u, sd[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"], act=ACT_QUICKGELU)
        ops.gemm(h, sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"], residual=x, out=x)
        return x

    # ------------------------------------------------------------------ decoder step on the tensor cores
    def _token_features_tc(self, tok: torch.Tensor, B: int, want_f32: bool = False):
        """split(f(tok)) [2, B*R*R, C]: gather from the pre-split embedding table, 3x3 conv as a tcgen05 implicit
        GEMM with the H/W positional map added in the epilogue (mage_model.py:674-676, 682)."""
        emb = ops.embedding_split(tok.view(B, self.R, self.R), self.ws["E"])
        o, f, _ = ops.conv2d_tc(emb, self.ws["Wc"], None, pad=(1, 1), residual=self.posHW, res_mode=3,
                                want=("f32", "split") if want_f32 else ("split",))
        if want_f32:
            return o.view(-1, self.C), f.view(2, -1, self.C)
        return f.view(2, -1, self.C)

    def _block_step_tc(self, i: int, x: torch.Tensor, pos: int, B: int, caches, bufs, last: bool, ln1_done: bool = False):
        """AxialAttentionBlock (mage_model.py:35-53) for one temporal position, dense contractions on tcgen05.
        x [B*R*R, C] fp32 residual stream (updated in place); LayerNorm / attention emit split operands.
        With `fused_ln` the two GEMMs that write x also emit split(LayerNorm(x)) for the GEMM that follows (ln_2 of this block after
        the out-projection, ln_1 of block i+1 after c_proj); `ln1_done`: bufs["u"] already holds split(ln_1(x)) on entry."""
        sd, ws, C, R = self.sd, self.ws, self.C, self.R
        p = f"generate_model.blocks.{i}"
        M = x.shape[0]
        u, a, h, qkv = bufs["u"], bufs["a"], bufs["h"], bufs["qkv"]
        fused_ln = self.fused_ln
        if not ln1_done:
            ops.layernorm(x, sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"], out_split=u)
        kind = i % 3
        if kind != 0 and self.fused_axial:
            ops.qkv_axial_attn_tc(u, ws[p + ".attn.in_proj_weight.axial"], self.axial_bias[i], a, n_img=B, R=R, n_head=self.n_head,
                                  axis=kind, scale=self.scale)
        else:
            ops.gemm_tc(u, ws[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"], out=qkv)
            if kind == 0:
                kc, vc = caches[i]
                if self.temporal_attn == "tma":
                    ops.temporal_attn_step(qkv, kc, vc, None, pos, self.scale, out_split=a)
                else:
                    ops.kv_append(qkv, kc, vc, pos)
                    Lmax = kc.shape[1]
                    ops.mha(qkv, kc, vc, None, n_outer=M, n_inner=1, n_head=self.n_head, Sq=1, Sk=pos + 1,
                            q_strides=(3 * C, 0, 0), k_strides=(Lmax * C, 0, C), v_strides=(Lmax * C, 0, C),
                            o_strides=(C, 0, 0), key_len=None, scale=self.scale, out_split=a)
            elif R == 16:
                ops.axial_attn(qkv, None, B=B, R=R, n_head=self.n_head, axis=kind, scale=self.scale, out_split=a)
            else:
                inner, seq = (1, R) if kind == 1 else (R, 1)
                ops.mha(qkv, qkv[:, C:], qkv[:, 2 * C:], None, n_outer=B, n_inner=R, n_head=self.n_head, Sq=R, Sk=R,
                        q_strides=(R * R * 3 * C, inner * 3 * C, seq * 3 * C), k_strides=(R * R * 3 * C, inner * 3 * C, seq * 3 * C),
                        v_strides=(R * R * 3 * C, inner * 3 * C, seq * 3 * C), o_strides=(R * R * C, inner * C, seq * C),
                        key_len=None, scale=self.scale, out_split=a)
        if fused_ln:
            ops.gemm_tc_ln(a, ws[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"], residual=x, out=x,
                           gamma=sd[p + ".ln_2.weight"], beta=sd[p + ".ln_2.bias"], ln_out=u, counters=bufs["ln_count"])
        else:
            ops.gemm_tc(a, ws[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"], residual=x, out=x)
            ops.layernorm(x, sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], out_split=u)
        ops.gemm_tc(u, ws[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"], act=ACT_QUICKGELU, want=(), out_split=h)
        if fused_ln and not last and i + 1 < self.n_blocks:
            q = f"generate_model.blocks.{i + 1}"
            ops.gemm_tc_ln(h, ws[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"], residual=x, out=x,
                           gamma=sd[q + ".ln_1.weight"], beta=sd[q + ".ln_1.bias"], ln_out=u, counters=bufs["ln_count"])
        else:
            ops.gemm_tc(h, ws[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"], residual=x, out=x,
                        out_split=u if last else None)  # the head reads split(x)
        return x

    # ------------------------------------------------------------------ whole path
    def _prelude(self, images0: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor], noise: Optional[torch.Tensor],
                 tok0_out: torch.Tensor, trace: Optional[dict]) -> dict:
        """Everything that runs once per prompt (SURVEY.md App. D `prelude`) for one chunk of the batch, up to and including the
        motion anchor's pass through the six blocks as temporal position 0 (fills the K/V caches).  Returns the chunk's decode
        state: residual stream x, K/V caches, scratch buffers."""
        sd, C, R, L = self.sd, self.C, self.R, self.L
        B, T = text.shape
        M = B * R * R
        p = "generate_model."
        tc = self.backend == "tc"
        anchor = self._motion_anchor(images0, text, speed, noise, tok0_out, trace)
        caches = {i: (torch.empty(M, L, C, device=self.device, dtype=torch.float32),
                      torch.empty(M, L, C, device=self.device, dtype=torch.float32))
                  for i in range(self.n_blocks) if i % 3 == 0}
        bufs = None
        if tc:
            x, _, _ = ops.gemm_tc(ops.split(anchor.view(M, C)), self.ws[p + "context_linear.weight"], self.bias_ctx0)
            bufs = {"u": torch.empty(2, M, C, device=self.device, dtype=torch.float16),
                    "a": torch.empty(2, M, C, device=self.device, dtype=torch.float16),
                    "h": torch.empty(2, M, 4 * C, device=self.device, dtype=torch.float16),
                    "qkv": torch.empty(M, 3 * C, device=self.device, dtype=torch.float32),
                    # arrival counters of mage_gemm_tc_ln, one per 128-row block; every launch leaves them zero
                    "ln_count": torch.zeros((M + 127) // 128, device=self.device, dtype=torch.int32)}
        else:
            x = ops.gemm(anchor.view(M, C), sd[p + "context_linear.weight"], self.bias_ctx0)
        for i in range(self.n_blocks):
            x = (self._block_step_tc(i, x, 0, B, caches, bufs, False, ln1_done=self.fused_ln and i > 0) if tc
                 else self._block_step(i, x, 0, B, caches))
        logits = torch.empty(M, sd[p + "out.weight"].shape[0], device=self.device, dtype=torch.float32)
        return dict(B=B, x=x, caches=caches, bufs=bufs, logits=logits, tok=tok0_out)

    def _motion_anchor(self, images0: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor], noise: Optional[torch.Tensor],
                       tok0_out: torch.Tensor, trace: Optional[dict]) -> torch.Tensor:
        """mage_model.py:642-668 (= :577-614 of the training forward): VQ tokens of frame 0 into `tok0_out`, their features, the text
        encoder, the motion-anchor cross-attention, AdaIN with `noise` [B,64,R,R] (the test-time draw, or the posterior sample of
        MAGE.forward), the speed embedding.  Returns the anchor fp32 [B,R,R,C].  images0=None: `tok0_out` already holds frame 0's
        tokens (MAGE.forward encodes all frames at once)."""
        sd, C, R = self.sd, self.C, self.R
        B, T = text.shape
        M = B * R * R
        if images0 is not None:
            z = self.vq.encode_features(images0)
            ops.vq_argmin(z.view(M, -1), self.vq.codebook, out=tok0_out.view(-1))
        tc = self.backend == "tc"
        temb, _ = self._text_encoder(text)
        if tc:
            f0, f0_split = self._token_features_tc(tok0_out, B, want_f32=True)
            anchor = self._ma_encoder_tc(f0, f0_split, temb, B, T).view(B, R, R, C)
        else:
            f0 = self._token_features(tok0_out, B)
            anchor = self._ma_encoder(f0, temb, B, T).view(B, R, R, C)
        if trace is not None:
            trace["text_emb"], trace["first_img"], trace["anchor_ma"] = temb.view(B, T, C).clone(), f0.view(B, R * R, C).clone(), anchor.clone()
        if noise is not None:
            anchor = self._adain_tc(anchor, noise, B) if tc else self._adain(anchor, noise, B)
        if speed is not None:
            ops.add_scaled_vec(anchor, speed, sd["speed_embedding"].view(-1))
        if trace is not None:
            trace["anchor"] = anchor.clone()
        return anchor

    def _decode_step(self, st: dict, j: int, tok_out: torch.Tensor, trace: Optional[dict]) -> None:
        """Temporal position j+1 of one chunk: features of the previous frame's tokens -> six blocks -> head -> greedy tokens of
        frame j+1 into `tok_out` (int64 [B*R*R], a slice of the step-major token buffer)."""
        sd, R = self.sd, self.R
        p = "generate_model."
        B, x, caches, bufs, logits, tok = st["B"], st["x"], st["caches"], st["bufs"], st["logits"], st["tok"]
        if self.backend == "tc":
            ws = self.ws
            ln1 = False
            if self.tok_table is not None and self.fused_ln_taps:
                ops.token_taps_ln(tok.view(B, R, R), self.tok_table, self.tok_posW, self.bias_in_T[j + 1], x,
                                  sd[p + "blocks.0.ln_1.weight"], sd[p + "blocks.0.ln_1.bias"], bufs["u"])
                ln1 = True
            elif self.tok_table is not None:
                ops.token_taps(tok.view(B, R, R), self.tok_table, self.tok_posW, self.bias_in_T[j + 1], x)
            else:
                f = self._token_features_tc(tok, B)
                ops.gemm_tc(f, ws[p + "in_linear.weight"], self.bias_in_T[j + 1], out=x)
            for i in range(self.n_blocks):
                x = self._block_step_tc(i, x, j + 1, B, caches, bufs, i + 1 == self.n_blocks,
                                        ln1_done=ln1 if i == 0 else self.fused_ln)
            ops.gemm_tc(bufs["u"], ws[p + "out.weight"], sd[p + "out.bias"], out=logits)
        else:
            f = self._token_features(tok, B)
            x = ops.gemm(f, sd[p + "in_linear.weight"], self.bias_in_T[j + 1])
            for i in range(self.n_blocks):
                x = self._block_step(i, x, j + 1, B, caches)
            ops.gemm(x, sd[p + "out.weight"], sd[p + "out.bias"], out=logits)
        st["x"] = x
        st["tok"] = ops.argmax_rows(logits, out=tok_out)
        if trace is not None:
            if trace.get("ce_rows") is not None:
                # stage-2 objective (mage_model.py:619): cross-entropy of this position's logits against the GIVEN next frame
                ops.cross_entropy_rows(logits, trace["force_tokens"][:, j].reshape(-1).contiguous(), trace["ce_rows"][j])
            if trace.get("keep_logits", True):
                trace.setdefault("logits", []).append(logits.clone())
            if trace.get("force_tokens") is not None:
                # teacher forcing (parity diagnostics): this step's prediction is recorded, but the NEXT step is fed the
                # given token map -- a near-tie flip then cannot cascade into later frames
                st["tok"] = trace["force_tokens"][:, j].reshape(-1).contiguous()

    def _plan(self, B: int):
        """(number of chunk streams, decode group) for a batch of B prompts.  The VQ-VAE decoder -- which nothing downstream depends
        on -- runs over groups of finished frames: 4 frames at large batches, more at small ones so that a decoder pass still
        covers >= 128 images (its low-resolution layers cannot fill 148 SMs otherwise).  Chunk streams (the batch cut into S
        chunks whose launch sequences run concurrently, `MAGE_STREAMS`) are OFF by default: measured slower at every batch size
        (profiles/r02b_small_batch_schedule_sweep.txt -- a decode step's kernels are persistent and each takes the whole machine,
        so two chains only interleave and every chunk pays the per-kernel latency again)."""
        S = self.n_streams if self.n_streams > 0 else 1
        S = max(1, min(S, B))
        G = self.decode_group
        if G <= 0:
            G = 4 if B >= 32 else max(4, min(self.L - 1, 128 // max(B, 1)))
        return S, G

    def _generate_impl(self, images0: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor],
                       noise: Optional[torch.Tensor], video: torch.Tensor, tokens: torch.Tensor, tok0_out: torch.Tensor,
                       trace: Optional[dict] = None, host_video: Optional[torch.Tensor] = None) -> None:
        """images0 [B,Cimg,Himg,Wimg]; text i64 [B,T]; video FRAME-MAJOR [L,B,Cimg,Himg,Wimg] (frame 0 = images0, each generated
        frame of the whole batch is one contiguous block); host_video: optional pinned host tensor of the same shape that
        receives every frame over a copy stream as soon as it is decoded (the D2H overlaps the following steps);
        tokens i64 [L-1, B, R*R] (step-major); tok0_out i64 [B, R*R].

        Schedule: the batch is cut into S contiguous chunks (`_plan`); chunk c runs its prelude and its L-1 decode steps on its
        own stream, the VQ-VAE decoder runs over groups of G finished frames of the WHOLE batch on a decode stream, frames leave
        for the host on a copy stream.  Rows are independent and every kernel is deterministic, so the schedule cannot change a
        bit of the result (tests/test_gpu_parity.py::test_optional_schedules_are_bit_exact)."""
        R, L = self.R, self.L
        B = text.shape[0]
        main = torch.cuda.current_stream()
        S, G = self._plan(B) if trace is None else (1, self.decode_group if self.decode_group > 0 else 4)
        side_decode = trace is None and (S > 1 or self.overlap_decode)
        while len(self._chunk_streams) < (S if S > 1 else 0):
            self._chunk_streams.append(torch.cuda.Stream(device=self.device))
        if side_decode and self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        streams = self._chunk_streams[:S] if S > 1 else [main]
        dec_stream = self._side if side_decode else main
        from .shard import shard_bounds
        bounds = [shard_bounds(B, S, c) for c in range(S)]

        video[0].copy_(images0)   # output frame 0 is the raw input frame (mage_model.py:691)
        self._frame_to_host(video, host_video, 0)
        side = self._side_plan(B) if trace is None and S == 1 and self.backend == "tc" else None
        if side is not None:
            return self._generate_side_by_side(images0, text, speed, noise, video, tokens, tok0_out, host_video, G, *side)
        states = []
        for c, (lo, hi) in enumerate(bounds):
            if streams[c] is not main:
                streams[c].wait_stream(main)
            with torch.cuda.stream(streams[c]):
                states.append(self._prelude(images0[lo:hi], text[lo:hi], speed[lo:hi] if speed is not None else None,
                                            noise[lo:hi] if noise is not None else None, tok0_out[lo:hi], trace))
        img_elems = video.shape[2] * video.shape[3] * video.shape[4]   # frame-major video: images of one frame are adjacent
        for j0 in range(0, L - 1, G):
            j1 = min(j0 + G, L - 1)           # this group generates frames j0+1 .. j1 (steps j0 .. j1-1)
            for c, (lo, hi) in enumerate(bounds):
                with torch.cuda.stream(streams[c]):
                    for j in range(j0, j1):
                        self._decode_step(states[c], j, tokens[j, lo:hi].reshape(-1), trace)
                if dec_stream is not streams[c]:
                    dec_stream.wait_stream(streams[c])
            if trace is not None and trace.get("skip_decode"):
                continue                                   # the objective reads logits only
            # decode the finished group of frames for every sample: video[j0+1 .. j1] (frame-major, contiguous)
            toks = tokens[j0:j1].view(-1, R, R)            # [(j1-j0)*B, R, R], frame-major like the video buffer
            with torch.cuda.stream(dec_stream):
                self.vq.decode_into(toks, video[j0 + 1], img_elems)
                for f in range(j0 + 1, j1 + 1):
                    self._frame_to_host(video, host_video, f)
        for s_ in streams:
            if s_ is not main:
                main.wait_stream(s_)
        if dec_stream is not main:
            main.wait_stream(dec_stream)   # join (also required before a graph capture ends)
        if host_video is not None:
            main.wait_stream(self._copy)

    def _side_plan(self, B: int):
        """(decoder SMs, frames decoded beside the steps, frames per side pass) or None.  A decode step of a FEW prompts is a chain
        of ~40 short dependent kernels that cannot use the machine (M = 2048 rows at 8 prompts: <= 128 CTAs, mostly waiting on
        latencies), while the VQ-VAE decoder of the frames already generated is throughput work nobody waits for.  So at small
        batches the steps get `sms - D` SMs and the decoder runs BESIDE them on a side stream with D SMs (mage_sm_share: the shares
        add up to the machine, so neither sequence ever waits for the other's CTAs); frames that do not fit beside the chain are
        decoded afterwards at full width.  MAGE_SIDE_SMS (D; 0 = off), MAGE_SIDE_FRAMES, MAGE_SIDE_GROUP override the plan."""
        D = self.side_sms
        if D <= 0 or D >= self.n_sms:   # off by default: measured slower (profiles/r02ab_side_by_side_decoder_ab.txt)
            return None
        n_side = self.side_frames if self.side_frames > 0 else (self.L - 1) // 3
        n_side = max(0, min(n_side, self.L - 1))
        if n_side == 0:
            return None
        Gs = self.side_group if self.side_group > 0 else 2
        return D, n_side, Gs

    def _generate_side_by_side(self, images0, text, speed, noise, video, tokens, tok0_out, host_video, G: int, D: int, n_side: int,
                               Gs: int) -> None:
        """The small-batch schedule of `_side_plan`: frames 1..n_side are decoded in passes of Gs frames on the side stream with D
        SMs while the steps continue on `n_sms - D`; the remaining frames after the last step, at full width, G at a time."""
        R, L = self.R, self.L
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        side = self._side
        img_elems = video.shape[2] * video.shape[3] * video.shape[4]

        def decode(j0: int, j1: int) -> None:   # frames j0+1 .. j1 (frame-major, contiguous)
            self.vq.decode_into(tokens[j0:j1].view(-1, R, R), video[j0 + 1], img_elems)
            for f in range(j0 + 1, j1 + 1):
                self._frame_to_host(video, host_video, f)

        st = self._prelude(images0, text, speed, noise, tok0_out, None)
        try:
            ops.sm_share(self.n_sms - D)
            j0 = 0
            for j in range(L - 1):
                self._decode_step(st, j, tokens[j].reshape(-1), None)
                done = j + 1
                if done <= n_side and (done - j0 == Gs or done == n_side):
                    side.wait_stream(main)
                    ops.sm_share(D)
                    with torch.cuda.stream(side):
                        decode(j0, done)
                    ops.sm_share(self.n_sms - D)
                    j0 = done
        finally:
            ops.sm_share(0)
        for j0 in range(n_side, L - 1, G):
            decode(j0, min(j0 + G, L - 1))
        main.wait_stream(side)
        if host_video is not None:
            main.wait_stream(self._copy)

    # ------------------------------------------------------------------ stage-2 objective, forward half
    def _posterior_weights(self) -> dict:
        """Split tensor-core copies of the video posterior (conv3d.*, conv_mu2 | conv_var2), packed on first use.  A 3x3x3 kernel
        [Cout,Cin,kt,ky,kx] becomes a 3x3 kernel over 3*Cin channels [Cout,ky,kx,kt*Cin+ci]: the three temporal taps are folded
        into the channel axis, so a Conv3d is ONE implicit GEMM over the frame triples gathered by `_frame_triples`."""
        if self._post is None:
            sd = self.sd
            if "conv3d.0.conv1.weight" not in sd:
                raise KeyError("this checkpoint was loaded without the train-only video posterior (conv3d.*, conv_mu2, conv_var2): "
                               "build the model with with_posterior=True to evaluate MAGE.forward")
            w = {}
            for i in range(4):
                for c in ("conv1", "conv2", "downsample.0"):
                    k3 = sd[f"conv3d.{i}.{c}.weight"]                                   # [Cout, Cin, 3, 3, 3]
                    w[f"{i}.{c}"] = ops.split(k3.permute(0, 3, 4, 2, 1).reshape(k3.shape[0], 3, 3, 3 * k3.shape[1]).contiguous())
            ml = torch.cat([sd["conv_mu2.weight"], sd["conv_var2.weight"]], 0)          # [128, C, 3, 3]
            w["mu_logvar"] = ops.split(ml.permute(0, 2, 3, 1).contiguous())
            w["mu_logvar.bias"] = torch.cat([sd["conv_mu2.bias"], sd["conv_var2.bias"]]).contiguous()
            ops.check_flag(self.device)
            self._post = w
        return self._post

    @staticmethod
    def _frame_triples(x: torch.Tensor, stride_t: int) -> torch.Tensor:
        """x fp32 [T,B,R,R,C] (frame-major) -> [T',B,R,R,3C]: for every output frame t' the input frames stride_t*t' - 1, +0, +1
        (zeros beyond the clip: Conv3d padding 1) side by side on the channel axis.  Pure data movement."""
        T = x.shape[0]
        To = (T - 1) // stride_t + 1
        z = torch.zeros_like(x[:1])
        xp = torch.cat([z, x, z], 0)
        idx = torch.arange(To, device=x.device) * stride_t
        return torch.cat([xp.index_select(0, idx + kt) for kt in range(3)], -1).contiguous()

    def _conv3d_gn(self, xt: torch.Tensor, wname: str, gn: str, B: int, *, relu: bool, residual: Optional[torch.Tensor] = None):
        """Conv3d 3x3x3 (bias-free, `xt` = its gathered frame triples [T',B,R,R,3C]) -> GroupNorm(16) (+ residual) (ReLU):
        BasicBlock's conv/bn pairs (mage_model.py:266-274).  Returns fp32 [T',B,R,R,C]."""
        To, R, C = xt.shape[0], self.R, self.C
        y, _, _ = ops.conv2d_tc(ops.split(xt.view(To * B, R, R, xt.shape[-1])), self._posterior_weights()[wname], None, pad=(1, 1))
        y = y.view(To * B * R * R, C)
        part = torch.empty(To, B, 16, 2, device=self.device, dtype=torch.float64)
        ops.gn_partial(y, part, B, R * R, groups=16)
        ops.gn_apply(y, part, self.sd[gn + ".weight"], self.sd[gn + ".bias"], B, R * R, relu=relu,
                     residual=residual.view(-1, C) if residual is not None else None, out=y)
        return y.view(To, B, R, R, C)

    def _teacher_forced_ce(self, tok: torch.Tensor, anchor: torch.Tensor) -> torch.Tensor:
        """FlatAxialDecoder.forward + F.cross_entropy(reduction='none') (mage_model.py:374-390, :619) in FULL-SEQUENCE form: all
        L temporal positions go through each block as one batch of images (`_block_seq_tc`: M = L*B*256 rows per GEMM instead of
        L passes over B*256), position 0 = context_linear(anchor), position j+1 = in_linear(features of frame j's GIVEN tokens).
        tok int64 [B,L,R,R]; returns the per-row losses of positions 1..L-1, fp32 [(L-1)*B*R*R] (position-major)."""
        sd, ws, C, R, L = self.sd, self.ws, self.C, self.R, self.L
        p = "generate_model."
        B = tok.shape[0]
        M = B * R * R
        dev = self.device
        x = torch.empty(L * M, C, device=dev, dtype=torch.float32)
        ops.gemm_tc(ops.split(anchor.view(M, C)), ws[p + "context_linear.weight"], self.bias_ctx0, out=x[:M])
        tok_t = tok.permute(1, 0, 2, 3).contiguous()                                   # frame-major [L,B,R,R]
        for j in range(L - 1):
            if self.tok_table is not None:
                ops.token_taps(tok_t[j], self.tok_table, self.tok_posW, self.bias_in_T[j + 1], x[(j + 1) * M:(j + 2) * M])
            else:
                ops.gemm_tc(self._token_features_tc(tok_t[j].reshape(-1), B), ws[p + "in_linear.weight"], self.bias_in_T[j + 1],
                            out=x[(j + 1) * M:(j + 2) * M])
        caches = {i: (torch.empty(M, L, C, device=dev, dtype=torch.float32), torch.empty(M, L, C, device=dev, dtype=torch.float32))
                  for i in range(self.n_blocks) if i % 3 == 0}
        u = torch.empty(2, L * M, C, device=dev, dtype=torch.float16)
        h = torch.empty(2, L * M, 4 * C, device=dev, dtype=torch.float16)
        qkv = torch.empty(L * M, 3 * C, device=dev, dtype=torch.float32)
        for i in range(self.n_blocks):
            self._block_seq_tc(i, x, 0, L, B, caches, u, h, qkv)
        del h, qkv
        logits, _, _ = ops.gemm_tc(ops.split(x[M:]), ws[p + "out.weight"], sd[p + "out.bias"])
        ce = torch.empty((L - 1) * M, device=dev, dtype=torch.float32)
        return ops.cross_entropy_rows(logits, tok_t[1:].reshape(-1), ce)

    def _posterior_sample(self, x: torch.Tensor, eps: torch.Tensor, test_flag: bool):
        """mage_model.py:605-611, 624-625: raw embeddings of all frames x fp32 [L,B,R,R,C] (frame-major) -> four BasicBlocks (each
        halves the frame axis) -> (mu | logvar) -> (z = eps * exp(logvar/2) + mu, or eps itself with test_flag, as [B,64,R,R];
        KL = -0.5 * mean_b sum(1 + logvar - mu^2 - exp(logvar)) as a device scalar)."""
        L, B = x.shape[:2]
        R, C = self.R, self.C
        assert eps is not None and tuple(eps.shape) == (B, 64, R, R)
        for i in range(4):
            p = f"conv3d.{i}"
            xt = self._frame_triples(x, 2)
            res = self._conv3d_gn(xt, f"{i}.downsample.0", p + ".downsample.1", B, relu=False)
            y = self._conv3d_gn(xt, f"{i}.conv1", p + ".bn1", B, relu=True)
            x = self._conv3d_gn(self._frame_triples(y, 1), f"{i}.conv2", p + ".bn2", B, relu=True, residual=res)
        if x.shape[0] != 1:
            raise ValueError(f"frames_length {L} leaves {x.shape[0]} frames after the posterior's four temporal halvings; the "
                             "reference squeezes that axis (mage_model.py:606) and fails as well")
        pw = self._posterior_weights()
        ml, _, _ = ops.conv2d_tc(ops.split(x.view(B, R, R, C)), pw["mu_logvar"], pw["mu_logvar.bias"], pad=(1, 1))
        self.last_mu_logvar = ml.view(B * R * R, -1)                                   # [B*R*R, mu(64) | logvar(64)] (tests)
        z, kl_rows = ops.reparam_kl(ml.view(B * R * R, -1), eps.contiguous().float(), B, R * R)
        z = eps.contiguous().float() if test_flag else z.view(B, 64, R, R)
        return z, ops.scaled_sum(kl_rows, -0.5 / B)                                    # mage_model.py:625

    def forward_loss_continuous(self, latents: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor],
                                eps: Optional[torch.Tensor], test_flag: bool = False) -> dict:
        """MAGE.forward for use_cids=False (MAGE+; mage_model.py:575-639 with :583, :621) in eval mode: latents fp32 [B,L,c,R,R] of
        ALL frames (the first stage's output, whatever module produced it) -> Linear embed -> [posterior, AdaIN] -> teacher-forced
        decoder in full-sequence form -> GroupNorm(32) over all slots -> SiLU -> 1x1x1 conv -> MSE against the latents of frames
        1..L-1.  Returns {'prediction': [1], 'kl_loss': [1] | None}."""
        assert not self.use_cids and self.backend == "tc", "the MAGE+ objective runs on the tensor-core back end"
        sd, ws, C, R, L = self.sd, self.ws, self.C, self.R, self.L
        p = "generate_model."
        B, Lz, c = latents.shape[:3]
        assert Lz == L, f"MAGE.forward needs frames_length = {L} frames, got {Lz}"
        M = B * R * R
        dev = self.device
        T = text.shape[1]
        # Linear embed of every frame's latents, frame-major rows (frame, b, h, w)
        z_rows = latents.permute(1, 0, 3, 4, 2).reshape(L * M, c).contiguous().float()
        emb = ops.gemm(z_rows, self.E, sd["visual_token_embedding.bias"])                # K = c = 4: FFMA kernel
        # 3x3 conv + H/W positions of frames 0..L-2 (the decoder's inputs), all frames as one batch of images
        prior, prior_split, _ = ops.conv2d_tc(ops.split(emb[:(L - 1) * M]).view(2, (L - 1) * B, R, R, C), ws["Wc"], None, pad=(1, 1),
                                              residual=self.posHW, res_mode=3, want=("f32", "split"))
        prior, prior_split = prior.view((L - 1) * M, C), prior_split.view(2, (L - 1) * M, C)
        temb, _ = self._text_encoder(text)
        anchor = self._ma_encoder_tc(prior[:M], prior_split[:, :M], temb, B, T).view(B, R, R, C)
        kl = None
        if self.randomness:
            z, kl = self._posterior_sample(emb.view(L, B, R, R, C), eps, test_flag)
            anchor = self._adain_tc(anchor, z, B)
        if speed is not None:
            ops.add_scaled_vec(anchor, speed, sd["speed_embedding"].view(-1))
        x = torch.empty(L * M, C, device=dev, dtype=torch.float32)
        ops.gemm_tc(ops.split(anchor.view(M, C)), ws[p + "context_linear.weight"], self.bias_ctx0, out=x[:M])
        for j in range(L - 1):
            ops.gemm_tc(prior_split[:, j * M:(j + 1) * M], ws[p + "in_linear.weight"], self.bias_in_T[j + 1], out=x[(j + 1) * M:(j + 2) * M])
        caches = {i: (torch.empty(M, L, C, device=dev, dtype=torch.float32), torch.empty(M, L, C, device=dev, dtype=torch.float32))
                  for i in range(self.n_blocks) if i % 3 == 0}
        u = torch.empty(2, L * M, C, device=dev, dtype=torch.float16)
        h = torch.empty(2, L * M, 4 * C, device=dev, dtype=torch.float16)
        qkv = torch.empty(L * M, 3 * C, device=dev, dtype=torch.float32)
        for i in range(self.n_blocks):
            self._block_seq_tc(i, x, 0, L, B, caches, u, h, qkv)
        del u, h, qkv
        hidden = x[M:]                                                                   # slots 0..L-2 = positions 1..L-1
        part = torch.empty(L - 1, B, 32, 2, device=dev, dtype=torch.float64)
        ops.gn_partial(hidden, part, B, R * R)
        Wo = sd[p + "out.2.weight"].reshape(sd[p + "out.2.weight"].shape[0], C).contiguous()
        pred = ops.gn_silu_head(hidden, part, sd[p + "out.0.weight"], sd[p + "out.0.bias"], Wo, sd[p + "out.2.bias"], B, R * R)
        self.last_prediction = pred.view(L - 1, B, R, R, -1)
        target = z_rows[M:]                                                              # rows (frame 1.., b, h, w) x c
        mse = ops.scaled_sqdiff_sum(pred, target, 1.0 / pred.numel())                    # F.mse_loss, :621
        ops.check_flag(dev)
        return {"prediction": mse, "kl_loss": kl}

    def forward_loss(self, images: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor], eps: Optional[torch.Tensor],
                     test_flag: bool = False, incremental: bool = False) -> dict:
        """MAGE.forward in eval mode (mage_model.py:575-639), the loss terms as device scalars: images [B,L,C,H,W] (all
        frames_length frames) -> VQ tokens -> [randomness: 3-D conv posterior -> (mu, logvar) -> z = eps * exp(logvar / 2) + mu (or
        z = eps with test_flag) -> AdaIN of the motion anchor] -> teacher-forced decoder -> cross-entropy against frames 1..L-1.
        The teacher-forced pass runs in full-sequence form (`_teacher_forced_ce`); `incremental=True` runs it as the sampling
        path's one-position-per-step pass with the given tokens fed back instead -- the same arithmetic (the temporal blocks are
        causal), kept as the cross-check.  Returns {'prediction': [1], 'kl_loss': [1] | None, 'tokens'}."""
        assert self.use_cids and self.backend == "tc", "the objective is built for the token model on the tensor-core back end"
        B, L = images.shape[:2]
        R, C = self.R, self.C
        assert L == self.L, f"MAGE.forward needs frames_length = {self.L} frames (mage_model.py:588,619), got {L}"
        images = images.contiguous().float()
        tok = self.vq.encode(images.view(B * L, *images.shape[2:])).view(B, L, R, R)
        kl = None
        z = None
        if self.randomness:
            # raw token embeddings of ALL frames, frame-major [L,B,R,R,C] (mage_model.py:581,605)
            x = ops.embedding(tok.permute(1, 0, 2, 3).reshape(-1), self.sd["visual_token_embedding.weight"]).view(L, B, R, R, C)
            z, kl = self._posterior_sample(x, eps, test_flag)
        M = B * R * R
        if incremental:
            trace = {"force_tokens": tok[:, 1:].contiguous(), "ce_rows": torch.empty(L - 1, M, device=self.device, dtype=torch.float32),
                     "keep_logits": False, "skip_decode": True}
            self.generate(images[:, 0], text, speed, z, trace=trace)
            ce_rows = trace["ce_rows"]
        else:
            tok0 = tok[:, 0].reshape(B, R * R).contiguous()
            anchor = self._motion_anchor(None, text, speed, z if self.randomness else None, tok0, None)
            ce_rows = self._teacher_forced_ce(tok, anchor)
        self.last_ce_rows = ce_rows.view(L - 1, M)
        pred = ops.scaled_sum(ce_rows, 1.0 / ((L - 1) * M))                             # F.cross_entropy's mean, :619
        ops.check_flag(self.device)
        return {"prediction": pred, "kl_loss": kl, "tokens": tok}

    # ------------------------------------------------------------------ MAGE+ branch (use_cids=False): continuous latents
    def _block_seq_tc(self, i: int, x: torch.Tensor, pos0: int, n_pos: int, B: int, caches, u, h, qkv) -> None:
        """AxialAttentionBlock (mage_model.py:35-53) over `n_pos` CONSECUTIVE temporal positions pos0 .. pos0+n_pos-1 at once (the
        reference-order / full-sequence form): x [n_pos*B*R*R, C] rows ordered (position, b, h, w), updated in place.  GEMMs,
        LayerNorms and the H/W attention see all positions as one batch of images; the temporal block appends and attends
        position by position (causal: position p reads cache entries 0..p, all of which are final by then)."""
        sd, ws, C, R = self.sd, self.ws, self.C, self.R
        p = f"generate_model.blocks.{i}"
        M = B * R * R
        ops.layernorm(x, sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"], out_split=u)
        kind = i % 3
        if kind == 0:
            ops.gemm_tc(u, ws[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"], out=qkv)
            kc, vc = caches[i]
            if n_pos == 1:
                ops.temporal_attn_step(qkv, kc, vc, None, pos0, self.scale, out_split=u)
            else:   # the K/V prefix of a location is staged once for all n_pos queries
                ops.temporal_attn_seq(qkv, kc, vc, pos0, n_pos, self.scale, out_split=u)
            a = u
        else:
            assert R == 16, "the full-sequence path uses the 16x16 axial kernels"
            if self.fused_axial:
                a = torch.empty_like(u)
                ops.qkv_axial_attn_tc(u, ws[p + ".attn.in_proj_weight.axial"], self.axial_bias[i], a, n_img=n_pos * B, R=R,
                                      n_head=self.n_head, axis=kind, scale=self.scale)
            else:
                ops.gemm_tc(u, ws[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"], out=qkv)
                ops.axial_attn(qkv, None, B=n_pos * B, R=R, n_head=self.n_head, axis=kind, scale=self.scale, out_split=u)
                a = u
        ops.gemm_tc(a, ws[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"], residual=x, out=x)
        ops.layernorm(x, sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], out_split=u)
        ops.gemm_tc(u, ws[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"], act=ACT_QUICKGELU, want=(), out_split=h)
        ops.gemm_tc(h, ws[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"], residual=x, out=x)

    def generate_continuous(self, z0: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor] = None,
                            noise: Optional[torch.Tensor] = None, trace: Optional[dict] = None) -> torch.Tensor:
        """MAGE+ sampling between the two first-stage calls (mage_model.py:642-689 with use_cids=False): z0 [B,c,R,R] latents of
        frame 0 -> predicted latents [B, L-1, c, R, R].

        The continuous head normalises over ALL L-1 temporal slots (GroupNorm over [C/32, L-1, H, W], :349-354,386-388), including
        the not-yet-generated slots that the reference pre-fills with frame 0's embedding (:670): slot j's hidden state still
        depends only on slots <= j (causal), but the prediction of slot i at iteration i needs every slot's hidden state.  So the
        one-position-per-step KV-cache form of the VQ path is NOT equivalent here; what is: at iteration i only slots >= i
        changed since iteration i-1 (slot i's input became the real prediction), so the SUFFIX i..L-2 is re-evaluated through the
        six blocks in full-sequence form against the K/V cache of the unchanged prefix, per-slot GroupNorm partial sums of the
        prefix are kept, and the head is evaluated for slot i only (all slots at the last iteration).  (L-1)L/2 position passes
        instead of the reference's (L-1)L, each on a batch of (suffix x B) images -- bit-for-bit the same dataflow otherwise."""
        assert not self.use_cids and self.backend == "tc", "the MAGE+ branch runs on the tensor-core back end"
        assert z0.is_cuda and text.is_cuda
        sd, ws, C, R, L = self.sd, self.ws, self.C, self.R, self.L
        p = "generate_model."
        B, c = z0.shape[0], z0.shape[1]
        T = text.shape[1]
        M = B * R * R
        F_ = L - 1                                                   # slots (frames to predict); slot f sits at temporal position f+1
        dev = self.device
        We, be = self.E, sd["visual_token_embedding.bias"]
        Wo = sd[p + "out.2.weight"].reshape(sd[p + "out.2.weight"].shape[0], C).contiguous()
        bo = sd[p + "out.2.bias"]
        gnw, gnb = sd[p + "out.0.weight"], sd[p + "out.0.bias"]

        def features(lat_rows: torch.Tensor):
            """latents [M, c] (rows (b,h,w)) -> Linear embed -> 3x3 conv + H/W pos: (fp32 [M,C], split [2,M,C])."""
            emb = ops.gemm(lat_rows, We, be)                          # K = c = 4: FFMA kernel
            o, f, _ = ops.conv2d_tc(ops.split(emb).view(2, B, R, R, C), ws["Wc"], None, pad=(1, 1), residual=self.posHW, res_mode=3,
                                    want=("f32", "split"))
            return o.view(M, C), f.view(2, M, C)

        if noise is None and self.randomness:
            raise AssertionError("randomness=True needs the N(0,1) noise [B,64,R,R]")
        z_rows = z0.permute(0, 2, 3, 1).reshape(M, c).contiguous().float()
        f0, f0_split = features(z_rows)
        temb, _ = self._text_encoder(text)
        anchor = self._ma_encoder_tc(f0, f0_split, temb, B, T).view(B, R, R, C)
        if noise is not None and self.randomness:
            anchor = self._adain_tc(anchor, noise, B)
        if speed is not None:
            ops.add_scaled_vec(anchor, speed, sd["speed_embedding"].view(-1))
        caches = {i: (torch.empty(M, L, C, device=dev, dtype=torch.float32), torch.empty(M, L, C, device=dev, dtype=torch.float32))
                  for i in range(self.n_blocks) if i % 3 == 0}
        # temporal position 0 = the motion anchor: one pass, fills cache entry 0
        x0, _, _ = ops.gemm_tc(ops.split(anchor.view(M, C)), ws[p + "context_linear.weight"], self.bias_ctx0)
        u1 = torch.empty(2, M, C, device=dev, dtype=torch.float16)
        h1 = torch.empty(2, M, 4 * C, device=dev, dtype=torch.float16)
        q1 = torch.empty(M, 3 * C, device=dev, dtype=torch.float32)
        for i in range(self.n_blocks):
            self._block_seq_tc(i, x0, 0, 1, B, caches, u1, h1, q1)
        del x0, u1, h1, q1
        hidden = torch.empty(F_, M, C, device=dev, dtype=torch.float32)           # final hidden state of every slot
        part = torch.empty(F_, B, 32, 2, device=dev, dtype=torch.float64)         # GroupNorm partial sums per slot
        f_real_split = f0_split                                                   # features of the slot that just became real
        pred = None
        for i in range(F_):
            n_s = F_ - i
            x = hidden[i:].view(n_s * M, C)
            # in_linear (+ T_pos of the slot): slot i reads the newest real features, the later slots frame 0's (mage_model.py:670)
            for s_ in range(n_s):
                ops.gemm_tc(f_real_split if s_ == 0 else f0_split, ws[p + "in_linear.weight"], self.bias_in_T[i + s_ + 1],
                            out=x[s_ * M:(s_ + 1) * M])
            u = torch.empty(2, n_s * M, C, device=dev, dtype=torch.float16)
            h = torch.empty(2, n_s * M, 4 * C, device=dev, dtype=torch.float16)
            qkv = torch.empty(n_s * M, 3 * C, device=dev, dtype=torch.float32)
            for blk in range(self.n_blocks):
                self._block_seq_tc(blk, x, i + 1, n_s, B, caches, u, h, qkv)
            del u, h, qkv
            ops.gn_partial(x, part[i:], B, R * R)
            if i + 1 < F_:
                pred_i = ops.gn_silu_head(hidden[i], part, gnw, gnb, Wo, bo, B, R * R)      # [M, c]: slot i only
                if trace is not None:
                    # parity diagnostics: record what this iteration would feed forward and, if asked, feed the GIVEN latents
                    # instead (teacher forcing: every iteration is then compared under identical inputs, no accumulation)
                    trace.setdefault("step_pred", []).append(pred_i.view(B, R, R, -1).permute(0, 3, 1, 2).clone())
                    if trace.get("force_step_pred") is not None:
                        pred_i = trace["force_step_pred"][:, i].permute(0, 2, 3, 1).reshape(M, -1).contiguous().float()
                _, f_real_split = features(pred_i)
            else:
                pred = ops.gn_silu_head(hidden.view(F_ * M, C), part, gnw, gnb, Wo, bo, B, R * R)   # every slot (mage_model.py:689)
        ops.check_flag(dev)
        if trace is not None and "step_pred" in trace:
            trace["step_pred"] = torch.stack(trace["step_pred"], 1)   # [B, L-2, c, R, R]
        return pred.view(F_, B, R, R, -1).permute(1, 0, 4, 2, 3).contiguous()

    def _frame_to_host(self, video: torch.Tensor, host_video: Optional[torch.Tensor], f: int) -> None:
        """Queue the D2H copy of frame f (one contiguous [B,C,H,W] block) on the copy stream, after the work queued so far."""
        if host_video is None:
            return
        cur = torch.cuda.current_stream()
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=self.device)
        self._copy.wait_stream(cur)
        with torch.cuda.stream(self._copy):
            host_video[f].copy_(video[f], non_blocking=True)

    def generate(self, images0: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None, trace: Optional[dict] = None, to_host: bool = False):
        """Returns (video [B,L,C,H,W] with frame 0 = images0, tokens i64 [B,L-1,R,R], tok0 i64 [B,R,R]).  The video is a view of a
        frame-major buffer that the next call overwrites.  to_host=True: the video comes back as a pinned HOST tensor, every
        frame having been copied out while later frames were still being generated (the call returns synchronised)."""
        assert images0.is_cuda and text.is_cuda and text.dtype == torch.int64
        if self.randomness:
            assert noise is not None, "randomness=True needs the N(0,1) noise [B,64,R,R] (drawn by the caller on the CPU, mage_model.py:661)"
        else:
            noise = None
        B, T = text.shape
        R, L = self.R, self.L
        images0 = images0.contiguous().float()
        if not self.use_cuda_graph or trace is not None:
            video = torch.empty(L, B, *images0.shape[1:], device=self.device, dtype=torch.float32)
            host = torch.empty(L, B, *images0.shape[1:], dtype=torch.float32, pin_memory=True) if to_host else None
            tokens = torch.empty(L - 1, B, R * R, device=self.device, dtype=torch.int64)
            tok0 = torch.empty(B, R * R, device=self.device, dtype=torch.int64)
            n0 = ops.launch_count()
            self._generate_impl(images0, text, speed, noise, video, tokens, tok0, trace, host)
            self.kernels_per_generate = ops.launch_count() - n0
            ops.check_flag(self.device)
            if to_host:
                torch.cuda.current_stream().synchronize()
                video = host
            return video.permute(1, 0, 2, 3, 4), tokens.permute(1, 0, 2).reshape(B, L - 1, R, R), tok0.view(B, R, R).clone()

        key = (B, T, tuple(images0.shape[1:]), speed is not None, noise is not None, to_host, self._plan(B), self.overlap_decode, self._side_plan(B),
               self.fused_ln_taps, self.fused_ln, self.fused_axial)
        st = self._graphs.pop(key, None)
        if st is None:
            st = self._capture(key, images0, text, speed, noise, to_host)
        self._graphs[key] = st     # most recently used last (dicts keep insertion order)
        st["images0"].copy_(images0)
        st["text"].copy_(text)
        if speed is not None:
            st["speed"].copy_(speed)
        if noise is not None:
            st["noise"].copy_(noise)
        st["graph"].replay()
        ops.check_flag(self.device)  # loud failure if an operand left the fp16 split range / a caption id left the vocabulary (syncs)
        video = st["video"]
        if to_host:
            torch.cuda.current_stream().synchronize()
            video = st["host_video"]
        return video.permute(1, 0, 2, 3, 4), st["tokens"].permute(1, 0, 2).reshape(B, L - 1, R, R), st["tok0"].view(B, R, R).clone()

    def _capture(self, key, images0, text, speed, noise, to_host=False):
        """Capture one whole generate call for this input signature.  All graphs of an engine replay serially, so they are captured
        into ONE shared memory pool (the K/V caches, activations and decoder maps of different signatures alias each other: the
        pool holds the largest signature, not the sum), and the cache keeps at most `max_graphs` signatures (LRU) -- captions
        cannot be padded to a common length because the motion anchor attends padded positions (mage_model.py:92), so real
        data produces one signature per caption length."""
        B, T = text.shape
        R, L = self.R, self.L
        dev = self.device
        while len(self._graphs) >= self.max_graphs:
            old = next(iter(self._graphs))
            self._graphs.pop(old)["graph"].reset()
        st = dict(images0=images0.clone(), text=text.clone(),
                  speed=speed.clone() if speed is not None else None,
                  noise=noise.clone() if noise is not None else None,
                  video=torch.empty(L, B, *images0.shape[1:], device=dev, dtype=torch.float32),
                  host_video=torch.empty(L, B, *images0.shape[1:], dtype=torch.float32, pin_memory=True) if to_host else None,
                  tokens=torch.empty(L - 1, B, R * R, device=dev, dtype=torch.int64),
                  tok0=torch.empty(B, R * R, device=dev, dtype=torch.int64))
        args = (st["images0"], st["text"], st["speed"], st["noise"], st["video"], st["tokens"], st["tok0"], None, st["host_video"])
        # warm-up on a side stream (sets kernel attributes, fills the allocator), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._generate_impl(*args)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        ops.check_flag(dev)
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(g, pool=self._pool):
            self._generate_impl(*args)
        self.kernels_per_generate = ops.launch_count() - n0
        st["graph"] = g
        return st
