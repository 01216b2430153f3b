"""Seeded synthetic checkpoints and batches for the MAGE sampling path.

The reference ships no checkpoints (README.md:8,41 are Google-Drive links) and no
tests, so every parity check in this repo runs on a *synthetic* state dict that
uses the reference's key names and shapes and is produced
deterministically from a seed with the CPU generator.  The same generator feeds
the golden-vector script (oracle/make_golden.py, run against /root/reference),
the CPU oracle, the CUDA path, bench.py and smoke().

Nothing here computes anything on the sampling path; it only fabricates inputs.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

Spec = List[Tuple[str, Tuple[int, ...], str]]

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# --------------------------------------------------------------------------------------
# model configs (same schema as the reference's config/*.yaml `model.params`)
# --------------------------------------------------------------------------------------
def model_params(family: str = "caterv2", frames_length: int = 10, randomness: Optional[bool] = None) -> dict:
    """`model.params` dict in the schema of /root/reference/config/mage_caterv2.yaml:11-53.

    family: 'caterv2' (f8 VQ-VAE, vocab 50, ctx 38), 'caterv1' (f8, vocab 30, ctx 32),
            'mnist' (f4 VQ-VAE, 1 channel, vocab 30; authored here, defaults of
            /root/reference/train_vqvae.py:197-200).
    """
    if family == "caterv2":
        fs = dict(input_dim=3, dim=256, down_ratio=8, K=512)
        vocab, ctx, rnd = 50, 38, True
    elif family == "caterv1":
        fs = dict(input_dim=3, dim=256, down_ratio=8, K=512)
        vocab, ctx, rnd = 30, 32, True
    elif family == "mnist":
        fs = dict(input_dim=1, dim=256, down_ratio=4, K=512)
        vocab, ctx, rnd = 30, 32, False
    elif family == "caterv2plus":
        # MAGE+ (config/mage+_caterv2.yaml): use_cids=False, 4-channel continuous latents.  The shipped first stage is
        # latent-diffusion's AutoencoderKL (not vendored); PatchLatentAE below stands in for it in tests / goldens.
        fs = dict(in_channels=3, embed_dim=4, down_ratio=8, seed=5)
        vocab, ctx, rnd = 50, 38, True
    else:
        raise KeyError(family)
    if randomness is not None:
        rnd = randomness
    plus = family == "caterv2plus"
    return dict(
        codebook_size=512, frames_length=frames_length, image_resolution=16, vision_width=512,
        dropout=0.1, use_cids=not plus, randomness=rnd, alpha=0.0001, beta=0.0005,
        first_stage_config=(dict(target="mage_b200.synthetic.PatchLatentAE", params=dict(**fs)) if plus else
                            dict(target="modules.vqvae_model.VectorQuantizedVAE", params=dict(ckpt_path=None, **fs))),
        text_encoder_config=dict(target="modules.mage_model.TransformerTextEncoder",
                                 params=dict(vocab_size=vocab, context_length=ctx, transformer_width=512,
                                             transformer_layers=2, output_dim=512, padding_idx=0, dropout=0.1)),
        ma_config=dict(target="modules.mage_model.MAEncoder", params=dict(layers=1, d_model=512)),
        generate_decoder_config=dict(target="modules.mage_model.FlatAxialDecoder",
                                     params=dict(in_channels=512, out_channels=4 if plus else 512, model_channels=512,
                                                 frames_length=frames_length, layers=6)),
    )


class PatchLatentAE(torch.nn.Module):
    """Stand-in first stage for the MAGE+ branch in tests and goldens: a seeded linear patch autoencoder with the interface the
    reference expects of `first_stage_model` when use_cids=False (mage_model.py:530-567): `.embed_dim`, `.encode(x[N,C,H,W]) ->
    Tensor [N,embed_dim,H/r,W/r]` (the reference also accepts an object with `.sample()`), `.decode(z) -> [N,C,H,W]`.
    Plain PyTorch on purpose: the real first stage (latent-diffusion's AutoencoderKL) is an external module that this repo
    treats as given -- parity unpinned for it, pinned for everything between the two calls."""

    def __init__(self, in_channels: int = 3, embed_dim: int = 4, down_ratio: int = 8, seed: int = 5):
        super().__init__()
        g = torch.Generator(device="cpu")
        g.manual_seed(seed)
        k = down_ratio
        self.embed_dim, self.down_ratio = embed_dim, down_ratio
        self.register_buffer("enc_w", torch.randn(embed_dim, in_channels, k, k, generator=g) * (2.0 / (k * math.sqrt(in_channels))))
        self.register_buffer("dec_w", torch.randn(embed_dim, in_channels, k, k, generator=g) * 0.5)

    # Patchify + plain fp32 matmul instead of conv2d / conv_transpose2d: on CUDA PyTorch's default lets cuDNN run convolutions in
    # TF32, which at batch 64 moved the latents by ~1e-3 -- the stand-in must be the same function on every device.
    @torch.no_grad()
    def encode(self, x):
        n, c, h, w = x.shape
        k = self.down_ratio
        p = x.float().reshape(n, c, h // k, k, w // k, k).permute(0, 2, 4, 1, 3, 5).reshape(n, h // k, w // k, c * k * k)
        z = p @ self.enc_w.reshape(self.embed_dim, -1).t()
        return z.permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def decode(self, z):
        n, e, h, w = z.shape
        k = self.down_ratio
        c = self.dec_w.shape[1]
        y = z.float().permute(0, 2, 3, 1) @ self.dec_w.reshape(e, -1)                 # [n, h, w, c*k*k]
        y = y.reshape(n, h, w, c, k, k).permute(0, 3, 1, 4, 2, 5).reshape(n, c, h * k, w * k)
        return torch.tanh(y)


# --------------------------------------------------------------------------------------
# parameter specs: (state_dict key, shape, init kind)
# --------------------------------------------------------------------------------------
def _conv(spec: Spec, name: str, cout: int, cin: int, k: int, bias: bool = True, transposed: bool = False):
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    spec.append((name + ".weight", shape, "conv"))
    if bias:
        spec.append((name + ".bias", (cout,), "bias"))


def _bn(spec: Spec, name: str, c: int):
    spec.append((name + ".weight", (c,), "ln_w"))
    spec.append((name + ".bias", (c,), "ln_b"))
    spec.append((name + ".running_mean", (c,), "bn_mean"))
    spec.append((name + ".running_var", (c,), "bn_var"))
    spec.append((name + ".num_batches_tracked", (), "count"))


def vqvae_param_spec(input_dim: int, dim: int, down_ratio: int, K: int = 512, **_) -> Spec:
    """Keys of VectorQuantizedVAE.state_dict() (/root/reference/modules/vqvae_model.py:168-217)."""
    s: Spec = []
    if down_ratio == 4:
        _conv(s, "encoder.0", dim, input_dim, 4)
        _bn(s, "encoder.1", dim)
        _conv(s, "encoder.3", dim, dim, 4)
        for blk in ("encoder.4", "encoder.5", "decoder.0", "decoder.1"):
            _conv(s, blk + ".block.1", dim, dim, 3)
            _bn(s, blk + ".block.2", dim)
            _conv(s, blk + ".block.4", dim, dim, 1)
            _bn(s, blk + ".block.5", dim)
        _conv(s, "decoder.3", dim, dim, 4, transposed=True)
        _bn(s, "decoder.4", dim)
        _conv(s, "decoder.6", input_dim, dim, 4, transposed=True)
        s.append(("codebook.embedding.weight", (K, dim), "codebook"))
    elif down_ratio == 8:
        _conv(s, "encoder.0", dim, input_dim, 7)

        def enc_block(name, cin, cout):
            hid = cout // 4
            if cin != cout:
                _conv(s, name + ".id_path", cout, cin, 1)
            _conv(s, name + ".block.1", hid, cin, 3)
            _conv(s, name + ".block.3", hid, hid, 3)
            _conv(s, name + ".block.5", hid, hid, 3)
            _conv(s, name + ".block.7", cout, hid, 1)

        def dec_block(name, cin, cout):
            hid = cout // 4
            if cin != cout:
                _conv(s, name + ".id_path", cout, cin, 1)
            _conv(s, name + ".block.1", hid, cin, 1)
            _conv(s, name + ".block.3", hid, hid, 3)
            _conv(s, name + ".block.5", hid, hid, 3)
            _conv(s, name + ".block.7", cout, hid, 3)

        enc_block("encoder.1", dim, dim)
        enc_block("encoder.3", dim, dim)
        enc_block("encoder.5", dim, 2 * dim)
        enc_block("encoder.7", 2 * dim, 4 * dim)
        dec_block("decoder.0", 4 * dim, 2 * dim)
        dec_block("decoder.2", 2 * dim, dim)
        dec_block("decoder.4", dim, dim)
        dec_block("decoder.6", dim, dim)
        _conv(s, "decoder.8", input_dim, dim, 1)
        s.append(("codebook.embedding.weight", (K, 4 * dim), "codebook"))
    else:
        raise ValueError(f"down_ratio {down_ratio}")
    return s


def _ln(spec: Spec, name: str, c: int):
    spec.append((name + ".weight", (c,), "ln_w"))
    spec.append((name + ".bias", (c,), "ln_b"))


def _mha(spec: Spec, name: str, d: int, w_std: float, o_std: float):
    spec.append((name + ".in_proj_weight", (3 * d, d), f"normal:{w_std}"))
    spec.append((name + ".in_proj_bias", (3 * d,), "bias"))
    spec.append((name + ".out_proj.weight", (d, d), f"normal:{o_std}"))
    spec.append((name + ".out_proj.bias", (d,), "bias"))


def mage_param_spec(params: dict) -> Spec:
    """Keys of MAGE.state_dict() that the sampling path reads.

    The train-only tensors (`conv3d.*`, `conv_mu2`, `conv_var2`) are left out; the
    drop-in accepts them in `load_state_dict` and ignores them.
    """
    d = params["vision_width"]
    K = params["codebook_size"]
    R = params["image_resolution"]
    s: Spec = []
    fs = params["first_stage_config"]["params"]
    if params["use_cids"]:
        for n, shp, kind in vqvae_param_spec(**{k: v for k, v in fs.items() if k != "ckpt_path"}):
            s.append(("first_stage_model." + n, shp, kind))

    te = params["text_encoder_config"]["params"]
    w = te["transformer_width"]
    for i in range(te["transformer_layers"]):
        p = f"text_encoder.transformer.layers.{i}"
        _mha(s, p + ".self_attn", w, 0.02, 0.02)
        s.append((p + ".linear1.weight", (4 * w, w), "normal:0.02"))
        s.append((p + ".linear1.bias", (4 * w,), "bias"))
        s.append((p + ".linear2.weight", (w, 4 * w), "normal:0.02"))
        s.append((p + ".linear2.bias", (w,), "bias"))
        _ln(s, p + ".norm1", w)
        _ln(s, p + ".norm2", w)
    s.append(("text_encoder.token_embedding.weight", (te["vocab_size"], w), "embed_pad0"))
    s.append(("text_encoder.positions.weight", (te["context_length"], w), "normal:0.02"))
    _ln(s, "text_encoder.layer_norm", w)
    _ln(s, "text_encoder.ln_text_final", w)
    s.append(("text_encoder.text_projection.weight", (te["output_dim"], w), "normal:0.02"))
    s.append(("text_encoder.text_projection.bias", (te["output_dim"],), "bias"))

    ma = params["ma_config"]["params"]
    dm = ma["d_model"]
    for i in range(ma["layers"]):
        p = f"ma_encoder.blocks.{i}"
        _mha(s, p + ".attn", dm, dm ** -0.5, dm ** -0.5)
        _ln(s, p + ".ln_q", dm)
        _ln(s, p + ".ln_kv", dm)
        s.append((p + ".mlp.c_fc.weight", (4 * dm, dm), f"normal:{(2 * dm) ** -0.5}"))
        s.append((p + ".mlp.c_fc.bias", (4 * dm,), "bias"))
        s.append((p + ".mlp.c_proj.weight", (dm, 4 * dm), f"normal:{(2 * dm) ** -0.5 * 0.5}"))
        s.append((p + ".mlp.c_proj.bias", (dm,), "bias"))
        _ln(s, p + ".ln_2", dm)

    gd = params["generate_decoder_config"]["params"]
    mc, layers, L = gd["model_channels"], gd["layers"], gd["frames_length"]
    proj_std = (mc ** -0.5) * ((2 * layers) ** -0.5)
    s.append(("generate_model.T_positional_embedding", (L, 1, 1, mc), f"normal:{mc ** -0.5}"))
    s.append(("generate_model.in_linear.weight", (mc, gd["in_channels"]), f"normal:{gd['in_channels'] ** -0.5}"))
    s.append(("generate_model.in_linear.bias", (mc,), "bias"))
    s.append(("generate_model.context_linear.weight", (mc, dm), f"normal:{dm ** -0.5}"))
    s.append(("generate_model.context_linear.bias", (mc,), "bias"))
    for i in range(layers):
        p = f"generate_model.blocks.{i}"
        _mha(s, p + ".attn", mc, mc ** -0.5, proj_std)
        _ln(s, p + ".ln_1", mc)
        s.append((p + ".mlp.c_fc.weight", (4 * mc, mc), f"normal:{(2 * mc) ** -0.5}"))
        s.append((p + ".mlp.c_fc.bias", (4 * mc,), "bias"))
        s.append((p + ".mlp.c_proj.weight", (mc, 4 * mc), f"normal:{proj_std}"))
        s.append((p + ".mlp.c_proj.bias", (mc,), "bias"))
        _ln(s, p + ".ln_2", mc)
    if params["use_cids"]:
        s.append(("generate_model.out.weight", (gd["out_channels"], mc), f"normal:{mc ** -0.5}"))
        s.append(("generate_model.out.bias", (gd["out_channels"],), "bias"))
        s.append(("visual_token_embedding.weight", (K, d), "normal:1.0"))
    else:
        # MAGE+ head: GroupNorm(32, mc) -> SiLU -> Conv3d(mc, out, 1) (mage_model.py:349-354; zero-initialised in the reference,
        # random here so that the autoregression is exercised); Linear embed of the latents (:482-483)
        _ln(s, "generate_model.out.0", mc)
        s.append(("generate_model.out.2.weight", (gd["out_channels"], mc, 1, 1, 1), f"normal:{mc ** -0.5}"))
        s.append(("generate_model.out.2.bias", (gd["out_channels"],), "bias"))
        s.append(("visual_token_embedding.weight", (d, fs["embed_dim"]), "normal:0.5"))
        s.append(("visual_token_embedding.bias", (d,), "bias"))

    s.append(("conv.0.weight", (d, d, 3, 3), "conv"))
    s.append(("speed_embedding", (1, d), f"normal:{d ** -0.5}"))
    s.append(("H_positional_embedding", (1, R, 1, d), f"normal:{d ** -0.5}"))
    s.append(("W_positional_embedding", (1, 1, R, d), f"normal:{d ** -0.5}"))
    if params["randomness"]:
        s.append(("conv_d2.weight", (d, 64, 3, 3), "conv"))
        for br in ("conv_mu", "conv_var"):
            _conv(s, f"adain.{br}.0", d, d, 3)
            _conv(s, f"adain.{br}.1", d, d, 3)
    return s


def posterior_param_spec(params: dict) -> Spec:
    """The train-only tensors of MAGE (randomness=True): the 3-D convolutional video posterior `conv3d` (four BasicBlocks,
    mage_model.py:264-297, 496-501) and the two Gaussian heads conv_mu2 / conv_var2 (:502-503).  Read by MAGE.forward only."""
    d, dm = params["vision_width"], params["ma_config"]["params"]["d_model"]
    s: Spec = []
    for i in range(4):
        cout = dm if i == 3 else d
        p = f"conv3d.{i}"
        s.append((p + ".conv1.weight", (cout, d, 3, 3, 3), "conv"))
        _ln(s, p + ".bn1", cout)
        s.append((p + ".conv2.weight", (cout, cout, 3, 3, 3), "conv"))
        _ln(s, p + ".bn2", cout)
        s.append((p + ".downsample.0.weight", (cout, d, 3, 3, 3), "conv"))
        _ln(s, p + ".downsample.1", cout)
    _conv(s, "conv_mu2", 64, d, 3)
    _conv(s, "conv_var2", 64, d, 3)
    return s


def make_state_dict(spec: Spec, seed: int) -> Dict[str, torch.Tensor]:
    """Fill a spec with seeded CPU-generator values.  Every tensor gets its own
    generator (seed, index) so inserting a key never shifts the others."""
    sd: Dict[str, torch.Tensor] = {}
    for i, (name, shape, kind) in enumerate(spec):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 100003 + i)
        if kind == "conv":
            taps = math.prod(shape[2:])   # 2-D and 3-D kernels
            fan_in = shape[1] * taps
            fan_out = shape[0] * taps
            a = math.sqrt(6.0 / (fan_in + fan_out)) * 1.4  # xavier-uniform with a ReLU-ish gain
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "ln_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "bn_mean":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "bn_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "count":
            t = torch.tensor(1, dtype=torch.long)
        elif kind == "codebook":
            t = torch.randn(shape, generator=g) * 0.5
        elif kind == "embed_pad0":
            t = torch.randn(shape, generator=g) * 0.02
            t[0].zero_()
        elif kind.startswith("normal:"):
            t = torch.randn(shape, generator=g) * float(kind.split(":")[1])
        else:
            raise KeyError(kind)
        sd[name] = t.contiguous()
    return sd


def conditioned_codebook(down_ratio: int) -> torch.Tensor:
    """Well-conditioned codebook fixture (tests/golden/codebook_f{4,8}.npy): farthest-point
    sample of this checkpoint family's own encoder outputs, rounded to fp16-representable
    values (see oracle/make_golden.py).  Default U(+-1/K) codebooks make
    the reference's own argmin ill-conditioned (fp32 vs fp64 already disagree)."""
    path = os.path.join(GOLDEN_DIR, f"codebook_f{down_ratio}.npy")
    return torch.from_numpy(np.load(path).astype(np.float32))


def make_vqvae_state_dict(fs_params: dict, seed: int = 7, conditioned: bool = True) -> Dict[str, torch.Tensor]:
    p = {k: v for k, v in fs_params.items() if k not in ("ckpt_path", "ignore_keys")}
    sd = make_state_dict(vqvae_param_spec(**p), seed)
    if conditioned:
        sd["codebook.embedding.weight"] = conditioned_codebook(p["down_ratio"])
    return sd


def make_mage_state_dict(params: dict, seed: int = 11, vq_seed: int = 7, conditioned: bool = True,
                         posterior: bool = False) -> Dict[str, torch.Tensor]:
    """Synthetic MAGE checkpoint `state_dict` for `params`: the sampling subset, plus (posterior=True) the train-only video
    posterior that MAGE.forward reads.  The sampling tensors do not depend on `posterior`."""
    sd = make_state_dict(mage_param_spec(params), seed)
    if posterior and params["randomness"]:
        sd.update(make_state_dict(posterior_param_spec(params), seed + 1))
    if not params["use_cids"]:
        for k, v in PatchLatentAE(**params["first_stage_config"]["params"]).state_dict().items():
            sd["first_stage_model." + k] = v
        return sd
    fs = make_vqvae_state_dict(params["first_stage_config"]["params"], vq_seed, conditioned)
    for k, v in fs.items():
        sd["first_stage_model." + k] = v
    return sd


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def structured_images(n: int, channels: int, res: int, seed: int, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    """Flat background + 6 random rectangles + 0.02 noise (uniform-noise images collapse the
    VQ-VAE to ~13 codes).  Returns [n, channels, res, res] in [lo, hi]."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    imgs = torch.empty(n, channels, res, res)
    for i in range(n):
        bg = torch.rand(channels, generator=g)
        img = bg.view(channels, 1, 1).expand(channels, res, res).clone()
        for _ in range(6):
            x0, y0 = [int(v) for v in torch.randint(0, res - 8, (2,), generator=g)]
            w, h = [int(v) for v in torch.randint(8, res // 2, (2,), generator=g)]
            col = torch.rand(channels, generator=g)
            img[:, y0:min(res, y0 + h), x0:min(res, x0 + w)] = col.view(channels, 1, 1)
        img = img + 0.02 * torch.randn(channels, res, res, generator=g)
        imgs[i] = img.clamp_(0, 1)
    return imgs * (hi - lo) + lo


def make_batch(params: dict, batch: int, seed: int = 1234, text_len: int = 20, padded: bool = False,
               frames: int = 1, with_speed: bool = True) -> Dict[str, torch.Tensor]:
    """Synthetic batch dict with the reference's contract (dataload.py:260,370):
    'images' f32 [B,frames,C,R,R], 'text' i64 [B,T] = [CLS]=1 ... [SEP]=2 (pad 0), 'speed' f32 [B].
    Only images[:,0] is read by sampling (mage_model.py:642,691)."""
    fs = params["first_stage_config"]["params"]
    C, res = fs.get("input_dim", fs.get("in_channels")), params["image_resolution"] * fs["down_ratio"]
    lo, hi = ((-0.5, 0.5) if fs["down_ratio"] == 4 else (-1.0, 1.0))
    imgs = structured_images(batch * frames, C, res, seed, lo, hi).view(batch, frames, C, res, res)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 1)
    vocab = params["text_encoder_config"]["params"]["vocab_size"]
    text = torch.zeros(batch, text_len, dtype=torch.long)
    for b in range(batch):
        n = text_len if not padded or b == 0 else int(torch.randint(8, text_len + 1, (1,), generator=g))
        text[b, 0] = 1
        text[b, 1:n - 1] = torch.randint(3, vocab, (n - 2,), generator=g)
        text[b, n - 1] = 2
    out = {"images": imgs, "text": text}
    if with_speed:
        out["speed"] = torch.rand(batch, generator=g)
    return out


def make_noise(batch: int, res: int = 16, seed: int = 99) -> torch.Tensor:
    """N(0,1) [B,64,res,res] as the reference draws it on the CPU (mage_model.py:661)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(batch, 64, res, res, generator=g)
