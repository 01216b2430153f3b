"""Drop-in classes for the reference's config-addressed plugin boundary.

`modules.mage_model.MAGE`, `modules.vqvae_model.VectorQuantizedVAE` and the sub-module classes
named as `target:` in config/*.yaml resolve (through the thin `modules/` package at the repo
root) to the classes below.  They keep the constructor kwargs, `state_dict` key names/shapes and
the two entry points of the sampling path -- `MAGE.autoregressive_generate(batch)` and
`VectorQuantizedVAE.encode / decode` -- and run them on libmage_sm100.so.  They are parameter
containers, not torch.nn compute graphs: there is no eager fallback, a CPU-resident model
refuses to sample.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib
from . import synthetic as syn
from .config import instantiate_from_config

_TRAIN_ONLY_PREFIXES = ("conv3d.", "conv_mu2.", "conv_var2.")  # mage_model.py:496-503, never run at sampling


class ParamTree(nn.Module):
    """nn.Module whose parameters/buffers are created from (dotted key, shape, kind) specs so
    that `state_dict()` reproduces the reference's key layout exactly."""

    def __init__(self, spec: Optional[syn.Spec] = None, seed: int = 0):
        super().__init__()
        if spec:
            values = syn.make_state_dict(spec, seed)
            for name, _, kind in spec:
                self._add(name.split("."), values[name], buffer=kind in ("bn_mean", "bn_var", "count"))

    def _add(self, parts: List[str], value: torch.Tensor, buffer: bool):
        if len(parts) == 1:
            if buffer:
                self.register_buffer(parts[0], value)
            else:
                self.register_parameter(parts[0], nn.Parameter(value, requires_grad=False))
            return
        child = self._modules.get(parts[0])
        if child is None:
            child = ParamTree()
            self.add_module(parts[0], child)
        child._add(parts[1:], value, buffer)


def _strip(spec: syn.Spec, prefix: str) -> syn.Spec:
    return [(n[len(prefix):], s, k) for n, s, k in spec if n.startswith(prefix)]


class _EngineOwner(nn.Module):
    """Owns the packed-weight engine (split fp16 copies, per-code tables, folded BatchNorm) and rebuilds it whenever the
    parameters it was packed from have changed: the engine is keyed on every parameter's / buffer's storage address and
    in-place version counter, so `.to()`, `load_state_dict` on this module OR on any child (e.g. `first_stage_model.
    init_from_ckpt`), and in-place edits all invalidate it -- never a silent sample from stale weights."""

    def __init__(self):
        super().__init__()
        self._engine = None
        self._engine_sig = None
        self._wide_engine = None

    def _signature(self):
        return tuple((t.data_ptr(), t._version, t.device.index) for t in list(self.parameters()) + list(self.buffers()))

    def _engine_valid(self) -> bool:
        return self._engine is not None and self._engine_sig == self._signature()

    def invalidate(self) -> None:
        """Drop the packed-weight engine explicitly (it is rebuilt on the next call)."""
        self._engine = None
        self._engine_sig = None
        self._wide_engine = None

    def _apply(self, fn, *a, **kw):
        self.invalidate()
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self.invalidate()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _cuda_state(self) -> Dict[str, torch.Tensor]:
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        dev = next(iter(sd.values())).device
        if dev.type != "cuda":
            raise RuntimeError("mage_b200 runs only on a CUDA device (sm_100a kernels, no CPU fallback): call .to('cuda') first")
        return {k: (v.float().contiguous() if v.is_floating_point() else v) for k, v in sd.items()}


class VectorQuantizedVAE(_EngineOwner):
    """modules.vqvae_model.VectorQuantizedVAE (vqvae_model.py:168-248): same ctor, keys, encode/decode."""

    def __init__(self, input_dim, down_ratio, dim, K=512, ckpt_path=None, ignore_keys=[]):
        super().__init__()
        spec = syn.vqvae_param_spec(input_dim, dim, down_ratio, K)
        self.input_dim, self.down_ratio, self.dim, self.K = input_dim, down_ratio, dim, K
        self.encoder = ParamTree(_strip(spec, "encoder."), seed=7)
        self.decoder = ParamTree(_strip(spec, "decoder."), seed=8)
        self.codebook = ParamTree(_strip(spec, "codebook."), seed=9)
        self.embed_dim = dim if down_ratio == 4 else 4 * dim
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=list()):
        """vqvae_model.py:222-231 (strict=False, keys starting with an ignore prefix dropped)."""
        sd = torch.load(path, map_location="cpu")
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                print("Deleting key {} from state_dict.".format(k))
                del sd[k]
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    def engine(self):
        if not self._engine_valid():
            from .engine import VQVAEEngine
            self._engine = VQVAEEngine(self._cuda_state())
            self._engine_sig = self._signature()
        return self._engine

    @torch.no_grad()
    def encode(self, x):
        """[N,C,H,W] float -> int64 [N,h,w] code indices (vqvae_model.py:233-237)."""
        return self.engine().encode(x.float())

    @torch.no_grad()
    def decode(self, latents=None):
        """int64 [N,h,w] -> [N,C,H,W] (vqvae_model.py:239-242)."""
        return self.engine().decode(latents.to(torch.int64))

    def forward(self, x):
        raise NotImplementedError("stage-1 training (vqvae_model.py:244-248) is outside the sampling path (SURVEY.md §2)")


class TransformerTextEncoder(ParamTree):
    """mage_model.py:180-262 parameter layout; evaluated inside SamplerEngine._text_encoder."""

    def __init__(self, vocab_size, transformer_width, transformer_layers, output_dim, context_length, padding_idx=0, dropout=0.1):
        cfg = dict(text_encoder_config=dict(params=dict(vocab_size=vocab_size, transformer_width=transformer_width,
                                                        transformer_layers=transformer_layers, output_dim=output_dim,
                                                        context_length=context_length)))
        super().__init__(_strip(_text_spec(cfg), "text_encoder."), seed=21)
        self.padding_idx = padding_idx


class MAEncoder(ParamTree):
    """mage_model.py:104-117 parameter layout.  `ln_qkv` (additive, default False = the shipped line 92) selects TransformerBlock
    line 93 -- `x = q + attn(ln_q(q), ln_kv(k), ln_kv(v))` -- which the reference asks MAGE+ users to enable by editing the
    source (mage_model.py:92-93); here it is a config switch (`ma_config.params.ln_qkv: true` in config/mage+_*.yaml)."""

    def __init__(self, layers, d_model, dropout=0.1, ln_qkv=False):
        super().__init__(_strip(_ma_spec(layers, d_model), "ma_encoder."), seed=22)
        self.layers, self.d_model, self.ln_qkv = layers, d_model, bool(ln_qkv)


class FlatAxialDecoder(ParamTree):
    """mage_model.py:317-390 parameter layout: `out` = Linear(model_channels, K) for use_cids=True, GroupNorm(32) -> SiLU ->
    1x1x1 Conv3d(model_channels, out_channels) (keys out.0.*, out.2.*) for the MAGE+ continuous head (:349-354)."""

    def __init__(self, in_channels, model_channels, out_channels, frames_length, layers, context_channels=None,
                 use_cids=True, dropout=0.1):
        super().__init__(_strip(_decoder_spec(in_channels, model_channels, out_channels, frames_length, layers,
                                              context_channels or in_channels, use_cids), "generate_model."), seed=23)
        self.frames_length, self.layers, self.use_cids = frames_length, layers, use_cids


def _full_spec_subset(params_like: dict, prefix: str) -> syn.Spec:
    return [e for e in syn.mage_param_spec(params_like) if e[0].startswith(prefix)]


def _template(**over) -> dict:
    p = syn.model_params("caterv2")
    p.update(over)
    return p


def _text_spec(cfg) -> syn.Spec:
    p = _template()
    p["text_encoder_config"]["params"].update(cfg["text_encoder_config"]["params"])
    return _full_spec_subset(p, "text_encoder.")


def _ma_spec(layers, d_model) -> syn.Spec:
    p = _template()
    p["ma_config"]["params"].update(layers=layers, d_model=d_model)
    return _full_spec_subset(p, "ma_encoder.")


def _decoder_spec(in_channels, model_channels, out_channels, frames_length, layers, context_channels, use_cids=True) -> syn.Spec:
    p = _template() if use_cids else syn.model_params("caterv2plus")
    p["ma_config"]["params"].update(d_model=context_channels)
    p["generate_decoder_config"]["params"].update(in_channels=in_channels, model_channels=model_channels,
                                                  out_channels=out_channels, frames_length=frames_length, layers=layers)
    return _full_spec_subset(p, "generate_model.")


class PIDControl:
    """mage_model.py:394-434: the position-form PI controller that sets the KL weight beta from the measured KL (auto_beta).
    State: the integral term I and the last output / error."""

    def __init__(self):
        self.I_k1 = 0.0
        self.W_k1 = 0.0
        self.e_k1 = 0.0

    @staticmethod
    def _Kp_fun(err, scale=1):
        return 1.0 / (1.0 + float(scale) * math.exp(err))

    def pid(self, exp_KL, KL_loss, Kp=0.01, Ki=-0.0001, Kd=0.0):
        """Returns (beta clipped to [0, 1], error).  The anti-windup test of the reference (`W < 0 and W >= 1`, :419) can never be
        true, so the integral always advances -- kept that way."""
        err = exp_KL - KL_loss
        P = Kp * self._Kp_fun(err)
        I = self.I_k1 + Ki * err
        W = P + I
        self.W_k1, self.I_k1, self.e_k1 = W, I, err
        return min(max(W, 0.0), 1.0), err


class MAGE(_EngineOwner):
    """modules.mage_model.MAGE (mage_model.py:446-693), sampling path only."""

    def __init__(self, first_stage_config, text_encoder_config, ma_config, generate_decoder_config, codebook_size: int,
                 frames_length: int, image_resolution: int, vision_width: int, dropout: float = 0.1, use_cids=False,
                 randomness=False, alpha=0., beta=1., v_kl=0., auto_beta=False, with_posterior: bool = False):
        super().__init__()
        # objective weights (mage_model.py:506-511) -- read by forward() only
        self.alpha, self.beta, self.auto_beta, self.KL_loss = alpha, beta, auto_beta, v_kl
        self.PID = PIDControl() if (randomness and auto_beta) else None
        # with_posterior (additive): also hold the train-only video posterior (conv3d.*, conv_mu2, conv_var2; 85 M parameters,
        # mage_model.py:496-503) that MAGE.forward evaluates.  Off by default: sampling never touches it.
        self.with_posterior = bool(with_posterior and randomness)
        self.frames_length, self.image_resolution, self.vision_width = frames_length, image_resolution, vision_width
        self.dropout, self.use_cids, self.randomness, self.codebook_size = dropout, use_cids, randomness, codebook_size
        try:
            self.first_stage_model = instantiate_from_config(first_stage_config).eval()
        except ImportError as e:
            # MAGE+ yamls name latent-diffusion's AutoencoderKL (requirements.txt:22, un-vendored): the first stage is a pluggable
            # torch module taken as given -- any class with .embed_dim, .encode(x) -> Tensor | object with .sample(), .decode(z)
            raise ImportError(f"cannot import the first stage {first_stage_config.get('target')!r} ({e}); for use_cids=False install "
                              "it or point first_stage_config.target at a module with encode()/decode()/embed_dim") from e
        for prm in self.first_stage_model.parameters():
            prm.requires_grad = False
        self.text_encoder = instantiate_from_config(text_encoder_config)
        self.ma_encoder = instantiate_from_config(ma_config, {"dropout": dropout})
        self.generate_model = instantiate_from_config(
            generate_decoder_config, {"use_cids": use_cids, "dropout": dropout, "context_channels": ma_config["params"]["d_model"]})
        d, R = vision_width, image_resolution
        if use_cids:
            emb: syn.Spec = [("visual_token_embedding.weight", (codebook_size, d), "normal:0.02")]
        else:   # nn.Linear(first_stage_model.embed_dim, vision_width), mage_model.py:482-483
            emb = [("visual_token_embedding.weight", (d, self.first_stage_model.embed_dim), "normal:0.02"),
                   ("visual_token_embedding.bias", (d,), "bias")]
        top: syn.Spec = emb + [("conv.0.weight", (d, d, 3, 3), "conv"),
                         ("speed_embedding", (1, d), f"normal:{d ** -0.5}"),
                         ("H_positional_embedding", (1, R, 1, d), f"normal:{d ** -0.5}"),
                         ("W_positional_embedding", (1, 1, R, d), f"normal:{d ** -0.5}")]
        if randomness:
            top.append(("conv_d2.weight", (d, 64, 3, 3), "conv"))
            for br in ("conv_mu", "conv_var"):
                for i in (0, 1):
                    top.append((f"adain.{br}.{i}.weight", (d, d, 3, 3), "conv"))
                    top.append((f"adain.{br}.{i}.bias", (d,), "bias"))
        if self.with_posterior:
            params_like = dict(vision_width=d, ma_config=ma_config)
            top = top + syn.posterior_param_spec(params_like)
        tree = ParamTree(top, seed=24)
        for name, child in list(tree._modules.items()):
            self.add_module(name, child)
        for name, prm in list(tree._parameters.items()):
            self.register_parameter(name, prm)
        self.last_tokens = None
        self.last_tok0 = None
        self.range_fallbacks = 0   # calls repeated on the fp32 kernels because an activation left the tensor-core operand range

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts released checkpoints: the train-only tensors (3-D conv posterior etc.) are dropped."""
        sd = {k: v for k, v in state_dict.items() if self.with_posterior or not k.startswith(_TRAIN_ONLY_PREFIXES)}
        return super().load_state_dict(sd, strict=strict, **kw)

    def engine(self):
        if not self._engine_valid():
            from .engine import SamplerEngine
            self._engine = SamplerEngine(self._cuda_state(), self.frames_length, self.randomness,
                                         padding_idx=getattr(self.text_encoder, "padding_idx", 0), use_cids=self.use_cids,
                                         ma_ln=getattr(self.ma_encoder, "ln_qkv", False))
            self._engine_sig = self._signature()
        return self._engine

    def _wide_range_engine(self):
        """The same sampler on the fp32 SIMT kernels (no fp16 operand format, hence no range limit): built on first need."""
        if getattr(self, "_wide_engine", None) is None or self._wide_engine_sig != self._signature():
            from .engine import SamplerEngine
            self._wide_engine = SamplerEngine(self._cuda_state(), self.frames_length, self.randomness,
                                              padding_idx=getattr(self.text_encoder, "padding_idx", 0), use_cids=self.use_cids,
                                              ma_ln=getattr(self.ma_encoder, "ln_qkv", False), backend="simt")
            self._wide_engine_sig = self._signature()
        return self._wide_engine

    def get_first_stage_encoding(self, encoder_posterior):
        """mage_model.py:542-549: a posterior object is sampled, a tensor is taken as is."""
        if isinstance(encoder_posterior, torch.Tensor):
            return encoder_posterior
        if hasattr(encoder_posterior, "sample"):
            return encoder_posterior.sample()
        raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")

    @torch.no_grad()
    def first_stage_encode(self, x):
        """mage_model.py:530-540: [B,T,C,H,W] -> [B,T,h,w] int64 (use_cids) / [B,T,c,h,w] latents."""
        out = self.get_first_stage_encoding(self.first_stage_model.encode(x.reshape(-1, *x.shape[-3:])))
        return out.view(*x.shape[:-3], *out.shape[1:]).contiguous().detach()

    @torch.no_grad()
    def first_stage_decode(self, x):
        """mage_model.py:551-567: [B,T,h,w] (use_cids) / [B,T,c,h,w] -> [B,T,C,H,W]."""
        out = self.first_stage_model.decode(x.reshape(-1, *x.shape[(-2 if self.use_cids else -3):]))
        return out.view(*x.shape[:2], *out.shape[1:]).contiguous().detach()

    @torch.no_grad()
    def autoregressive_generate(self, batch, noise: Optional[torch.Tensor] = None, to_host: bool = False):
        """mage_model.py:641-693.  batch: 'images' [B,>=1,C,H,W] (frame 0 read), 'text' i64 [B,T], optional
        'speed' [B].  Returns [B, frames_length, C, H, W]; frame 0 is the input frame.  With
        randomness=True the N(0,1) noise [B,64,h,w] is drawn like the reference does -- on the CPU
        default generator -- unless passed explicitly.  to_host=True (additive): the clip is returned as a pinned host
        tensor whose frames were copied out while later frames were still being generated; it is valid until the next call."""
        eng = self.engine()
        dev = eng.device
        # token ids must index the vocabulary table (nn.Embedding raises, mage_model.py:228): checked for free on a host tensor
        # here, and for a device tensor inside text_embed_kernel (flagged, never dereferenced; raised after the call)
        if not batch["text"].is_cuda:
            vocab = self.text_encoder.state_dict()["token_embedding.weight"].shape[0]
            if batch["text"].numel() and (int(batch["text"].max()) >= vocab or int(batch["text"].min()) < 0):
                raise IndexError(f"caption token id outside the vocabulary [0, {vocab})")
        images0 = batch["images"][:, 0].to(dev, non_blocking=True)
        text = batch["text"].to(dev, non_blocking=True)
        speed = batch["speed"].to(dev, non_blocking=True).float() if "speed" in batch else None
        z0 = None
        if not self.use_cids:
            # the posterior is sampled BEFORE the AdaIN noise is drawn, like the reference (mage_model.py:642 then :661)
            z0 = self.first_stage_encode(images0.unsqueeze(1))[:, 0].float().contiguous()
        if self.randomness:
            if noise is None:
                noise = torch.randn([text.shape[0], 64, self.image_resolution, self.image_resolution])
            noise = noise.to(dev, non_blocking=True).float().contiguous()
        if not self.use_cids:
            # MAGE+ (mage_model.py:646,684,689): continuous latents between the two first-stage calls; no tokens, no argmax
            latents = eng.generate_continuous(z0, text, speed, noise)
            self.last_latents, self.last_tokens, self.last_tok0 = latents, None, None
            video = torch.cat([images0.unsqueeze(1).float(), self.first_stage_decode(latents).float()], 1)
            if to_host:
                host = torch.empty(video.shape, dtype=torch.float32, pin_memory=True)
                host.copy_(video)
                return host
            return video
        try:
            video, tokens, tok0 = eng.generate(images0, text, speed, noise, to_host=to_host)
        except _lib.MageSplitRangeError:
            # An activation left the fp16 hi/lo operand range of the tensor-core kernels (|x| > 65504 -- e.g. a residual stream
            # far outside anything the shipped checkpoints produce -- or NaN).  The reference computes such a call in plain fp32
            # (NaN in, NaN out), so the call is REPEATED on the fp32 SIMT kernels of the same library: ~9x slower, same
            # algorithm, never a wrong clip.  MAGE_RANGE_FALLBACK=0 re-raises instead.
            if os.environ.get("MAGE_RANGE_FALLBACK", "1") == "0" or eng.backend != "tc":
                raise
            warnings.warn("an activation left the tensor-core operand range (|x| > 65504 or NaN): this call is repeated on the "
                          "fp32 SIMT kernels (about 9x slower)", RuntimeWarning, stacklevel=2)
            self.range_fallbacks += 1
            video, tokens, tok0 = self._wide_range_engine().generate(images0, text, speed, noise, to_host=to_host)
        self.last_tokens, self.last_tok0 = tokens, tok0
        if to_host:
            return video
        # the engine's buffers are reused by the next call (CUDA graph); the reference hands out a fresh CONTIGUOUS
        # [B,L,C,H,W] tensor (torch.cat, mage_model.py:691) that its caller clamps in place and views (main_mage.py:242)
        return video.clone(memory_format=torch.contiguous_format)

    @torch.no_grad()
    def teacher_forced_tokens(self, batch, force_tokens: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """Per-step greedy predictions when every step is fed the GIVEN previous tokens (`force_tokens` int64 [B,L-1,h,w], e.g. the
        reference's) instead of its own: what `prediction[:, j]` of the reference's last iteration holds (mage_model.py:686-687).
        Returns (tokens int64 [B,L-1,h,w], logits f32 [B,L-1,h*w,K]).  Runs eagerly (no CUDA graph)."""
        eng = self.engine()
        dev = eng.device
        noise = noise.to(dev).float().contiguous() if (self.randomness and noise is not None) else None
        if self.randomness and noise is None:
            noise = torch.randn([batch["text"].shape[0], 64, self.image_resolution, self.image_resolution]).to(dev)
        trace = {"force_tokens": force_tokens.to(dev, torch.int64)}
        _, tokens, _ = eng.generate(batch["images"][:, 0].to(dev), batch["text"].to(dev),
                                    batch["speed"].to(dev).float() if "speed" in batch else None, noise, trace=trace)
        B = tokens.shape[0]
        logits = torch.stack([l.view(B, -1, l.shape[-1]) for l in trace["logits"]], 1)
        return tokens, logits

    @torch.no_grad()
    def forward(self, batch, test_flag=False, eps: Optional[torch.Tensor] = None):
        """mage_model.py:575-639, the FORWARD half of the stage-2 objective in eval mode -- what the reference's periodic
        validation computes (main_mage.py:163-176): (final_loss 0-dim tensor, loss_dict with 'val/prediction', 'val/kl_loss',
        ['val/beta',] 'val/final_loss').  batch: 'images' [B,frames_length,C,H,W], 'text', 'speed'.  `eps` (additive) [B,64,h,w]
        stands for the pass's one random draw, torch.randn_like in reparameterize (:571; with test_flag the draw that replaces
        the posterior sample, :610); None draws it on the device like the reference.  Gradients, dropout and the optimiser step
        are not built: in training mode this raises."""
        if self.training:
            raise NotImplementedError("MAGE.forward is built for eval mode (validation loss); training mode -- dropout, gradients, the "
                                      "optimiser step of main_mage.py:127-152 -- is not part of this library: call .eval()")
        eng = self.engine()
        dev = eng.device
        images = batch["images"].to(dev, non_blocking=True)
        text = batch["text"].to(dev, non_blocking=True)
        speed = batch["speed"].to(dev, non_blocking=True).float() if "speed" in batch else None
        B = text.shape[0]
        if self.randomness:
            if not self.with_posterior:
                raise RuntimeError("MAGE.forward with randomness=True evaluates the video posterior (conv3d.*, conv_mu2, conv_var2): build "
                                   "the model with with_posterior=True (config params) and load a checkpoint that has those tensors")
            if eps is None:
                eps = torch.randn(B, 64, self.image_resolution, self.image_resolution, device=dev)
            eps = eps.to(dev).float().contiguous()
        if self.use_cids:
            out = eng.forward_loss(images, text, speed, eps if self.randomness else None, bool(test_flag))
            self.last_tokens_all = out["tokens"]
        else:   # MAGE+: the first stage (a plain torch module, run as given) encodes every frame, mage_model.py:579
            latents = self.first_stage_encode(images).float().contiguous()
            out = eng.forward_loss_continuous(latents, text, speed, eps if self.randomness else None, bool(test_flag))
        prefix = "val"   # `'train' if self.training else 'val'` (:601); training mode is refused above
        recon = out["prediction"].reshape(())
        loss_dict = {f"{prefix}/prediction": recon.item()}
        if self.randomness:
            kl = out["kl_loss"].reshape(())
            loss_dict[f"{prefix}/kl_loss"] = kl.item()
            if self.auto_beta:
                self.beta, _ = self.PID.pid(self.KL_loss, kl.item())
                loss_dict[f"{prefix}/beta"] = self.beta
                final = recon + self.beta * kl
            else:
                if speed is None:   # the reference reads speed_emb here, which exists only with batch['speed'] (:612-614, :631)
                    raise KeyError("speed")
                speed_emb = speed.view(B, 1) @ self.speed_embedding
                l2 = torch.mean(torch.pow(torch.norm(speed_emb, dim=-1), 2))
                final = recon + self.beta * kl + self.alpha * l2
        else:
            final = recon
        loss_dict[f"{prefix}/final_loss"] = final.item()
        return final, loss_dict
