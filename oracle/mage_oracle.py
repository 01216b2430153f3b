"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MAGE sampling path.

A functional fp32 restatement, on torch CPU ops, of what the reference computes on the
path `main_mage.py --split test` -> `MAGE.autoregressive_generate`
(/root/reference/modules/mage_model.py:641-693) and `VectorQuantizedVAE.encode/decode`
(/root/reference/modules/vqvae_model.py:233-242).  It takes a plain `state_dict` with the
reference's key names instead of nn.Modules.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference in the
authoring container (via oracle/ref_shims.py) on seeded synthetic checkpoints/batches and
commits the outputs under tests/golden/; tests/test_oracle_golden.py holds this file to
those vectors (bit-exact tokens/VQ indices, pixels to 1e-6), and, where /root/reference is
present, to the reference itself.  The reference has no tests or golden vectors of its own
(SURVEY.md §4).

Only tests/, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` / `--impl reference`
legs may import this package.  The product (mage_b200/) never does.

Two evaluation orders are provided:
  * `generate`             -- the reference's own order: every step re-runs the 3x3 conv and
                              the 6-block decoder over all L positions (mage_model.py:673-684).
  * `generate_incremental` -- App. D of SURVEY.md: one position per step with a K/V cache for
                              the two temporal blocks.  Mathematically identical because the
                              temporal blocks are causal; used for full-length parity runs.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# VQ-VAE  (vqvae_model.py)
# --------------------------------------------------------------------------------------
def _conv(sd: SD, name: str, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _bn(sd: SD, name: str, x):
    # eval-mode BatchNorm2d (first stage is frozen in eval: mage_model.py:518-519)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], training=False, eps=1e-5)


def _res_block(sd: SD, name: str, x):
    """vqvae_model.py:111-124.  The leading ReLU is in-place, so the skip adds relu(x)."""
    x = F.relu(x)
    h = _bn(sd, name + ".block.2", _conv(sd, name + ".block.1", x, padding=1))
    h = _bn(sd, name + ".block.5", _conv(sd, name + ".block.4", F.relu(h)))
    return x + h


def _enc_block(sd: SD, name: str, x):
    """vqvae_model.py:126-145 (3x3, 3x3, 3x3, 1x1; non-in-place ReLUs)."""
    idp = _conv(sd, name + ".id_path", x) if (name + ".id_path.weight") in sd else x
    h = _conv(sd, name + ".block.1", F.relu(x), padding=1)
    h = _conv(sd, name + ".block.3", F.relu(h), padding=1)
    h = _conv(sd, name + ".block.5", F.relu(h), padding=1)
    h = _conv(sd, name + ".block.7", F.relu(h))
    return idp + h


def _dec_block(sd: SD, name: str, x):
    """vqvae_model.py:147-166 (1x1, 3x3, 3x3, 3x3)."""
    idp = _conv(sd, name + ".id_path", x) if (name + ".id_path.weight") in sd else x
    h = _conv(sd, name + ".block.1", F.relu(x))
    h = _conv(sd, name + ".block.3", F.relu(h), padding=1)
    h = _conv(sd, name + ".block.5", F.relu(h), padding=1)
    h = _conv(sd, name + ".block.7", F.relu(h), padding=1)
    return idp + h


def vqvae_down_ratio(sd: SD) -> int:
    return 4 if "encoder.1.running_mean" in sd else 8


def vqvae_encoder(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """x [N,C,H,W] -> z_e [N,D,h,w].  f4: vqvae_model.py:172-179, f8: :192-202."""
    if vqvae_down_ratio(sd) == 4:
        h = F.relu(_bn(sd, "encoder.1", _conv(sd, "encoder.0", x, stride=2, padding=1)))
        h = _conv(sd, "encoder.3", h, stride=2, padding=1)
        h = _res_block(sd, "encoder.4", h)
        return _res_block(sd, "encoder.5", h)
    h = _conv(sd, "encoder.0", x, padding=3)
    h = F.max_pool2d(_enc_block(sd, "encoder.1", h), 2)
    h = F.max_pool2d(_enc_block(sd, "encoder.3", h), 2)
    h = F.max_pool2d(_enc_block(sd, "encoder.5", h), 2)
    return F.relu(_enc_block(sd, "encoder.7", h))


def vq_distances(z_flat: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """vqvae_model.py:14-19: ||c||^2 + ||z||^2 - 2 z.c^T through one addmm."""
    c_sqr = torch.sum(codebook ** 2, dim=1)
    z_sqr = torch.sum(z_flat ** 2, dim=1, keepdim=True)
    return torch.addmm(c_sqr + z_sqr, z_flat, codebook.t(), alpha=-2.0, beta=1.0)


def vq_argmin(z_flat: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """vqvae_model.py:21: indices of the row minima (int64)."""
    return torch.min(vq_distances(z_flat, codebook), dim=1)[1]


def vqvae_quantize(sd: SD, z_e: torch.Tensor) -> torch.Tensor:
    """VQEmbedding.forward (vqvae_model.py:93-96): NCHW -> NHWC -> nearest code, int64 [N,h,w]."""
    cb = sd["codebook.embedding.weight"]
    z = z_e.permute(0, 2, 3, 1).contiguous()
    return vq_argmin(z.view(-1, cb.shape[1]), cb).view(*z.shape[:-1])


def vqvae_encode(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """VectorQuantizedVAE.encode (vqvae_model.py:233-237)."""
    return vqvae_quantize(sd, vqvae_encoder(sd, x))


def vqvae_decode(sd: SD, latents: torch.Tensor) -> torch.Tensor:
    """VectorQuantizedVAE.decode (vqvae_model.py:239-242); f4 decoder :180-189, f8 :203-214."""
    z = F.embedding(latents, sd["codebook.embedding.weight"]).permute(0, 3, 1, 2)
    if vqvae_down_ratio(sd) == 4:
        h = _res_block(sd, "decoder.0", z.contiguous())
        h = F.relu(_res_block(sd, "decoder.1", h))
        h = F.conv_transpose2d(h, sd["decoder.3.weight"], sd["decoder.3.bias"], stride=2, padding=1)
        h = F.relu(_bn(sd, "decoder.4", h))
        h = F.conv_transpose2d(h, sd["decoder.6.weight"], sd["decoder.6.bias"], stride=2, padding=1)
        return torch.tanh(h)
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    h = up(_dec_block(sd, "decoder.0", z))
    h = up(_dec_block(sd, "decoder.2", h))
    h = up(_dec_block(sd, "decoder.4", h))
    h = _dec_block(sd, "decoder.6", h)
    return torch.tanh(_conv(sd, "decoder.8", F.relu(h)))


def _sub(sd: SD, prefix: str) -> SD:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------
# transformer pieces  (mage_model.py)
# --------------------------------------------------------------------------------------
def _mha(sd: SD, name: str, q, k, v, n_head: int, attn_mask=None, key_padding_mask=None):
    """nn.MultiheadAttention.forward, seq-first (batch_first=False everywhere in the
    reference: mage_model.py:20,75,193-199), need_weights=False -> SDPA."""
    return F.multi_head_attention_forward(
        q, k, v, q.shape[-1], n_head,
        sd[name + ".in_proj_weight"], sd[name + ".in_proj_bias"], None, None, False, 0.0,
        sd[name + ".out_proj.weight"], sd[name + ".out_proj.bias"],
        training=False, key_padding_mask=key_padding_mask, need_weights=False, attn_mask=attn_mask)[0]


def _ln(sd: SD, name: str, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)  # mage_model.py:11-13


def _mlp(sd: SD, name: str, x):
    h = F.linear(x, sd[name + ".c_fc.weight"], sd[name + ".c_fc.bias"])
    return F.linear(_quick_gelu(h), sd[name + ".c_proj.weight"], sd[name + ".c_proj.bias"])


def text_encoder(sd: SD, text: torch.Tensor, padding_idx: int = 0) -> torch.Tensor:
    """TransformerTextEncoder.forward (mage_model.py:223-250).  text i64 [B,T] -> [B,T,C]."""
    p = "text_encoder."
    width = sd[p + "token_embedding.weight"].shape[1]
    n_head = width // 32
    B, T = text.shape
    text_length = (text != padding_idx).float().sum(-1)
    pos = torch.arange(T, dtype=text.dtype, device=text.device).unsqueeze(0).expand(B, T)
    x = F.embedding(text, sd[p + "token_embedding.weight"], padding_idx=padding_idx)
    x = _ln(sd, p + "layer_norm", x + F.embedding(pos, sd[p + "positions.weight"]), eps=1e-8)
    x = x * (text != padding_idx).unsqueeze(-1).type(x.dtype)
    caption_mask = text_length.unsqueeze(1) < torch.ones_like(text).cumsum(dim=1)  # True = padded key
    x = x.permute(1, 0, 2)
    i = 0
    while (p + f"transformer.layers.{i}.linear1.weight") in sd:  # post-norm nn.TransformerEncoderLayer, exact GELU
        lp = p + f"transformer.layers.{i}"
        x = _ln(sd, lp + ".norm1", x + _mha(sd, lp + ".self_attn", x, x, x, n_head, key_padding_mask=caption_mask))
        h = F.linear(F.gelu(F.linear(x, sd[lp + ".linear1.weight"], sd[lp + ".linear1.bias"])),
                     sd[lp + ".linear2.weight"], sd[lp + ".linear2.bias"])
        x = _ln(sd, lp + ".norm2", x + h)
        i += 1
    x = _ln(sd, p + "ln_text_final", x.permute(1, 0, 2))
    return F.linear(x, sd[p + "text_projection.weight"], sd[p + "text_projection.bias"])


def ma_encoder(sd: SD, q: torch.Tensor, kv: torch.Tensor) -> torch.Tensor:
    """MAEncoder.forward (mage_model.py:114-117) with the shipped TransformerBlock line 92:
    no LayerNorm on q/kv and no key-padding mask.  q [HW,B,C], kv [T,B,C] seq-first."""
    i = 0
    x = q
    while f"ma_encoder.blocks.{i}.attn.in_proj_weight" in sd:
        p = f"ma_encoder.blocks.{i}"
        x = x + _mha(sd, p + ".attn", x, kv, kv, x.shape[-1] // 32)
        x = x + _mlp(sd, p + ".mlp", _ln(sd, p + ".ln_2", x))
        i += 1
    return x


def adain(sd: SD, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """ADAIN2D.forward (mage_model.py:309-314): InstanceNorm(x) * conv_mu(y) + conv_var(y)."""
    out = F.instance_norm(x, eps=1e-5)
    g = _conv(sd, "adain.conv_mu.1", _conv(sd, "adain.conv_mu.0", y, padding=1), padding=1)
    b = _conv(sd, "adain.conv_var.1", _conv(sd, "adain.conv_var.0", y, padding=1), padding=1)
    return g * out + b


def axial_block(sd: SD, name: str, x: torch.Tensor, axial_dim: int, attn_mask=None) -> torch.Tensor:
    """AxialAttentionBlock.forward (mage_model.py:35-53) on x [B,L,H,W,C]: move `axial_dim`
    next to the channel dim, flatten the rest into the MHA batch, pre-LN attention + MLP."""
    nd = x.dim()
    rest = [d for d in range(nd) if d not in (axial_dim, nd - 1)]
    perm = rest + [axial_dim, nd - 1]
    inv = [perm.index(d) for d in range(nd)]
    xp = x.permute(perm).contiguous()
    shp = xp.shape
    s = xp.view(-1, shp[-2], shp[-1]).transpose(0, 1)  # [S, N, C]
    u = _ln(sd, name + ".ln_1", s)
    s = s + _mha(sd, name + ".attn", u, u, u, shp[-1] // 32, attn_mask=attn_mask)
    s = s + _mlp(sd, name + ".mlp", _ln(sd, name + ".ln_2", s))
    return s.transpose(0, 1).reshape(shp).permute(inv).contiguous()


def flat_axial_decoder(sd: SD, motion: torch.Tensor, imgs: torch.Tensor, return_hidden: bool = False):
    """FlatAxialDecoder.forward (mage_model.py:374-390).  motion [B,H,W,C] takes temporal
    position 0, imgs [B,F,H,W,C] positions 1..F; logits come from positions 1.. only."""
    p = "generate_model."
    x = torch.cat([F.linear(motion, sd[p + "context_linear.weight"], sd[p + "context_linear.bias"]).unsqueeze(1),
                   F.linear(imgs, sd[p + "in_linear.weight"], sd[p + "in_linear.bias"])], 1)
    Lmax = sd[p + "T_positional_embedding"].shape[0]
    assert x.shape[1] == Lmax, "reference adds the full T_positional_embedding (needs F == frames_length-1)"
    x = x + sd[p + "T_positional_embedding"]
    mask = torch.full((Lmax, Lmax), float("-inf")).triu_(1).to(x.device)  # mage_model.py:367-372 (built on the CPU, moved in attention(), :32)
    i = 0
    while (p + f"blocks.{i}.ln_1.weight") in sd:
        x = axial_block(sd, p + f"blocks.{i}", x, i % 3 + 1, mask if i % 3 == 0 else None)
        i += 1
    logits = F.linear(x[:, 1:], sd[p + "out.weight"], sd[p + "out.bias"])
    return (logits, x) if return_hidden else logits


def token_features(sd: SD, x_emb: torch.Tensor) -> torch.Tensor:
    """3x3 bias-free conv over token embeddings + H/W positional embeddings
    (mage_model.py:648-649 / :674-676).  x_emb [B,F,C,H,W] -> [B,F,H,W,C]."""
    B, Fr, C, H, W = x_emb.shape
    f = F.conv2d(x_emb.reshape(-1, C, H, W), sd["conv.0.weight"], None, padding=1)
    f = f.view(B, Fr, C, H, W).permute(0, 1, 3, 4, 2).contiguous()
    return f + sd["H_positional_embedding"] + sd["W_positional_embedding"]


def embed_tokens(sd: SD, tok: torch.Tensor) -> torch.Tensor:
    """visual_token_embedding gather, [..,H,W] i64 -> [..,C,H,W] (mage_model.py:644,682)."""
    e = F.embedding(tok, sd["visual_token_embedding.weight"])
    nd = e.dim()
    return e.permute(*range(nd - 3), nd - 1, nd - 3, nd - 2).contiguous()


def motion_anchor(sd: SD, tok0: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor],
                  noise: Optional[torch.Tensor], trace: Optional[dict] = None) -> torch.Tensor:
    """Prelude of autoregressive_generate (mage_model.py:644-668).  tok0 [B,H,W] i64,
    noise [B,64,H,W] or None (randomness=False) -> anchor [B,H,W,C]."""
    B, H, W = tok0.shape
    x_emb = embed_tokens(sd, tok0).unsqueeze(1)
    C = x_emb.shape[2]
    first = token_features(sd, x_emb)[:, 0].reshape(B, -1, C).permute(1, 0, 2).contiguous()
    t = text_encoder(sd, text).permute(1, 0, 2).contiguous()
    a = ma_encoder(sd, first, t).permute(1, 0, 2).contiguous().view(B, H, W, C)
    if trace is not None:
        trace["text_emb"], trace["first_img"], trace["anchor_ma"] = t, first, a
    if noise is not None:
        y = F.conv2d(noise, sd["conv_d2.weight"], None, padding=1)
        a = adain(sd, a.permute(0, 3, 1, 2).contiguous(), y).permute(0, 2, 3, 1).contiguous()
    if speed is not None:
        a = a + (speed.view(B, 1) @ sd["speed_embedding"]).unsqueeze(1).unsqueeze(1)
    if trace is not None:
        trace["anchor"] = a
    return a


def _top2_gap(logits: torch.Tensor) -> torch.Tensor:
    t = torch.topk(logits, 2, dim=-1)[0]
    return t[..., 0] - t[..., 1]


@torch.no_grad()
def generate(sd: SD, batch: Dict[str, torch.Tensor], noise: Optional[torch.Tensor] = None,
             trace: Optional[dict] = None) -> torch.Tensor:
    """MAGE.autoregressive_generate in the reference's own evaluation order
    (mage_model.py:641-693).  Returns [B,L,C,H,W]; `trace` (if given) receives tokens
    [B,L-1,h,w], the final-step logit top1-top2 gaps and intermediates."""
    fsd = _sub(sd, "first_stage_model.")
    L = sd["generate_model.T_positional_embedding"].shape[0]
    img0 = batch["images"][:, 0]
    tok0 = vqvae_encode(fsd, img0)
    anchor = motion_anchor(sd, tok0, batch["text"], batch.get("speed"), noise, trace)
    x_emb = embed_tokens(sd, tok0).unsqueeze(1)
    inp = x_emb.repeat(1, L - 1, 1, 1, 1)
    prediction = None
    for i in range(L - 1):
        prediction = flat_axial_decoder(sd, anchor, token_features(sd, inp))
        if i != L - 2:
            ids = torch.max(prediction, -1)[1]
            inp[:, i + 1] = embed_tokens(sd, ids[:, i])
    tokens = torch.max(prediction, -1)[1]
    if trace is not None:
        trace["tok0"], trace["tokens"], trace["gap"] = tok0, tokens, _top2_gap(prediction)
    B = tokens.shape[0]
    pix = vqvae_decode(fsd, tokens.view(-1, *tokens.shape[-2:]))
    pix = pix.view(B, L - 1, *pix.shape[1:])
    return torch.cat([batch["images"][:, 0:1], pix], 1)


# --------------------------------------------------------------------------------------
# incremental evaluation order
# --------------------------------------------------------------------------------------
def _heads(t: torch.Tensor, n_head: int) -> torch.Tensor:
    return t.view(*t.shape[:-1], n_head, t.shape[-1] // n_head)


def _block_step(sd: SD, name: str, x: torch.Tensor, kind: int, cache: Optional[list]) -> torch.Tensor:
    """One axial block on ONE temporal position.  x [B,H,W,C]; kind 0 = temporal (attend
    the cache of positions 0..p), 1 = along H, 2 = along W."""
    C = x.shape[-1]
    nh = C // 32
    u = _ln(sd, name + ".ln_1", x)
    qkv = F.linear(u, sd[name + ".attn.in_proj_weight"], sd[name + ".attn.in_proj_bias"])
    q, k, v = [_heads(t, nh) for t in qkv.split(C, dim=-1)]  # [B,H,W,nh,32]
    if kind == 0:
        cache.append((k, v))
        K = torch.stack([c[0] for c in cache], dim=-2)  # [B,H,W,nh,S,32]
        V = torch.stack([c[1] for c in cache], dim=-2)
        a = F.scaled_dot_product_attention(q.unsqueeze(-2), K, V).squeeze(-2)
    else:
        ax = 1 if kind == 1 else 2  # attended spatial dim of [B,H,W,nh,32]
        mv = lambda t: t.movedim(ax, -2)  # [B,other,nh,S,32]
        a = F.scaled_dot_product_attention(mv(q), mv(k), mv(v)).movedim(-2, ax)
    a = a.reshape(*x.shape)
    x = x + F.linear(a, sd[name + ".attn.out_proj.weight"], sd[name + ".attn.out_proj.bias"])
    return x + _mlp(sd, name + ".mlp", _ln(sd, name + ".ln_2", x))


@torch.no_grad()
def generate_incremental(sd: SD, batch: Dict[str, torch.Tensor], noise: Optional[torch.Tensor] = None,
                         trace: Optional[dict] = None, frames_length: Optional[int] = None) -> torch.Tensor:
    """Same result as `generate`, O(L) work: per step only the newest temporal position is
    evaluated; temporal blocks (i % 3 == 0) keep K/V of earlier positions."""
    p = "generate_model."
    fsd = _sub(sd, "first_stage_model.")
    Tp = sd[p + "T_positional_embedding"]
    L = frames_length or Tp.shape[0]
    n_blocks = 0
    while (p + f"blocks.{n_blocks}.ln_1.weight") in sd:
        n_blocks += 1
    tok0 = vqvae_encode(fsd, batch["images"][:, 0])
    anchor = motion_anchor(sd, tok0, batch["text"], batch.get("speed"), noise, trace)
    caches = [[] for _ in range(n_blocks)]

    def step(pos: int, x: torch.Tensor) -> torch.Tensor:
        x = x + Tp[pos]
        for i in range(n_blocks):
            x = _block_step(sd, p + f"blocks.{i}", x, i % 3, caches[i])
        return x

    step(0, F.linear(anchor, sd[p + "context_linear.weight"], sd[p + "context_linear.bias"]))
    tok = tok0
    toks, gaps = [], []
    for j in range(L - 1):
        f = token_features(sd, embed_tokens(sd, tok).unsqueeze(1))[:, 0]
        h = step(j + 1, F.linear(f, sd[p + "in_linear.weight"], sd[p + "in_linear.bias"]))
        logits = F.linear(h, sd[p + "out.weight"], sd[p + "out.bias"])
        tok = torch.max(logits, -1)[1]
        toks.append(tok)
        gaps.append(_top2_gap(logits))
        if trace is not None and trace.get("want_logits"):
            trace.setdefault("logits", []).append(logits.reshape(logits.shape[0], -1, logits.shape[-1]))
    tokens = torch.stack(toks, 1)
    if trace is not None:
        trace["tok0"], trace["tokens"], trace["gap"] = tok0, tokens, torch.stack(gaps, 1)
        if trace.get("want_logits"):
            trace["logits"] = torch.stack(trace["logits"], 1)   # [B, L-1, h*w, K]
    B = tokens.shape[0]
    pix = vqvae_decode(fsd, tokens.view(-1, *tokens.shape[-2:]))
    pix = pix.view(B, L - 1, *pix.shape[1:])
    return torch.cat([batch["images"][:, 0:1], pix], 1)


# --------------------------------------------------------------------------------------
# MAGE+ branch: use_cids=False (continuous latents, no codebook / argmax)
#   mage_model.py:482-483 (Linear embed), :349-354 + :386-388 (GroupNorm -> SiLU -> 1x1x1 Conv3d head),
#   :646,:684,:689 (continuous autoregression), :93 (ln_q / ln_kv motion anchor -- the documented manual edit).
# The first stage (AutoencoderKL of un-vendored latent-diffusion@main in the shipped yaml) is a pluggable module whose
# encode()/decode() are taken as given: PARITY UNPINNED for that module, pinned for everything below it.
# --------------------------------------------------------------------------------------
def embed_latents(sd: SD, z: torch.Tensor) -> torch.Tensor:
    """visual_token_embedding = Linear(embed_dim, C) on channels-last latents (mage_model.py:482-483, :646).
    z [..., c, H, W] -> [..., C, H, W]."""
    nd = z.dim()
    zl = z.permute(*range(nd - 3), nd - 2, nd - 1, nd - 3)
    e = F.linear(zl, sd["visual_token_embedding.weight"], sd["visual_token_embedding.bias"])
    return e.permute(*range(nd - 3), nd - 1, nd - 3, nd - 2).contiguous()


def ma_encoder_ln(sd: SD, q: torch.Tensor, kv: torch.Tensor) -> torch.Tensor:
    """MAEncoder with TransformerBlock line 93 (the MAGE+ edit): x = q + attn(ln_q(q), ln_kv(kv), ln_kv(kv), key_mask=None)."""
    i = 0
    x = q
    while f"ma_encoder.blocks.{i}.attn.in_proj_weight" in sd:
        p = f"ma_encoder.blocks.{i}"
        k = _ln(sd, p + ".ln_kv", kv)
        x = x + _mha(sd, p + ".attn", _ln(sd, p + ".ln_q", x), k, k, x.shape[-1] // 32)
        x = x + _mlp(sd, p + ".mlp", _ln(sd, p + ".ln_2", x))
        i += 1
    return x


def continuous_head(sd: SD, hidden: torch.Tensor) -> torch.Tensor:
    """FlatAxialDecoder.out for use_cids=False (mage_model.py:349-354, :386-388): hidden [B,F,H,W,C] -> GroupNorm(32) over
    (C/32 channels x F x H x W) per sample -> SiLU -> 1x1x1 Conv3d -> [B,F,H,W,c_out].  The statistics span ALL F slots."""
    p = "generate_model.out."
    x = hidden.permute(0, 4, 1, 2, 3).contiguous()
    x = F.group_norm(x, 32, sd[p + "0.weight"], sd[p + "0.bias"], eps=1e-5)
    x = F.conv3d(F.silu(x), sd[p + "2.weight"], sd[p + "2.bias"])
    return x.permute(0, 2, 3, 4, 1).contiguous()


def flat_axial_decoder_continuous(sd: SD, motion: torch.Tensor, imgs: torch.Tensor) -> torch.Tensor:
    """FlatAxialDecoder.forward, use_cids=False: same blocks as `flat_axial_decoder`, continuous head."""
    p = "generate_model."
    x = torch.cat([F.linear(motion, sd[p + "context_linear.weight"], sd[p + "context_linear.bias"]).unsqueeze(1),
                   F.linear(imgs, sd[p + "in_linear.weight"], sd[p + "in_linear.bias"])], 1)
    Lmax = sd[p + "T_positional_embedding"].shape[0]
    assert x.shape[1] == Lmax
    x = x + sd[p + "T_positional_embedding"]
    mask = torch.full((Lmax, Lmax), float("-inf")).triu_(1).to(x.device)
    i = 0
    while (p + f"blocks.{i}.ln_1.weight") in sd:
        x = axial_block(sd, p + f"blocks.{i}", x, i % 3 + 1, mask if i % 3 == 0 else None)
        i += 1
    return continuous_head(sd, x[:, 1:])


@torch.no_grad()
def generate_continuous(sd: SD, z0: torch.Tensor, text: torch.Tensor, speed: Optional[torch.Tensor],
                        noise: Optional[torch.Tensor] = None, ma_ln: bool = False, trace: Optional[dict] = None) -> torch.Tensor:
    """MAGE.autoregressive_generate for use_cids=False between the two first-stage calls (mage_model.py:642-689), reference
    evaluation order.  z0 [B,c,H,W] = first_stage_encode(images[:, 0:1])[:, 0]; returns the predicted latents [B,L-1,c,H,W]
    (what the reference hands to first_stage_decode, :689-690)."""
    L = sd["generate_model.T_positional_embedding"].shape[0]
    B, c, H, W = z0.shape
    x_emb = embed_latents(sd, z0.unsqueeze(1))                                   # [B,1,C,H,W]
    C = x_emb.shape[2]
    first = token_features(sd, x_emb)[:, 0].reshape(B, -1, C).permute(1, 0, 2).contiguous()
    t = text_encoder(sd, text).permute(1, 0, 2).contiguous()
    a = (ma_encoder_ln if ma_ln else ma_encoder)(sd, first, t).permute(1, 0, 2).contiguous().view(B, H, W, C)
    if noise is not None:
        y = F.conv2d(noise, sd["conv_d2.weight"], None, padding=1)
        a = adain(sd, a.permute(0, 3, 1, 2).contiguous(), y).permute(0, 2, 3, 1).contiguous()
    if speed is not None:
        a = a + (speed.view(B, 1) @ sd["speed_embedding"]).unsqueeze(1).unsqueeze(1)
    if trace is not None:
        trace["anchor"] = a
    inp = x_emb.repeat(1, L - 1, 1, 1, 1)
    pred = None
    for i in range(L - 1):
        pred = flat_axial_decoder_continuous(sd, a, token_features(sd, inp))        # [B,L-1,H,W,c]
        if i != L - 2:
            if trace is not None:   # what iteration i feeds into slot i+1 (differs from the FINAL prediction of slot i: the head's
                trace.setdefault("step_pred", []).append(pred[:, i].permute(0, 3, 1, 2).clone())   # GroupNorm spans all slots)
            inp[:, i + 1] = embed_latents(sd, pred.permute(0, 1, 4, 2, 3))[:, i]
    if trace is not None and "step_pred" in trace:
        trace["step_pred"] = torch.stack(trace["step_pred"], 1)                     # [B, L-2, c, H, W]
    return pred.permute(0, 1, 4, 2, 3).contiguous()


# --------------------------------------------------------------------------------------
# stage-2 objective, forward half: MAGE.forward in eval mode
# --------------------------------------------------------------------------------------
def _gn(sd: SD, name: str, x: torch.Tensor, groups: int = 16) -> torch.Tensor:
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps=1e-5)


def basic_block_3d(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """BasicBlock.forward (mage_model.py:264-297) as MAGE builds it (:497-500: stride 1, stride_t 2, downsample=True):
    Conv3d 3x3x3 (temporal stride 2, no bias) -> GroupNorm(16) -> ReLU -> Conv3d 3x3x3 -> GroupNorm(16); the residual goes
    through its own strided Conv3d + GroupNorm; sum -> ReLU.  x [B,C,T,H,W]."""
    out = F.conv3d(x, sd[name + ".conv1.weight"], None, stride=(2, 1, 1), padding=1)
    out = F.relu(_gn(sd, name + ".bn1", out))
    out = _gn(sd, name + ".bn2", F.conv3d(out, sd[name + ".conv2.weight"], None, stride=1, padding=1))
    res = _gn(sd, name + ".downsample.1", F.conv3d(x, sd[name + ".downsample.0.weight"], None, stride=(2, 1, 1), padding=1))
    return F.relu(out + res)


def video_posterior(sd: SD, x_emb: torch.Tensor):
    """mage_model.py:605-607 + reparameterize (:569-573) without the draw: x_emb [B,L,C,H,W] (raw token embeddings of ALL L
    frames) -> four BasicBlocks over [B,C,L,H,W] (each halves T; T must end at 1, `.squeeze(2)`) -> (mu, logvar) [B,64,H,W]."""
    v = x_emb.permute(0, 2, 1, 3, 4).contiguous()
    for i in range(4):
        v = basic_block_3d(sd, f"conv3d.{i}", v)
    assert v.shape[2] == 1, "the reference squeezes the temporal axis: frames_length must reduce to 1 in four halvings"
    v = v.squeeze(2)
    return _conv(sd, "conv_mu2", v, padding=1), _conv(sd, "conv_var2", v, padding=1)


@torch.no_grad()
def forward_loss(sd: SD, batch: Dict[str, torch.Tensor], eps: Optional[torch.Tensor], *, randomness: bool = True,
                 beta: float = 1.0, alpha: float = 0.0, test_flag: bool = False, trace: Optional[dict] = None,
                 latents: Optional[torch.Tensor] = None, ma_ln: bool = False) -> Dict[str, float]:
    """MAGE.forward (mage_model.py:575-639) in eval mode (dropout off) with a GIVEN beta: the teacher-forced full-sequence pass
    and its losses.  batch['images'] [B,L,C,H,W] (all L frames), 'text', optional 'speed'; `eps` [B,64,H,W] stands for the
    `torch.randn_like(logvar)` draw of reparameterize (:571) -- or, with test_flag, for the `torch.randn_like(video_emb)` that
    replaces the posterior sample (:609-610).  use_cids=True (default): VQ tokens, cross-entropy (:619).  `latents` [B,L,c,h,w]
    given = the MAGE+ branch (use_cids=False): the first stage's latents of all frames (whatever module produced them), Linear
    embed, continuous head, MSE (:621); `ma_ln` = TransformerBlock line 93.  With auto_beta the caller obtains beta from PIDControl
    and the L2 term is absent (:627-630): pass alpha=0.  Returns the reference's loss_dict values (without the prefix)."""
    imgs = batch["images"]
    B, L = imgs.shape[:2]
    if latents is None:
        fsd = _sub(sd, "first_stage_model.")
        tok = vqvae_encode(fsd, imgs.reshape(-1, *imgs.shape[2:]))
        tok = tok.view(B, L, *tok.shape[1:])
        x_emb = embed_tokens(sd, tok)                                      # [B,L,C,H,W]
    else:
        tok = None
        x_emb = embed_latents(sd, latents)
    prior = token_features(sd, x_emb[:, :L - 1])                           # [B,L-1,H,W,C]
    C = x_emb.shape[2]
    first = prior[:, 0].reshape(B, -1, C).permute(1, 0, 2).contiguous()
    t = text_encoder(sd, batch["text"]).permute(1, 0, 2).contiguous()
    H, W = x_emb.shape[-2:]
    a = (ma_encoder_ln if ma_ln else ma_encoder)(sd, first, t).permute(1, 0, 2).contiguous().view(B, H, W, C)
    out: Dict[str, float] = {}
    kl = None
    if randomness:
        mu, logvar = video_posterior(sd, x_emb)
        video_emb = eps * torch.exp(0.5 * logvar) + mu
        if test_flag:
            video_emb = eps
        y = F.conv2d(video_emb, sd["conv_d2.weight"], None, padding=1)
        a = adain(sd, a.permute(0, 3, 1, 2).contiguous(), y).permute(0, 2, 3, 1).contiguous()
        m2, lv2 = mu.reshape(B, -1), logvar.reshape(B, -1)
        kl = -0.5 * torch.mean(torch.sum(1 + lv2 - m2.pow(2) - lv2.exp(), dim=1))
        if trace is not None:
            trace["mu"], trace["logvar"] = mu, logvar
    speed_emb = None
    if batch.get("speed") is not None:
        speed_emb = batch["speed"].view(B, 1) @ sd["speed_embedding"]
        a = a + speed_emb.unsqueeze(1).unsqueeze(1)
    if latents is None:
        logits = flat_axial_decoder(sd, a, prior)                           # [B,L-1,H,W,K]
        K = logits.shape[-1]
        pred = F.cross_entropy(logits.reshape(-1, K), tok[:, 1:L].reshape(-1))
        if trace is not None:
            trace["tokens"], trace["logits"] = tok, logits
    else:
        model_predict = flat_axial_decoder_continuous(sd, a, prior)         # [B,L-1,H,W,c]
        pred = F.mse_loss(model_predict.permute(0, 1, 4, 2, 3).contiguous(), latents[:, 1:])
        if trace is not None:
            trace["prediction"] = model_predict
    out["prediction"] = float(pred)
    final = pred
    if randomness:
        out["kl_loss"] = float(kl)
        # mage_model.py:631-632 (auto_beta=False); the L2 term needs batch['speed'] in the reference too (speed_emb, :613)
        l2 = torch.mean(torch.pow(torch.norm(speed_emb, dim=-1), 2)) if (speed_emb is not None and alpha != 0.0) else torch.zeros(())
        final = pred + beta * kl + alpha * l2
    out["final_loss"] = float(final)
    return out
