"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shims.py) on seeded synthetic inputs.

Run in the authoring container (the GPU box has no /root/reference):

    python -m oracle.make_golden codebooks   # conditioned codebook fixtures (uses the oracle encoder)
    python -m oracle.make_golden goldens     # reference outputs -> tests/golden/*.npz

Every case is fully determined by (family, frames_length, batch, seeds) recorded inside the
.npz, so tests can rebuild the checkpoint/batch with mage_b200.synthetic and compare.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from mage_b200 import synthetic as syn
from oracle import mage_oracle as orc
from oracle import ref_shims

GOLDEN_DIR = syn.GOLDEN_DIR

# (name, family, frames_length, batch, text_len, padded, with_speed, noise_seed)
CASES = [
    ("cater_L4_b2", "caterv2", 4, 2, 12, False, True, 99),
    ("cater_L4_b2_pad", "caterv2", 4, 2, 14, True, True, 98),
    ("caterv1_L3_b1_norand", "caterv1", 3, 1, 10, False, False, None),
    ("mnist_L5_b2", "mnist", 5, 2, 9, False, True, None),
    ("cater_L10_b1", "caterv2", 10, 1, 20, False, True, 97),
]


def farthest_point_sample(x: torch.Tensor, k: int) -> torch.Tensor:
    """Greedy farthest-point sampling in fp64 (SURVEY.md H1 recipe)."""
    x = x.double()
    n = x.shape[0]
    chosen = [int(torch.argmax((x - x.mean(0)).pow(2).sum(1)))]
    d = (x - x[chosen[0]]).pow(2).sum(1)
    for _ in range(k - 1):
        i = int(torch.argmax(d))
        chosen.append(i)
        d = torch.minimum(d, (x - x[i]).pow(2).sum(1))
    return torch.tensor(chosen)


def make_codebooks():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for family in ("caterv2", "mnist"):
        params = syn.model_params(family)
        fs = params["first_stage_config"]["params"]
        sd = syn.make_vqvae_state_dict(fs, conditioned=False)
        C, res = fs["input_dim"], 16 * fs["down_ratio"]
        lo, hi = (-0.5, 0.5) if fs["down_ratio"] == 4 else (-1.0, 1.0)
        imgs = syn.structured_images(16, C, res, seed=555, lo=lo, hi=hi)
        with torch.no_grad():
            z = orc.vqvae_encoder(sd, imgs).permute(0, 2, 3, 1).reshape(-1, sd["codebook.embedding.weight"].shape[1])
        idx = farthest_point_sample(z, fs["K"])
        cb = z[idx].half()  # fp16-representable values: the fixture is exact in any float width
        d = torch.cdist(cb.double(), cb.double()) + torch.eye(cb.shape[0], dtype=torch.float64) * 1e9
        print(f"{family}: codebook {tuple(cb.shape)} min inter-code distance {d.min():.4f} |z| mean {z.norm(dim=1).mean():.3f}")
        np.save(os.path.join(GOLDEN_DIR, f"codebook_f{fs['down_ratio']}.npy"), cb.numpy())


def _np(t):
    return t.detach().cpu().numpy()


def make_goldens():
    assert ref_shims.reference_available(), "needs /root/reference"
    out = {}
    # ---- VQ-VAE round trips (config C1 and the f8 equivalent) ----
    for family in ("mnist", "caterv2"):
        params = syn.model_params(family)
        fs = params["first_stage_config"]["params"]
        sd = syn.make_vqvae_state_dict(fs)
        ref = ref_shims.build_reference_vqvae(fs, sd)
        C, res = fs["input_dim"], 16 * fs["down_ratio"]
        lo, hi = (-0.5, 0.5) if fs["down_ratio"] == 4 else (-1.0, 1.0)
        x = syn.structured_images(2, C, res, seed=4242, lo=lo, hi=hi)
        with torch.no_grad():
            idx = ref.encode(x.clone())
            rec = ref.decode(idx)
            z = ref.encoder(x.clone()).permute(0, 2, 3, 1).reshape(-1, sd["codebook.embedding.weight"].shape[1])
            dist = orc.vq_distances(z, sd["codebook.embedding.weight"])
            top2 = torch.topk(dist, 2, dim=1, largest=False)[0]
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"vqvae_f{fs['down_ratio']}.npz"),
                            family=family, image_seed=4242, idx=_np(idx).astype(np.int16),
                            rec=_np(rec), vq_gap=_np(top2[:, 1] - top2[:, 0]).astype(np.float32),
                            z_sample=_np(z[::37, ::13]))
        print(f"vqvae {family}: {idx.unique().numel()} codes used, min VQ gap {float((top2[:,1]-top2[:,0]).min()):.4g}")

    # ---- MAGE sampling ----
    for name, family, L, B, T, padded, with_speed, noise_seed in CASES:
        params = syn.model_params(family, frames_length=L, randomness=noise_seed is not None)
        sd = syn.make_mage_state_dict(params)
        batch = syn.make_batch(params, B, seed=1234, text_len=T, padded=padded, with_speed=with_speed)
        model = ref_shims.build_reference_mage(params, sd)
        captured = {}
        # capture the final prediction (for the top1-top2 gaps) without touching reference code
        h = model.generate_model.register_forward_hook(lambda m, i, o: captured.__setitem__("pred", o.detach()))
        t0 = time.time()
        with torch.no_grad():
            if noise_seed is not None:
                torch.manual_seed(noise_seed)  # reference draws torch.randn([B,64,H,W]) on the CPU generator (:661)
            video = model.autoregressive_generate({k: v.clone() for k, v in batch.items()})
        dt = time.time() - t0
        h.remove()
        pred = captured["pred"]
        tokens = torch.max(pred, -1)[1]
        top2 = torch.topk(pred, 2, dim=-1)[0]
        with torch.no_grad():
            tok0 = model.first_stage_encode(batch["images"][:, 0:1])[:, 0]
        rec = dict(family=family, frames_length=L, batch=B, text_len=T, padded=padded, with_speed=with_speed,
                   noise_seed=-1 if noise_seed is None else noise_seed,
                   tok0=_np(tok0).astype(np.int16), tokens=_np(tokens).astype(np.int16),
                   gap=_np(top2[..., 0] - top2[..., 1]).astype(np.float32),
                   logits_sample=_np(pred[:, :, ::5, ::5, ::16]).astype(np.float32))
        # pixels: full for the small cases, strided for the long one
        gen = video[:, 1:]
        rec["pixels"] = _np(gen if L <= 5 else gen[:, :, :, ::4, ::4]).astype(np.float32)
        rec["pixel_stride"] = 1 if L <= 5 else 4
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"mage_{name}.npz"), **rec)
        print(f"mage {name}: {dt:.1f}s, tokens {tuple(tokens.shape)}, {tokens.unique().numel()} codes, "
              f"min logit gap {float(rec['gap'].min()):.3g}, logit std {float(pred.std()):.3f}")


# MAGE+ branch (use_cids=False): (name, frames_length, batch, text_len, padded, noise_seed, the reference's line-93 edit)
PLUS_CASES = [
    ("plus_L4_b2", 4, 2, 12, False, 91, False),        # reference exactly as shipped (TransformerBlock line 92)
    ("plus_L5_b2_ln_pad", 5, 2, 14, True, 92, True),   # with the edit the reference documents for MAGE+ (line 93: ln_q / ln_kv)
    ("plus_L10_b1_ln", 10, 1, 20, False, 93, True),    # frames_length of the shipped mage+_cater*.yaml
]


def make_plus_goldens():
    """Reference outputs of the MAGE+ branch: the unmodified reference (and the reference with its own documented line 92->93
    edit, applied in memory) with use_cids=False and the stand-in first stage `mage_b200.synthetic.PatchLatentAE` (the shipped
    first stage, latent-diffusion's AutoencoderKL, is not vendored).  Stored: the latents the reference hands to
    first_stage_decode (captured at the module boundary by a forward hook on `generate_model`) and the pixels."""
    assert ref_shims.reference_available(), "needs /root/reference"
    for name, L, B, T, padded, noise_seed, edit in PLUS_CASES:
        params = syn.model_params("caterv2plus", frames_length=L)
        sd = syn.make_mage_state_dict(params)
        batch = syn.make_batch(params, B, seed=1234, text_len=T, padded=padded)
        model = ref_shims.build_reference_mage(params, sd, mage_plus_edit=edit)
        captured = {}
        h = model.generate_model.register_forward_hook(lambda m, i, o: captured.__setitem__("pred", o.detach()))
        t0 = time.time()
        with torch.no_grad():
            torch.manual_seed(noise_seed)
            video = model.autoregressive_generate({k: v.clone() for k, v in batch.items()})
            z0 = model.first_stage_encode(batch["images"][:, 0:1])[:, 0]
        h.remove()
        lat = captured["pred"].permute(0, 1, 4, 2, 3).contiguous()            # [B, L-1, c, h, w] (mage_model.py:689)
        rec = dict(frames_length=L, batch=B, text_len=T, padded=padded, noise_seed=noise_seed, ma_ln=edit,
                   z0=_np(z0).astype(np.float32), latents=_np(lat).astype(np.float32),
                   pixels=_np(video[:, 1:, :, ::4, ::4]).astype(np.float32), pixel_stride=4)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"mage_{name}.npz"), **rec)
        print(f"mage+ {name}: {time.time() - t0:.1f}s, latents {tuple(lat.shape)} std {float(lat.std()):.3f} max {float(lat.abs().max()):.3f}")


# stage-2 objective, forward half (MAGE.forward in eval mode): (name, family, frames_length, batch, text_len, padded, eps seed, test_flag)
FORWARD_CASES = [
    ("forward_L4_b2", "caterv2", 4, 2, 12, False, 61, False),
    ("forward_L8_b2_pad", "caterv2", 8, 2, 14, True, 62, False),
    ("forward_L16_b1", "caterv2", 16, 1, 20, False, 63, False),
    ("forward_L4_b2_testflag", "caterv2", 4, 2, 12, False, 64, True),
    ("forward_mnist_L4_b2", "mnist", 4, 2, 6, False, 65, False),
]


def forward_case_inputs(name):
    """(params, state_dict incl. the train-only posterior, batch with ALL frames, eps) of a FORWARD_CASES entry -- shared with the tests."""
    _, family, L, B, T, padded, eps_seed, test_flag = next(c for c in FORWARD_CASES if c[0] == name)
    params = syn.model_params(family, frames_length=L)
    sd = syn.make_mage_state_dict(params, posterior=True)
    batch = syn.make_batch(params, B, seed=4321, text_len=T, padded=padded, frames=L)
    eps = syn.make_noise(B, res=params["image_resolution"], seed=eps_seed) if params["randomness"] else None
    return params, sd, batch, eps, test_flag


def make_forward_goldens():
    """Loss values of the unmodified reference's MAGE.forward (mage_model.py:575-639) in eval mode (dropout off) on seeded
    synthetic checkpoints that include the 3-D conv posterior.  The one random draw of the pass, `torch.randn_like` in
    reparameterize (:571; with test_flag also :610), is replaced by a stored tensor `eps` for the duration of the call."""
    assert ref_shims.reference_available(), "needs /root/reference"
    for case in FORWARD_CASES:
        name, family, L, B, T, padded, eps_seed, test_flag = case
        params, sd, batch, eps, _ = forward_case_inputs(name)
        model = ref_shims.build_reference_mage(params, sd)
        assert not [k for k in model.state_dict() if k not in sd and not k.startswith("first_stage_model.")], "posterior keys missing"
        captured = {}
        hooks = []
        if params["randomness"]:
            hooks = [model.conv_mu2.register_forward_hook(lambda m, i, o: captured.__setitem__("mu", o.detach())),
                     model.conv_var2.register_forward_hook(lambda m, i, o: captured.__setitem__("logvar", o.detach()))]
        real = torch.randn_like
        torch.randn_like = lambda t, *a, **k: eps.clone().to(t.dtype) if tuple(t.shape) == tuple(eps.shape) else real(t, *a, **k)
        t0 = time.time()
        try:
            with torch.no_grad():
                loss, loss_dict = model({k: v.clone() for k, v in batch.items()}, test_flag=test_flag)
                tok = model.first_stage_encode(batch["images"])
        finally:
            torch.randn_like = real
            for h in hooks:
                h.remove()
        rec = dict(frames_length=L, batch=B, text_len=T, padded=padded, eps_seed=eps_seed, test_flag=test_flag,
                   tokens=_np(tok).astype(np.int16), final_loss=np.float64(float(loss)),
                   prediction=np.float64(loss_dict["val/prediction"]))
        if params["randomness"]:
            rec.update(kl_loss=np.float64(loss_dict["val/kl_loss"]), mu=_np(captured["mu"]).astype(np.float32),
                       logvar=_np(captured["logvar"]).astype(np.float32))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"mage_{name}.npz"), **rec)
        print(f"{name}: {time.time() - t0:.1f}s", {k: round(float(v), 6) for k, v in loss_dict.items()})


# MAGE+ (use_cids=False) forward: (name, frames_length, batch, text_len, padded, eps seed, test_flag, line-93 edit, auto_beta, v_kl)
FORWARD_PLUS_CASES = [
    ("forward_plus_L4_b2", 4, 2, 12, False, 81, False, False, False, 0.0),
    ("forward_plus_L10_b1_ln_pid", 10, 1, 20, False, 82, False, True, True, 100.0),   # the shipped mage+_cater*.yaml objective: PID beta
    ("forward_plus_L8_b2_pad_testflag", 8, 2, 14, True, 83, True, True, False, 0.0),
]


def forward_plus_case_inputs(name):
    _, L, B, T, padded, eps_seed, test_flag, edit, auto_beta, v_kl = next(c for c in FORWARD_PLUS_CASES if c[0] == name)
    params = syn.model_params("caterv2plus", frames_length=L)
    params = dict(params, auto_beta=auto_beta, v_kl=v_kl)
    sd = syn.make_mage_state_dict(params, posterior=True)
    batch = syn.make_batch(params, B, seed=4321, text_len=T, padded=padded, frames=L)
    eps = syn.make_noise(B, res=params["image_resolution"], seed=eps_seed)
    return params, sd, batch, eps, test_flag, edit


def make_forward_plus_goldens():
    """MAGE.forward of the unmodified reference (and of the reference with its documented line 92->93 edit) for use_cids=False in
    eval mode, stand-in first stage, the reparameterisation draw replaced by a stored tensor: MSE / KL / [PID beta] / final loss."""
    assert ref_shims.reference_available(), "needs /root/reference"
    for case in FORWARD_PLUS_CASES:
        name, L, B, T, padded, eps_seed, test_flag, edit, auto_beta, v_kl = case
        params, sd, batch, eps, _, _ = forward_plus_case_inputs(name)
        model = ref_shims.build_reference_mage(params, sd, mage_plus_edit=edit)
        captured = {}
        hooks = [model.conv_mu2.register_forward_hook(lambda m, i, o: captured.__setitem__("mu", o.detach())),
                 model.conv_var2.register_forward_hook(lambda m, i, o: captured.__setitem__("logvar", o.detach()))]
        real = torch.randn_like
        torch.randn_like = lambda t, *a, **k: eps.clone().to(t.dtype) if tuple(t.shape) == tuple(eps.shape) else real(t, *a, **k)
        t0 = time.time()
        try:
            with torch.no_grad():
                loss, loss_dict = model({k: v.clone() for k, v in batch.items()}, test_flag=test_flag)
        finally:
            torch.randn_like = real
            for h in hooks:
                h.remove()
        rec = dict(frames_length=L, batch=B, text_len=T, padded=padded, eps_seed=eps_seed, test_flag=test_flag, ma_ln=edit,
                   auto_beta=auto_beta, v_kl=v_kl, final_loss=np.float64(float(loss)), prediction=np.float64(loss_dict["val/prediction"]),
                   kl_loss=np.float64(loss_dict["val/kl_loss"]), beta=np.float64(loss_dict.get("val/beta", params["beta"])),
                   mu=_np(captured["mu"]).astype(np.float32), logvar=_np(captured["logvar"]).astype(np.float32))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"mage_{name}.npz"), **rec)
        print(f"{name}: {time.time() - t0:.1f}s", {k: round(float(v), 6) for k, v in loss_dict.items()})


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("codebooks", "all"):
        make_codebooks()
    if what in ("goldens", "all"):
        make_goldens()
    if what in ("plus", "all"):
        make_plus_goldens()
    if what in ("forward", "all"):
        make_forward_goldens()
    if what in ("forward_plus", "all"):
        make_forward_plus_goldens()
