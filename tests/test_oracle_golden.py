"""CPU: pin the oracle (oracle/mage_oracle.py) to the golden vectors produced by the
unmodified reference (oracle/make_golden.py), and -- where /root/reference exists -- to the
reference itself."""
import os

import numpy as np
import pytest
import torch

from mage_b200 import synthetic as syn
from oracle import mage_oracle as orc
from oracle import ref_shims
from tests.helpers import (FORWARD_CASES, FORWARD_PLUS_CASES, GOLDEN_DIR, MAGE_CASES, PLUS_CASES, free_running_token_report, load_case,
                           load_forward_case, load_forward_plus_case, load_plus_case)


@pytest.mark.parametrize("ratio", [4, 8])
def test_vqvae_round_trip_matches_reference_golden(ratio):
    g = np.load(os.path.join(GOLDEN_DIR, f"vqvae_f{ratio}.npz"))
    params = syn.model_params(str(g["family"]))
    fs = params["first_stage_config"]["params"]
    sd = syn.make_vqvae_state_dict(fs)
    lo, hi = (-0.5, 0.5) if ratio == 4 else (-1.0, 1.0)
    x = syn.structured_images(2, fs["input_dim"], 16 * ratio, seed=int(g["image_seed"]), lo=lo, hi=hi)
    with torch.no_grad():
        idx = orc.vqvae_encode(sd, x)
        rec = orc.vqvae_decode(sd, idx)
    gap = g["vq_gap"].reshape(idx.shape)
    neq = idx.numpy() != g["idx"]
    assert not (neq & (gap >= 1e-4)).any(), "VQ indices differ from the reference away from ties"
    if not neq.any():
        np.testing.assert_allclose(rec.numpy(), g["rec"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", MAGE_CASES)
def test_generate_matches_reference_golden(name):
    params, sd, batch, noise, g = load_case(name)
    tr = {}
    video = orc.generate(sd, batch, noise, tr)
    assert np.array_equal(tr["tok0"].numpy(), g["tok0"]), "first-frame VQ indices differ from the reference"
    rep = free_running_token_report(tr["tokens"].numpy(), g["tokens"], g["gap"], eps=1e-5)
    assert rep["excused"] == 0 and rep["compared"] == rep["positions"], "the oracle must reproduce the reference's tokens exactly"
    s = int(g["pixel_stride"])
    pix = video[:, 1:][..., ::s, ::s].numpy()
    np.testing.assert_allclose(pix, g["pixels"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(tr["gap"].numpy(), g["gap"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", ["cater_L4_b2_pad", "mnist_L5_b2", "cater_L10_b1"])
def test_incremental_order_equals_reference_order(name):
    params, sd, batch, noise, g = load_case(name)
    tr = {}
    video = orc.generate_incremental(sd, batch, noise, tr)
    rep = free_running_token_report(tr["tokens"].numpy(), g["tokens"], g["gap"], eps=1e-5)
    assert rep["excused"] == 0 and rep["compared"] == rep["positions"], "incremental and reference order must give the same tokens"
    s = int(g["pixel_stride"])
    np.testing.assert_allclose(video[:, 1:][..., ::s, ::s].numpy(), g["pixels"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", PLUS_CASES)
def test_mage_plus_branch_matches_reference_golden(name):
    """MAGE+ (use_cids=False): the oracle's continuous-latent restatement against the reference's own latents -- as shipped
    (TransformerBlock line 92) and with the reference's documented line-93 edit (ln_q / ln_kv)."""
    params, sd, batch, noise, g = load_plus_case(name)
    ae = syn.PatchLatentAE(**params["first_stage_config"]["params"])
    z0 = ae.encode(batch["images"][:, 0])
    np.testing.assert_allclose(z0.numpy(), g["z0"], rtol=0, atol=1e-6)
    lat = orc.generate_continuous(sd, z0, batch["text"], batch.get("speed"), noise, ma_ln=bool(g["ma_ln"]))
    np.testing.assert_allclose(lat.numpy(), g["latents"], rtol=0, atol=2e-5)
    B, Fr = lat.shape[:2]
    pix = ae.decode(lat.reshape(-1, *lat.shape[2:])).view(B, Fr, 3, 128, 128)[..., ::4, ::4]
    np.testing.assert_allclose(pix.numpy(), g["pixels"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", FORWARD_CASES)
def test_forward_loss_matches_reference_golden(name):
    """SURVEY.md §8 row N2, forward half: the oracle's restatement of MAGE.forward (teacher-forced full-sequence pass, 3-D conv
    video posterior, cross-entropy / KL / final loss; mage_model.py:575-639) against the unmodified reference's values in eval
    mode with the stored draw: VQ tokens of all frames identical, posterior mean / log-variance and the three losses to fp32
    rounding."""
    params, sd, batch, eps, test_flag, g = load_forward_case(name)
    tr = {}
    out = orc.forward_loss(sd, batch, eps, randomness=params["randomness"], beta=params.get("beta", 1.0), alpha=params.get("alpha", 0.0),
                           test_flag=test_flag, trace=tr)
    assert np.array_equal(tr["tokens"].numpy(), g["tokens"])
    assert abs(out["prediction"] - float(g["prediction"])) <= 2e-6 * abs(float(g["prediction"]))
    assert abs(out["final_loss"] - float(g["final_loss"])) <= 2e-6 * abs(float(g["final_loss"]))
    if params["randomness"]:
        assert abs(out["kl_loss"] - float(g["kl_loss"])) <= 2e-6 * abs(float(g["kl_loss"]))
        np.testing.assert_allclose(tr["mu"].numpy(), g["mu"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(tr["logvar"].numpy(), g["logvar"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", FORWARD_PLUS_CASES)
def test_forward_loss_mage_plus_matches_reference_golden(name):
    """The same for the MAGE+ branch (use_cids=False: continuous latents, MSE, `:621`): shipped line 92 and the documented line-93
    edit, fixed beta and the shipped objective's PID-controlled beta (auto_beta, v_kl = 100)."""
    params, sd, batch, eps, test_flag, g = load_forward_plus_case(name)
    ae = syn.PatchLatentAE(**params["first_stage_config"]["params"])
    ae.load_state_dict({k[len("first_stage_model."):]: v for k, v in sd.items() if k.startswith("first_stage_model.")})
    B, L = batch["images"].shape[:2]
    with torch.no_grad():
        z = ae.encode(batch["images"].reshape(B * L, *batch["images"].shape[2:]))
    z = z.view(B, L, *z.shape[1:])
    auto = bool(g["auto_beta"])
    tr = {}
    out = orc.forward_loss(sd, batch, eps, randomness=True, beta=float(g["beta"]), alpha=0.0 if auto else params["alpha"],
                           test_flag=test_flag, trace=tr, latents=z, ma_ln=bool(g["ma_ln"]))
    for key in ("prediction", "kl_loss", "final_loss"):
        assert abs(out[key] - float(g[key])) <= 3e-6 * abs(float(g[key])), (key, out[key], float(g[key]))
    np.testing.assert_allclose(tr["mu"].numpy(), g["mu"], rtol=0, atol=2e-5)
    if auto:   # the reference's first PID step from a fresh controller
        from mage_b200.model import PIDControl
        assert PIDControl().pid(float(g["v_kl"]), float(g["kl_loss"]))[0] == float(g["beta"])


def test_incremental_can_run_longer_than_checkpoint_positions_is_rejected():
    params, sd, batch, noise, _ = load_case("caterv1_L3_b1_norand")
    with pytest.raises(IndexError):
        orc.generate_incremental(sd, batch, noise, frames_length=5)


@pytest.mark.needs_reference
def test_oracle_bit_identical_to_live_reference():
    params = syn.model_params("caterv2", frames_length=3)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, 1, seed=77, text_len=9)
    model = ref_shims.build_reference_mage(params, sd)
    assert set(sd) <= set(model.state_dict()), "synthetic checkpoint has keys the reference does not"
    torch.manual_seed(5)
    with torch.no_grad():
        ref = model.autoregressive_generate({k: v.clone() for k, v in batch.items()})
    torch.manual_seed(5)
    noise = torch.randn(1, 64, 16, 16)
    out = orc.generate(sd, batch, noise)
    assert torch.equal(ref, out)


@pytest.mark.needs_reference
def test_synthetic_state_dict_loads_strict_into_reference_vqvae():
    for family in ("mnist", "caterv2"):
        fs = syn.model_params(family)["first_stage_config"]["params"]
        ref_shims.build_reference_vqvae(fs, syn.make_vqvae_state_dict(fs))
