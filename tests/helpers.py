"""Shared helpers for the parity tests: rebuild a golden case's checkpoint / batch / noise, and the
tie-aware parity report every whole-path test goes through."""
import os

import numpy as np
import torch

from mage_b200 import synthetic as syn

GOLDEN_DIR = syn.GOLDEN_DIR
MAGE_CASES = ["cater_L4_b2", "cater_L4_b2_pad", "caterv1_L3_b1_norand", "mnist_L5_b2", "cater_L10_b1"]

# A greedy-token mismatch is excusable only where the REFERENCE's own top1-top2 logit gap is below LOGIT_EPS.
# Measured logit error of the CUDA path against the reference (teacher-forced, printed by every test below): <= ~2e-5 at logit
# magnitudes ~2, so a flip needs a gap below ~4e-5; LOGIT_EPS is that with a small margin.
LOGIT_EPS = 5e-5
LOGIT_TOL = 5e-5      # teacher-forced logits vs the reference, relative to max(1, |logit|max)
PIX_REL = 1e-3        # BASELINE.json north_star: decoded pixels within 1e-3 relative (L2 over the compared frames)
PIX_ABS = 8e-3        # one 8-bit quantisation step of a [-1,1] pixel (own bound; north_star only states the relative one)
# the excuse window must stay a rare event: at most this fraction of the reference's positions may sit below LOGIT_EPS
NEAR_TIE_FRACTION_MAX = 1e-3


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"mage_{name}.npz")))
    family = str(g["family"])
    noise_seed = int(g["noise_seed"])
    params = syn.model_params(family, frames_length=int(g["frames_length"]), randomness=noise_seed >= 0)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, int(g["batch"]), seed=1234, text_len=int(g["text_len"]),
                           padded=bool(g["padded"]), with_speed=bool(g["with_speed"]))
    noise = None
    if noise_seed >= 0:
        torch.manual_seed(noise_seed)  # how the reference draws it (mage_model.py:661)
        noise = torch.randn(int(g["batch"]), 64, 16, 16)
    return params, sd, batch, noise, g


PLUS_CASES = ["plus_L4_b2", "plus_L5_b2_ln_pad", "plus_L10_b1_ln"]


def load_plus_case(name):
    """MAGE+ golden (use_cids=False, stand-in first stage): params (with the additive `ln_qkv` switch set like the golden's
    reference edit), checkpoint, batch, AdaIN noise as the reference drew it, golden arrays."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"mage_{name}.npz")))
    params = syn.model_params("caterv2plus", frames_length=int(g["frames_length"]))
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, int(g["batch"]), seed=1234, text_len=int(g["text_len"]), padded=bool(g["padded"]))
    torch.manual_seed(int(g["noise_seed"]))
    noise = torch.randn(int(g["batch"]), 64, 16, 16)
    return params, sd, batch, noise, g


FORWARD_CASES = ["forward_L4_b2", "forward_L8_b2_pad", "forward_L16_b1", "forward_L4_b2_testflag", "forward_mnist_L4_b2"]
FORWARD_FAMILY = {"forward_mnist_L4_b2": "mnist"}


def load_forward_case(name):
    """Golden of the reference's MAGE.forward in eval mode (stage-2 objective, forward half): params, checkpoint INCLUDING the
    train-only video posterior, a batch with all frames_length frames, the stored stand-in `eps` of the pass's one random draw
    (None without randomness), test_flag, golden arrays.  Mirrors oracle/make_golden.py::forward_case_inputs."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"mage_{name}.npz")))
    L, B = int(g["frames_length"]), int(g["batch"])
    params = syn.model_params(FORWARD_FAMILY.get(name, "caterv2"), frames_length=L)
    sd = syn.make_mage_state_dict(params, posterior=True)
    batch = syn.make_batch(params, B, seed=4321, text_len=int(g["text_len"]), padded=bool(g["padded"]), frames=L)
    eps = syn.make_noise(B, res=params["image_resolution"], seed=int(g["eps_seed"])) if params["randomness"] else None
    return params, sd, batch, eps, bool(g["test_flag"]), g


FORWARD_PLUS_CASES = ["forward_plus_L4_b2", "forward_plus_L10_b1_ln_pid", "forward_plus_L8_b2_pad_testflag"]


def load_forward_plus_case(name):
    """Golden of the reference's MAGE.forward for use_cids=False (MAGE+, stand-in first stage) in eval mode: params (with `ln_qkv`
    set like the golden's reference edit and the golden's auto_beta / v_kl), checkpoint incl. the posterior, batch with all
    frames, stored draw, test_flag, golden arrays.  Mirrors oracle/make_golden.py::forward_plus_case_inputs."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"mage_{name}.npz")))
    L, B = int(g["frames_length"]), int(g["batch"])
    params = syn.model_params("caterv2plus", frames_length=L)
    params = dict(params, auto_beta=bool(g["auto_beta"]), v_kl=float(g["v_kl"]))
    sd = syn.make_mage_state_dict(params, posterior=True)
    batch = syn.make_batch(params, B, seed=4321, text_len=int(g["text_len"]), padded=bool(g["padded"]), frames=L)
    eps = syn.make_noise(B, res=params["image_resolution"], seed=int(g["eps_seed"]))
    return params, sd, batch, eps, bool(g["test_flag"]), g


def pix_check(got, want, what="pixels"):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    rel = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)
    mx = np.abs(got - want).max() if got.size else 0.0
    assert rel <= PIX_REL, f"{what}: rel-L2 error {rel:.3e} > {PIX_REL}"
    assert mx <= PIX_ABS, f"{what}: max-abs error {mx:.3e} > {PIX_ABS}"
    return rel, mx


def free_running_token_report(tokens, ref_tokens, ref_gap, eps=LOGIT_EPS):
    """Greedy tokens of a FREE-RUNNING generate against the reference's.  A mismatch is tolerated only where the reference's
    own top1-top2 logit gap is below `eps`; after such a flip the sample's later frames legitimately diverge (the cascade), so
    they are not compared HERE -- the caller then re-checks every position teacher-forced (`parity_check`).
    tokens/ref_tokens [B,F,h,w], ref_gap [B,F,h,w].  Returns a dict:
      positions     B*F*h*w
      compared      positions actually compared (frames up to and including each sample's first diverging frame)
      excused       mismatches inside the excuse window
      skipped       list of (sample, first skipped frame) -- frames after a sample's first excused flip
      near_ties     reference positions with gap < eps (an upper bound for `excused` by construction)
      frames_equal  bool [B,F]: frames whose tokens are identical to the reference's (their pixels are comparable)"""
    tokens = np.asarray(tokens)
    ref_tokens = np.asarray(ref_tokens)
    ref_gap = np.asarray(ref_gap).reshape(tokens.shape)
    B, Fr = tokens.shape[:2]
    rep = dict(positions=int(tokens.size), compared=0, excused=0, skipped=[], near_ties=int((ref_gap < eps).sum()),
               frames_equal=np.zeros((B, Fr), dtype=bool))
    for b in range(B):
        for f in range(Fr):
            neq = tokens[b, f] != ref_tokens[b, f]
            rep["compared"] += int(neq.size)
            if not neq.any():
                rep["frames_equal"][b, f] = True
                continue
            bad = neq & (ref_gap[b, f] >= eps)
            assert not bad.any(), (
                f"sample {b} frame {f}: {int(bad.sum())} token mismatches where the reference's top1-top2 gap is >= {eps} "
                f"(gaps at the mismatches: {np.sort(ref_gap[b, f][neq])[:5]})")
            rep["excused"] += int(neq.sum())
            if f + 1 < Fr:
                rep["skipped"].append((b, f + 1))
            break
    return rep


def parity_check(model, batch, noise, video, ref_tokens, ref_gap, ref_frames, *, label, eps=LOGIT_EPS, pixel_stride=1,
                 ref_logits=None, always_teacher_forced=False, rows=None):
    """The whole-path parity gate (tokens bit-exact away from reference near-ties, pixels <= 1e-3), with its resolution made
    explicit: prints and bounds how much was compared and how much was excused.

    video        [B,L,C,H,W] of a free-running `model.autoregressive_generate(batch, noise)`; `model.last_tokens` are its tokens
    ref_tokens   [B,F,h,w] reference (golden or oracle) tokens, ref_gap their top1-top2 logit gaps
    ref_frames   [B,F,C,H/s,W/s] reference pixels of the generated frames (sampled with `pixel_stride` = s)
    ref_logits   optional [B,F,h*w,K] (or a golden's sample of it, see `logits_sample`) for the teacher-forced logit check
    rows         optional slice of the model's batch rows that the reference covers (first `rows` prompts)

    1. free-running tokens vs the reference (tie-aware); frames whose tokens agree have their pixels compared -- always;
    2. if a flip was excused (or `always_teacher_forced`): the model is re-run TEACHER-FORCED on the reference's tokens, so that
       every position of every frame is compared without the cascade (nothing is skipped), and the reference's tokens are pushed
       through the model's VQ-VAE decoder so every frame's pixels are compared as well."""
    ref_tokens_t = torch.as_tensor(np.asarray(ref_tokens)).to(torch.int64)
    n = ref_tokens_t.shape[0] if rows is None else rows
    tokens = model.last_tokens[:n].cpu().numpy()
    rep = free_running_token_report(tokens, ref_tokens, ref_gap, eps)
    s = pixel_stride
    gen = video[:n, 1:][..., ::s, ::s].cpu().numpy()
    ref_frames = np.asarray(ref_frames)
    eq = rep["frames_equal"]
    rel = mx = float("nan")
    if eq.any():
        rel, mx = pix_check(gen[eq], ref_frames[eq], f"{label}: pixels of the {int(eq.sum())} token-identical frames")
    tf = None
    if rep["excused"] > 0 or always_teacher_forced:
        sub = {k: v[:n] for k, v in batch.items()}
        tf_tokens, tf_logits = model.teacher_forced_tokens({k: v.to("cuda") for k, v in sub.items()}, ref_tokens_t,
                                                           noise=noise[:n] if noise is not None else None)
        neq = tf_tokens.cpu().numpy() != np.asarray(ref_tokens)
        gap = np.asarray(ref_gap).reshape(neq.shape)
        bad = neq & (gap >= eps)
        assert not bad.any(), f"{label}: {int(bad.sum())} teacher-forced mismatches where the reference gap is >= {eps}: {np.sort(gap[neq])[:5]}"
        tf = dict(compared=int(neq.size), excused=int(neq.sum()))
        assert tf["compared"] == rep["positions"]
        if ref_logits is not None:
            got = tf_logits.cpu().numpy()
            want = np.asarray(ref_logits)
            if want.shape != got.shape:   # golden sample: [:, :, ::5, ::5, ::16] of [B,F,16,16,K]
                B_, F_ = got.shape[:2]
                got = got.reshape(B_, F_, 16, 16, -1)[:, :, ::5, ::5, ::16]
            tf["logit_err"] = float(np.abs(got - want).max())
            tf["logit_max"] = float(np.abs(want).max())
            assert tf["logit_err"] <= LOGIT_TOL * max(1.0, tf["logit_max"]), \
                f"{label}: teacher-forced logits differ from the reference by {tf['logit_err']:.3e}"
        # decoder parity on EVERY frame: the reference's tokens through the model's own VQ-VAE decoder
        dec = model.first_stage_decode(ref_tokens_t.to("cuda"))[..., ::s, ::s].cpu().numpy()
        tf["pix_rel"], tf["pix_abs"] = pix_check(dec, ref_frames, f"{label}: decoder on the reference's tokens, all frames")
    print(f"[parity] {label}: positions {rep['positions']}, compared free-running {rep['compared']}, excused {rep['excused']}, "
          f"reference near-ties (gap < {eps:g}) {rep['near_ties']}, frames skipped after a flip {rep['skipped']}, "
          f"pixel rel/max-abs on token-identical frames {rel:.2e}/{mx:.2e}"
          + (f"; teacher-forced: compared {tf['compared']}, excused {tf['excused']}"
             + (f", logit err {tf['logit_err']:.2e} at |logit| <= {tf['logit_max']:.2f}" if "logit_err" in tf else "")
             + f", decoder-on-reference-tokens pixels {tf['pix_rel']:.2e}/{tf['pix_abs']:.2e}" if tf else ""))
    assert rep["excused"] <= rep["near_ties"]
    assert rep["near_ties"] <= max(3, NEAR_TIE_FRACTION_MAX * rep["positions"]), \
        f"{label}: the excuse window covers {rep['near_ties']} of {rep['positions']} reference positions -- too wide to mean anything"
    if not rep["skipped"]:
        assert rep["compared"] == rep["positions"]
    rep["teacher_forced"] = tf
    return rep
