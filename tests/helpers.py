"""Shared helpers for the parity tests: rebuild a golden case's checkpoint / batch / noise."""
import os

import numpy as np
import torch

from mage_b200 import synthetic as syn

GOLDEN_DIR = syn.GOLDEN_DIR
MAGE_CASES = ["cater_L4_b2", "cater_L4_b2_pad", "caterv1_L3_b1_norand", "mnist_L5_b2", "cater_L10_b1"]


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


def tie_aware_token_check(tokens, ref_tokens, ref_gap, eps):
    """Greedy tokens must equal the reference's except where the reference's own top1-top2
    logit gap is below `eps`.  After such an excusable flip the sample's
    later frames are free to diverge (the cascade), so they are not compared.
    tokens/ref_tokens [B,F,h,w], ref_gap [B,F,h,w].  Returns (#compared, #excused flips)."""
    tokens = np.asarray(tokens)
    ref_tokens = np.asarray(ref_tokens)
    compared = excused = 0
    for b in range(tokens.shape[0]):
        for f in range(tokens.shape[1]):
            neq = tokens[b, f] != ref_tokens[b, f]
            compared += neq.size
            if neq.any():
                bad = neq & (ref_gap[b, f] >= eps)
                assert not bad.any(), (
                    f"sample {b} frame {f}: {int(bad.sum())} token mismatches with reference gap >= {eps} "
                    f"(min gap at mismatch {float(ref_gap[b, f][neq].min()):.3g})")
                excused += int(neq.sum())
                break  # later frames of this sample legitimately diverge
    return compared, excused
