#!/usr/bin/env python
"""CPU emulation of reduced-precision schedules for the f8 VQ-VAE DECODER (vqvae_model.py:203-214), to size the pixel-error
budget before building the kernels (north_star: decoded pixels within 1e-3 relative; the decoder feeds no token).

Modes per convolution (what the tensor-core kernel would issue per product):
  3  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo     (fp32-grade, the shipped scheme)
  2a a_hi*w_hi + a_lo*w_hi                 (activations fp32-grade, weights rounded to fp16)
  2w a_hi*w_hi + a_hi*w_lo                 (weights fp32-grade, activations rounded to fp16)
  1  a_hi*w_hi                             (both operands rounded to fp16, fp32 accumulation)
Usage: python tools/experiments/decoder_precision_emulation.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from mage_b200 import synthetic as syn

LAYERS = ["decoder.%d.%s" % (b, l) for b in (0, 2, 4, 6) for l in ("id_path", "block.1", "block.3", "block.5", "block.7")] + ["decoder.8"]


def h(x):
    return x.half().float()


def conv(sd, name, x, mode, padding=0):
    w, b = sd[name + ".weight"].double(), sd[name + ".bias"].double()
    x = x.double()
    if mode in ("1", "2w"):
        xa = h(x.float()).double()
    else:
        xa = x
    if mode in ("1", "2a"):
        wa = h(w.float()).double()
    else:
        wa = w
    return F.conv2d(xa, wa, b, padding=padding).float()


def dec_block(sd, name, x, modes):
    idp = conv(sd, name + ".id_path", x, modes.get(name + ".id_path", "3")) if (name + ".id_path.weight") in sd else x
    y = conv(sd, name + ".block.1", F.relu(x), modes.get(name + ".block.1", "3"))
    y = conv(sd, name + ".block.3", F.relu(y), modes.get(name + ".block.3", "3"), 1)
    y = conv(sd, name + ".block.5", F.relu(y), modes.get(name + ".block.5", "3"), 1)
    y = conv(sd, name + ".block.7", F.relu(y), modes.get(name + ".block.7", "3"), 1)
    return idp + y


def decode(sd, idx, modes):
    z = F.embedding(idx, sd["codebook.embedding.weight"]).permute(0, 3, 1, 2)
    x = dec_block(sd, "decoder.0", z, modes)
    for name in ("decoder.2", "decoder.4", "decoder.6"):
        x = dec_block(sd, name, F.interpolate(x, scale_factor=2, mode="nearest"), modes)
    return torch.tanh(conv(sd, "decoder.8", F.relu(x), modes.get("decoder.8", "3")))


def main():
    torch.set_num_threads(os.cpu_count())
    fs = syn.model_params("caterv2")["first_stage_config"]["params"]
    schedules = {
        "pixel-head conv (6.b7) 1-pass": {"decoder.6.block.7": "1"},
        "6.b7 2a": {"decoder.6.block.7": "2a"},
        "6.b7 2w": {"decoder.6.block.7": "2w"},
        "6.b5+6.b7 1-pass": {"decoder.6.block.5": "1", "decoder.6.block.7": "1"},
        "6.b3+6.b5+6.b7 1-pass": {"decoder.6.block.3": "1", "decoder.6.block.5": "1", "decoder.6.block.7": "1"},
        "block 6 all 1-pass": {"decoder.6." + l: "1" for l in ("block.1", "block.3", "block.5", "block.7")},
        "block 6 all 2a": {"decoder.6." + l: "2a" for l in ("block.1", "block.3", "block.5", "block.7")},
        "blocks 4+6 all 1-pass": {f"decoder.{b}." + l: "1" for b in (4, 6) for l in ("block.1", "block.3", "block.5", "block.7")},
        "blocks 4+6 3x3 convs 1-pass": {f"decoder.{b}." + l: "1" for b in (4, 6) for l in ("block.3", "block.5", "block.7")},
        "blocks 4+6 all 2a": {f"decoder.{b}." + l: "2a" for b in (4, 6) for l in ("block.1", "block.3", "block.5", "block.7")},
        "whole decoder 1-pass": {l: "1" for l in LAYERS},
        "whole decoder 2a": {l: "2a" for l in LAYERS},
        "whole decoder 2w": {l: "2w" for l in LAYERS},
    }
    print(f"{'schedule':38s} " + " ".join(f"{'seed %d rel-L2 | max-abs' % s:>26s}" for s in (7, 8, 9)))
    rows = {k: [] for k in schedules}
    for seed in (7, 8, 9):
        sd = {k: v for k, v in syn.make_vqvae_state_dict(fs, seed=seed).items()}
        g = torch.Generator().manual_seed(seed)
        idx = torch.randint(0, 512, (4, 16, 16), generator=g)
        with torch.no_grad():
            ref = decode(sd, idx, {})
            for name, modes in schedules.items():
                out = decode(sd, idx, modes)
                rows[name].append(((out - ref).norm() / ref.norm(), (out - ref).abs().max()))
    for name, vals in rows.items():
        print(f"{name:38s} " + " ".join(f"{float(r):14.2e} | {float(m):9.2e}" for r, m in vals))


if __name__ == "__main__":
    main()
