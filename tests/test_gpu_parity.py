"""GPU parity proper: the CUDA sampling path (through the drop-in classes -> C ABI) against the
golden vectors of the unmodified reference and against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): greedy VQ-token indices identical to the reference -- checked
tie-aware, i.e. a mismatch is tolerated only where the reference's own top1-top2 logit gap is
below LOGIT_EPS (fp32 summation-order noise) -- and decoded pixels within 1e-3
relative (PIX_REL, on the L2 norm of the generated frames) plus a max-abs bound."""
import os

import numpy as np
import pytest
import torch

from mage_b200 import synthetic as syn
from tests.helpers import (FORWARD_CASES, FORWARD_PLUS_CASES, GOLDEN_DIR, LOGIT_EPS, MAGE_CASES, PLUS_CASES, load_case, load_forward_case,
                           load_forward_plus_case, load_plus_case, parity_check, pix_check)

pytestmark = pytest.mark.gpu

VQ_EPS = 1e-3      # VQ distances are O(10..100)


@pytest.fixture(params=["tc", "simt"], autouse=True)
def backend(request, monkeypatch):
    """Every parity case runs on both back ends: tcgen05 split-fp16 (default) and fp32 FFMA."""
    monkeypatch.setenv("MAGE_BACKEND", request.param)
    return request.param


def _build(params, sd):
    from mage_b200.config import instantiate_from_config
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": params})
    model.load_state_dict(sd)
    return model.to("cuda").eval()


def _pix_check(got, want):
    pix_check(got, want)


@pytest.mark.parametrize("ratio", [4, 8])
def test_vqvae_round_trip_vs_reference_golden(ratio):
    from modules.vqvae_model import VectorQuantizedVAE
    g = np.load(os.path.join(GOLDEN_DIR, f"vqvae_f{ratio}.npz"))
    fs = syn.model_params(str(g["family"]))["first_stage_config"]["params"]
    sd = syn.make_vqvae_state_dict(fs)
    m = VectorQuantizedVAE(**{k: v for k, v in fs.items()})
    m.load_state_dict(sd)
    m = m.to("cuda")
    lo, hi = (-0.5, 0.5) if ratio == 4 else (-1.0, 1.0)
    x = syn.structured_images(2, fs["input_dim"], 16 * ratio, seed=int(g["image_seed"]), lo=lo, hi=hi)
    idx = m.encode(x.to("cuda"))
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (2, 16, 16)
    neq = idx.cpu().numpy() != g["idx"]
    gap = g["vq_gap"].reshape(neq.shape)
    assert not (neq & (gap >= VQ_EPS)).any(), f"{int(neq.sum())} VQ index mismatches, min gap {gap[neq].min():.3g}"
    assert neq.sum() == 0, "expected bit-exact VQ indices on the conditioned codebook"
    rec = m.decode(torch.from_numpy(g["idx"].astype(np.int64)).to("cuda"))
    _pix_check(rec.cpu().numpy(), g["rec"])


def test_vqvae_encoder_features_vs_oracle():
    from mage_b200.engine import VQVAEEngine
    from oracle import mage_oracle as orc
    for family in ("mnist", "caterv2"):
        fs = syn.model_params(family)["first_stage_config"]["params"]
        sd = syn.make_vqvae_state_dict(fs)
        lo, hi = (-0.5, 0.5) if fs["down_ratio"] == 4 else (-1.0, 1.0)
        x = syn.structured_images(3, fs["input_dim"], 16 * fs["down_ratio"], seed=31, lo=lo, hi=hi)
        eng = VQVAEEngine({k: v.to("cuda") for k, v in sd.items()})
        z = eng.encode_features(x.to("cuda")).permute(0, 3, 1, 2).cpu()
        with torch.no_grad():
            want = orc.vqvae_encoder(sd, x)
        err = (z - want).abs().max().item()
        assert err <= 2e-4 * want.abs().max().item() + 1e-5, f"{family}: encoder feature max err {err:.3e}"


@pytest.mark.parametrize("name", MAGE_CASES)
def test_generate_vs_reference_golden(name):
    params, sd, batch, noise, g = load_case(name)
    model = _build(params, sd)
    video = model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=noise)
    B, L = int(g["batch"]), int(g["frames_length"])
    assert tuple(video.shape[:2]) == (B, L)
    assert torch.equal(video[:, 0].cpu(), batch["images"][:, 0]), "frame 0 must be the raw input frame (mage_model.py:691)"
    assert np.array_equal(model.last_tok0.cpu().numpy(), g["tok0"]), "first-frame VQ indices differ from the reference"
    parity_check(model, batch, noise, video, g["tokens"], g["gap"], g["pixels"], label=f"golden {name}",
                 pixel_stride=int(g["pixel_stride"]), ref_logits=g["logits_sample"])


def test_decoder_precision_budget(backend, monkeypatch):
    """DESIGN.md "decoder precision budget": the last decoder block's block.5 / block.7 (+ pixel head) convolutions -- 63 % of the
    decoder's FLOPs, feeding no token -- run single-pass fp16 by default.  Held here to the north_star bar (decoded pixels within
    1e-3 relative) against the fp32 CPU oracle on the tokens of all five reference goldens' family plus three seeds of random
    token maps and random-weight decoders, next to the all-fp32-grade mode (MAGE_DECODER_PRECISION=fp32: ~1e-6)."""
    if backend != "tc":
        pytest.skip("the FFMA back end has no reduced-precision mode")
    from mage_b200.engine import VQVAEEngine
    from oracle import mage_oracle as orc
    fs = syn.model_params("caterv2")["first_stage_config"]["params"]
    worst = 0.0
    for seed in (7, 8, 9):
        sd = syn.make_vqvae_state_dict(fs, seed=seed)
        g = torch.Generator().manual_seed(seed)
        idx = torch.randint(0, 512, (8, 16, 16), generator=g)
        if seed == 7:   # the goldens' checkpoint: decode the reference's own greedy tokens of the L=10 golden as well
            gold = np.load(os.path.join(GOLDEN_DIR, "mage_cater_L10_b1.npz"))
            idx = torch.cat([idx, torch.from_numpy(gold["tokens"].astype(np.int64)).view(-1, 16, 16)])
        with torch.no_grad():
            want = orc.vqvae_decode(sd, idx).numpy().astype(np.float64)
        errs = {}
        for mode in ("budget", "fp32"):
            monkeypatch.setenv("MAGE_DECODER_PRECISION", mode)
            eng = VQVAEEngine({k: v.to("cuda") for k, v in sd.items()})
            got = eng.decode(idx.to("cuda")).cpu().numpy().astype(np.float64)
            errs[mode] = (np.linalg.norm(got - want) / np.linalg.norm(want), np.abs(got - want).max())
        print(f"[parity] decoder precision, seed {seed}, {idx.shape[0]} frames: budget rel-L2 {errs['budget'][0]:.2e} max-abs "
              f"{errs['budget'][1]:.2e}; fp32-grade rel-L2 {errs['fp32'][0]:.2e} max-abs {errs['fp32'][1]:.2e}")
        assert errs["fp32"][0] <= 2e-5
        assert errs["budget"][0] <= 8e-4 and errs["budget"][1] <= 8e-3, errs   # bar 1e-3 with margin
        worst = max(worst, errs["budget"][0])
    assert worst > 1e-5, "the budget mode did not engage (results equal the fp32-grade path)"


def test_prelude_intermediates_vs_oracle():
    """Text encoder, motion anchor (cross-attention, AdaIN, speed) and per-step logits, teacher-forced by
    construction as long as the tokens agree."""
    from oracle import mage_oracle as orc
    params, sd, batch, noise, g = load_case("cater_L4_b2_pad")
    model = _build(params, sd)
    eng = model.engine()
    tr = {}
    eng.generate(batch["images"][:, 0].to("cuda"), batch["text"].to("cuda"), batch["speed"].to("cuda"), noise.to("cuda"), trace=tr)
    otr = {}
    orc.generate(sd, batch, noise, otr)
    B = batch["text"].shape[0]

    def close(a, b, what, tol=3e-5):
        err = (a.cpu() - b).abs().max().item()
        assert err <= tol * max(1.0, b.abs().max().item()), f"{what}: max err {err:.3e} (ref max {b.abs().max().item():.3e})"

    valid = (batch["text"] != 0)
    te = tr["text_emb"].cpu()            # [B,T,C]
    ote = otr["text_emb"].permute(1, 0, 2)  # oracle is [T,B,C]
    close(te[valid], ote[valid], "text encoder (valid tokens)")
    close(te, ote, "text encoder (all positions, padded keys feed the un-masked motion anchor)", 1e-4)
    close(tr["first_img"], otr["first_img"].permute(1, 0, 2), "first-frame token features")
    close(tr["anchor_ma"], otr["anchor_ma"], "motion anchor after MAEncoder", 1e-4)
    close(tr["anchor"], otr["anchor"], "motion anchor after AdaIN + speed", 2e-4)


def test_full_length_c5_shape_vs_incremental_oracle():
    """C5 geometry (CATER-v2, 128x128x32) at B=2: all 31 steps, K/V cache up to 32 positions."""
    from oracle import mage_oracle as orc
    params = syn.model_params("caterv2", frames_length=32)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, 2, seed=4321, text_len=20)
    noise = syn.make_noise(2, seed=5)
    model = _build(params, sd)
    video = model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=noise)
    otr = {"want_logits": True}
    want = orc.generate_incremental(sd, batch, noise, otr)
    assert np.array_equal(model.last_tok0.cpu().numpy(), otr["tok0"].numpy())
    # free-running AND teacher-forced (all 31 x 256 x B positions, logits to 5e-5, every frame's pixels): no cascade can hide anything
    parity_check(model, batch, noise, video, otr["tokens"].numpy(), otr["gap"].numpy(), want[:, 1:].numpy(),
                 label="C5 shape (CATER-v2 128x128x32) B=2", ref_logits=otr["logits"].numpy(), always_teacher_forced=True)


def test_batch_invariance_and_graph_replay_are_bit_exact():
    """Size-independent property used at full batch: a sample's result does not depend on which
    batch it is generated in (kernels are row-independent and deterministic), and CUDA-graph
    replays equal the eager launch sequence bit for bit."""
    params = syn.model_params("caterv2", frames_length=6)
    sd = syn.make_mage_state_dict(params)
    big = syn.make_batch(params, 16, seed=777, text_len=16)
    noise = syn.make_noise(16, seed=6)
    model = _build(params, sd)
    eng = model.engine()
    cu = lambda d: {k: v.to("cuda") for k, v in d.items()}
    v16 = model.autoregressive_generate(cu(big), noise=noise)
    t16 = model.last_tokens.clone()
    v16b = model.autoregressive_generate(cu(big), noise=noise)  # graph replay
    assert torch.equal(v16, v16b) and torch.equal(t16, model.last_tokens)
    small = {k: v[3:5] for k, v in big.items()}
    v2 = model.autoregressive_generate(cu(small), noise=noise[3:5])
    assert torch.equal(model.last_tokens, t16[3:5]), "tokens depend on the batch composition"
    assert torch.equal(v2, v16[3:5]), "pixels depend on the batch composition"
    eng.use_cuda_graph = False
    v2e = model.autoregressive_generate(cu(small), noise=noise[3:5])
    assert torch.equal(v2e, v2), "eager launches and graph replay disagree"
    eng.temporal_attn = "generic"
    v2g = model.autoregressive_generate(cu(small), noise=noise[3:5])
    assert (v2g - v2).abs().max().item() < 1e-4, "TMA-staged and generic temporal attention disagree"
    assert eng.kernels_per_generate and eng.kernels_per_generate > 100


def test_state_dict_round_trip_and_module_prefix():
    """main_mage.py:218-223 strips a DDP `module.` prefix; released checkpoints carry train-only tensors."""
    params = syn.model_params("mnist", frames_length=3)
    sd = syn.make_mage_state_dict(params)
    model = _build(params, sd)
    ck = {"module." + k: v for k, v in model.state_dict().items()}
    ck["module.conv3d.0.conv1.weight"] = torch.zeros(4)
    stripped = {k[7:]: v for k, v in ck.items()}
    model2 = _build(params, stripped)
    batch = syn.make_batch(params, 1, seed=9, text_len=9)
    cu = {k: v.to("cuda") for k, v in batch.items()}
    assert torch.equal(model.autoregressive_generate(cu), model2.autoregressive_generate(cu))


def test_streamed_host_output_equals_device_output():
    """autoregressive_generate(..., to_host=True) returns the clip in pinned host memory (frames copied out while later frames
    are generated): same values as the device tensor, frame 0 = the raw input frame (mage_model.py:691), on graph replay too."""
    params = syn.model_params("caterv2", frames_length=4)
    sd = syn.make_mage_state_dict(params)
    model = _build(params, sd)
    batch = syn.make_batch(params, 3, seed=77, text_len=10)
    noise = syn.make_noise(3, seed=5)
    cu = lambda d: {k: v.to("cuda") for k, v in d.items()}
    dev = model.autoregressive_generate(cu(batch), noise=noise)
    for _ in range(2):   # capture, then replay
        host = model.autoregressive_generate(cu(batch), noise=noise, to_host=True)
        assert not host.is_cuda and host.is_pinned() and tuple(host.shape) == tuple(dev.shape)
        assert torch.equal(host, dev.cpu())
    assert torch.equal(host[:, 0], batch["images"][:, 0])


def test_main_mage_split_test_entry(tmp_path):
    """`python main_mage.py --split test --test_model <dir>/model_best.pth` (main_mage.py:201-248): the yaml beside the checkpoint
    is loaded, the checkpoint's `module.`-prefixed state dict is accepted, every prompt is sampled and clamped."""
    import yaml

    import main_mage
    params = syn.model_params("caterv2", frames_length=3)
    sd = syn.make_mage_state_dict(params)
    (tmp_path / "config.yaml").write_text(yaml.safe_dump({"model": {"target": "modules.mage_model.MAGE", "params": params},
                                                          "data": {"target": "dataload.CATER", "params": {}}}))
    torch.save({"epoch": 1, "state_dict": {"module." + k: v for k, v in sd.items()}, "optimizer": {}}, tmp_path / "model_best.pth")
    out = tmp_path / "clips"
    opt = main_mage.parser.parse_args(["--split", "test", "--test_model", str(tmp_path / "model_best.pth"), "--synthetic", "3",
                                       "--batch-size", "2", "--out", str(out)])
    frames = main_mage.sampling(opt)
    assert frames == 3 * 2
    clips = sorted(out.glob("*.npy"))
    assert len(clips) == 3
    clip = np.load(clips[0])
    assert clip.shape == (3, 3, 128, 128) and np.abs(clip).max() <= 1.0
    # frame 0 is the (clamped) input frame of that prompt (mage_model.py:691, main_mage.py:242); the generated frames depend on
    # the AdaIN noise drawn from the global CPU generator inside the call, which the DataLoader also advances -- like the reference
    from dataload import SyntheticCaptionVideos
    ds = SyntheticCaptionVideos(params, 3, seed=1234)
    assert np.array_equal(clip[0], ds[0]["images"][0].clamp(-1, 1).numpy())
    assert np.isfinite(clip).all() and np.abs(clip[1:]).max() > 0.05


@pytest.mark.parametrize("family,L,B,B_check", [("mnist", 16, 1, 1),      # BASELINE configs[1]: single Moving MNIST 64x64x16, batch 1
                                                 ("mnist", 20, 32, 2),     # configs[2]: double Moving MNIST 64x64x20, batch 32
                                                 ("caterv1", 16, 16, 2)])  # configs[3]: CATER-GEN-v1 128x128x16, batch 16
def test_baseline_configs_full_length_vs_incremental_oracle(family, L, B, B_check):
    """The other BASELINE.json configurations at their full frame counts and batch sizes: the first B_check prompts are held to the
    CPU oracle (incremental order, proven equal to the reference order in tests/test_oracle_golden.py) -- tokens tie-aware, pixels
    to 1e-3 -- and, being batch-invariant, stand for the whole batch (test_batch_invariance_and_graph_replay_are_bit_exact)."""
    from oracle import mage_oracle as orc
    params = syn.model_params(family, frames_length=L)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, B, seed=99, text_len=14, padded=B > 1)
    noise = syn.make_noise(B, seed=8) if params["randomness"] else None
    model = _build(params, sd)
    video = model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=noise)
    assert tuple(video.shape[:2]) == (B, L)
    sub = {k: v[:B_check] for k, v in batch.items()}
    if B > 1:
        # the motion anchor attends padded caption positions (mage_model.py:92), so a prompt's result depends on the batch's
        # maximum caption length: keep the full batch's padded width for the oracle's sub-batch
        assert sub["text"].shape[1] == batch["text"].shape[1]
    otr = {"want_logits": True}
    want = orc.generate_incremental(sd, sub, noise[:B_check] if noise is not None else None, otr)
    assert np.array_equal(model.last_tok0[:B_check].cpu().numpy(), otr["tok0"].numpy())
    # teacher-forced as well (L = 16 / 20 at full length): every position and every frame is compared, no cascade
    parity_check(model, sub, noise, video, otr["tokens"].numpy(), otr["gap"].numpy(), want[:, 1:].numpy(),
                 label=f"{family} L={L} B={B} (first {B_check} prompts)", ref_logits=otr["logits"].numpy(),
                 always_teacher_forced=True, rows=B_check)


def test_optional_schedules_are_bit_exact():
    """Schedules: chunk streams (the batch cut into S chunks whose launch sequences run concurrently), decoder on a side stream,
    decode group size, programmatic dependent launch.  Rows are independent and every kernel is deterministic, so none of them
    may change a single bit -- of the tokens, of the device clip, or of the streamed host clip."""
    from mage_b200 import ops
    params = syn.model_params("caterv2", frames_length=6)
    sd = syn.make_mage_state_dict(params)
    batch = {k: v.to("cuda") for k, v in syn.make_batch(params, 5, seed=31, text_len=12).items()}
    noise = syn.make_noise(5, seed=2)
    model = _build(params, sd)
    eng = model.engine()
    eng.n_streams, eng.decode_group, eng.overlap_decode = 1, 4, False
    ref = model.autoregressive_generate(batch, noise=noise)
    ref_tok = model.last_tokens.clone()
    try:
        for streams, overlap, group, use_pdl in ((1, True, 4, False), (1, False, 1, False), (1, False, 4, True), (1, True, 2, True),
                                                 (2, False, 4, False), (3, False, 2, True), (5, False, 5, False), (0, False, 0, False)):
            eng.n_streams, eng.overlap_decode, eng.decode_group = streams, overlap, group
            eng._graphs.clear()
            ops.pdl(use_pdl)
            for it in range(2):   # capture + replay
                got = model.autoregressive_generate(batch, noise=noise, to_host=it == 1 and streams != 1)
                assert torch.equal(model.last_tokens, ref_tok), (streams, overlap, group, use_pdl)
                assert torch.equal(got.cuda(), ref), (streams, overlap, group, use_pdl)
        eng.use_cuda_graph = False   # the same multi-stream schedule launched eagerly
        eng.n_streams, eng.decode_group = 2, 2
        got = model.autoregressive_generate(batch, noise=noise)
        assert torch.equal(model.last_tokens, ref_tok) and torch.equal(got, ref)
        if getattr(eng, "fused_axial", False):
            # the two-kernel form of the H / W attention (QKV GEMM + axial kernel) differs from the fused kernel only in fp32
            # summation order inside the 16x16 attention: same tokens, pixels to fp32 noise
            eng.fused_axial = False
            got2 = model.autoregressive_generate(batch, noise=noise)
            assert torch.equal(model.last_tokens, ref_tok) and (got2 - ref).abs().max() < 2e-4
            eng.fused_axial = True
        # small-batch schedule: the decoder of the first frames BESIDE the steps on a share of the SMs (mage_sm_share), captured and eager
        keep_side = eng.side_sms, eng.side_frames, eng.side_group
        for graph in (True, False):
            eng.use_cuda_graph = graph
            for sms, frames, group in ((20, 3, 2), (64, 5, 1), (8, 2, 4)):
                eng.side_sms, eng.side_frames, eng.side_group = sms, frames, group
                eng.n_streams, eng.decode_group, eng.overlap_decode = 1, 4, False
                for it in range(2):
                    got4 = model.autoregressive_generate(batch, noise=noise, to_host=it == 1)
                    assert torch.equal(model.last_tokens, ref_tok) and torch.equal(got4.cuda(), ref), (graph, sms, frames, group, it)
        eng.side_sms, eng.side_frames, eng.side_group = keep_side
        eng.use_cuda_graph = False
        if hasattr(eng, "fused_ln"):
            # LayerNorm inside the producing kernel (token_taps only = the default; the residual-stream GEMMs too; not at all): the
            # same arithmetic row by row, so not a bit may change
            keep = eng.fused_ln_taps, eng.fused_ln
            for taps, gemms in ((True, True), (False, False), (False, True)):
                eng.fused_ln_taps, eng.fused_ln = taps, gemms
                got3 = model.autoregressive_generate(batch, noise=noise)
                assert torch.equal(model.last_tokens, ref_tok) and torch.equal(got3, ref), (taps, gemms)
            eng.fused_ln_taps, eng.fused_ln = keep
    finally:
        ops.pdl(True)   # the default


def test_activation_outside_the_tensor_core_operand_range_is_recomputed_in_fp32(monkeypatch, backend):
    """The tensor-core kernels carry fp32 values as fp16 hi/lo pairs, so an activation beyond +-65504 cannot be represented (the
    kernels flag it).  The shipped checkpoints stay far inside (every GEMM operand but the head's is a LayerNorm / attention /
    QuickGELU output), but the path must not depend on that: such a call is repeated on the fp32 SIMT kernels with a warning --
    same algorithm, the reference's behaviour (plain fp32), never a wrong clip and never an exception the reference would not
    raise.  Here the last block's c_proj bias pushes the residual stream that feeds the head to ~3e5."""
    import warnings

    from mage_b200 import _lib
    if backend != "tc":
        pytest.skip("the operand range is a property of the tensor-core back end")
    params = syn.model_params("caterv2", frames_length=4)
    sd = syn.make_mage_state_dict(params)
    big = {k: v.clone() for k, v in sd.items()}
    last = max(int(k.split(".")[2]) for k in big if k.startswith("generate_model.blocks."))
    big[f"generate_model.blocks.{last}.mlp.c_proj.bias"] += 3e5
    batch = {k: v.to("cuda") for k, v in syn.make_batch(params, 2, seed=3, text_len=9).items()}
    noise = syn.make_noise(2, seed=4)
    model = _build(params, big)
    with pytest.warns(RuntimeWarning, match="repeated on the fp32"):
        got = model.autoregressive_generate(batch, noise=noise)
    assert model.range_fallbacks == 1 and torch.isfinite(got).all()
    tokens = model.last_tokens.clone()
    # the same checkpoint on the fp32 kernels from the start: the identical clip
    monkeypatch.setenv("MAGE_BACKEND", "simt")
    want_model = _build(params, big)
    want = want_model.autoregressive_generate(batch, noise=noise)
    assert want_model.range_fallbacks == 0 and torch.equal(want_model.last_tokens, tokens) and torch.equal(want, got)
    monkeypatch.delenv("MAGE_BACKEND")
    # an in-range checkpoint never takes the detour ...
    ok = _build(params, sd)
    with warnings.catch_warnings(record=True) as seen:
        warnings.simplefilter("always")
        ok.autoregressive_generate(batch, noise=noise)
    assert ok.range_fallbacks == 0 and not [w for w in seen if "repeated on the fp32" in str(w.message)]
    # ... and the detour can be refused
    monkeypatch.setenv("MAGE_RANGE_FALLBACK", "0")
    with pytest.raises(_lib.MageSplitRangeError):
        model.autoregressive_generate(batch, noise=noise)


@pytest.mark.parametrize("name", FORWARD_CASES)
def test_forward_loss_vs_reference_golden(name, backend):
    """SURVEY.md §8 row N2, forward half: MAGE.forward in eval mode (what the reference's periodic validation computes,
    main_mage.py:163-176) through the CUDA kernels against the unmodified reference's golden values: VQ tokens of all frames
    identical, cross-entropy to 2e-5 relative, KL and final loss to 1e-4 (fp32 summation order; the KL term sums exp(logvar), which
    multiplies the rounding of logvar -- the output of twelve 3-D convolutions and GroupNorms -- by exp(logvar), and the final
    loss of these cases is mostly beta * KL).  The 3-D conv posterior (84.9 M parameters) runs as tensor-core implicit GEMMs with the temporal taps folded into the
    channel axis; the teacher-forced pass is the sampling path's incremental pass with the given tokens fed back."""
    if backend != "tc":
        pytest.skip("the objective's forward pass is built on the tensor-core back end")
    params, sd, batch, eps, test_flag, g = load_forward_case(name)
    from mage_b200.config import instantiate_from_config
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": dict(params, with_posterior=True)})
    model.load_state_dict(sd)
    model = model.to("cuda").eval()
    final, loss_dict = model({k: v.to("cuda") for k, v in batch.items()}, test_flag=test_flag, eps=eps)
    tok = model.last_tokens_all.cpu().numpy()
    print(f"[parity] forward {name}: token mismatches {int((tok != g['tokens']).sum())} of {tok.size}; "
          + ", ".join(f"{k} {v:.7g}" for k, v in loss_dict.items()) + f"; reference prediction {float(g['prediction']):.7g} "
          f"final {float(g['final_loss']):.7g}")
    assert np.array_equal(tok, g["tokens"])
    assert final.dim() == 0 and abs(final.item() - loss_dict["val/final_loss"]) == 0
    for key in ("prediction", "kl_loss", "final_loss"):
        if key in g:
            got, want = loss_dict["val/" + key], float(g[key])
            assert abs(got - want) <= (2e-5 if key == "prediction" else 1e-4) * abs(want), (key, got, want)
    if "mu" in g:   # the posterior's two heads themselves, before the exponential
        ml = model.engine().last_mu_logvar.view(-1, 16, 16, 128).permute(0, 3, 1, 2).cpu().numpy()
        for k, arr in (("mu", ml[:, :64]), ("logvar", ml[:, 64:])):
            err = np.abs(arr - g[k]).max()
            print(f"[parity] forward {name}: max |{k} - reference| = {err:.2e} (max |{k}| {np.abs(g[k]).max():.2f})")
            assert err <= 4e-5 * np.abs(g[k]).max()   # fp32 summation order through 12 convolutions of K = 13824 + GroupNorms
    assert ("val/kl_loss" in loss_dict) == bool(params["randomness"])
    # the teacher-forced pass in full-sequence form (all L positions per GEMM: the default) against the sampling path's
    # one-position-per-step pass fed the given tokens: the same arithmetic per row
    eng = model.engine()
    ce_full = eng.last_ce_rows.clone()
    dev_batch = {k: v.to("cuda") for k, v in batch.items()}
    inc = eng.forward_loss(dev_batch["images"], dev_batch["text"], dev_batch["speed"].float(), eps.to("cuda") if eps is not None else None,
                           test_flag, incremental=True)
    diff = (eng.last_ce_rows - ce_full).abs().max().item()
    print(f"[parity] forward {name}: per-row cross-entropy, full-sequence vs incremental pass: max |diff| {diff:.2e}")
    assert diff <= 2e-5 and abs(inc["prediction"].item() - loss_dict["val/prediction"]) <= 1e-6 * loss_dict["val/prediction"]
    # training mode is refused loudly (dropout / gradients are not built), and so is a model without the posterior tensors
    with pytest.raises(NotImplementedError):
        model.train()({k: v.to("cuda") for k, v in batch.items()})
    if params["randomness"]:
        plain = _build(params, sd)
        with pytest.raises(RuntimeError, match="with_posterior"):
            plain({k: v.to("cuda") for k, v in batch.items()}, eps=eps)


def test_graph_cache_is_bounded_and_shares_one_pool():
    """Captions cannot be padded to a common length (the motion anchor attends padded positions, mage_model.py:92), so real data
    yields one CUDA graph per caption length: the cache keeps at most `max_graphs` signatures (LRU) and all graphs live in ONE
    memory pool -- device memory must not grow with the number of signatures seen."""
    params = syn.model_params("caterv2", frames_length=4)
    sd = syn.make_mage_state_dict(params)
    model = _build(params, sd)
    eng = model.engine()
    eng.max_graphs = 3
    noise = syn.make_noise(2, seed=2)
    outs, mem = {}, []
    for T in (8, 9, 10, 11, 12, 8, 12):
        batch = {k: v.to("cuda") for k, v in syn.make_batch(params, 2, seed=31, text_len=T).items()}
        v = model.autoregressive_generate(batch, noise=noise)
        if T in outs:
            assert torch.equal(outs[T], v), "a re-captured signature must reproduce the earlier result"
        outs[T] = v
        assert len(eng._graphs) <= 3
        torch.cuda.synchronize()
        mem.append(torch.cuda.memory_reserved())
    assert mem[-1] <= mem[1] * 1.25 + (64 << 20), f"reserved memory grew with the number of caption lengths: {mem}"


def test_out_of_vocabulary_caption_on_the_device_is_refused():
    """main_mage.py moves the batch to the device before the call, so the host-side range check never ran there and
    text_embed_kernel would have read past the vocabulary table.  The kernel now flags the id (never dereferences it) and the
    call raises IndexError like nn.Embedding (mage_model.py:228)."""
    params = syn.model_params("caterv1", frames_length=3)   # vocabulary of 30 ids
    sd = syn.make_mage_state_dict(params)
    model = _build(params, sd)
    batch = syn.make_batch(params, 2, seed=3, text_len=10)
    batch["text"][1, 4] = 45                                  # a CATER-v2 id (vocab 50) fed to a CATER-v1 checkpoint
    with pytest.raises(IndexError):
        model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=syn.make_noise(2))
    batch["text"][1, 4] = 5
    out = model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=syn.make_noise(2))
    assert torch.isfinite(out).all() and out.is_contiguous()


def test_engine_follows_child_and_in_place_weight_updates(tmp_path):
    """loading into a submodule (`first_stage_model.init_from_ckpt`, the way every shipped yaml loads the VQ-VAE,
    vqvae_model.py:222-231) or editing a parameter in place must not leave the sampler on stale packed weights."""
    from modules.vqvae_model import VectorQuantizedVAE
    params = syn.model_params("caterv2", frames_length=3)
    sd = syn.make_mage_state_dict(params)
    fs = params["first_stage_config"]["params"]
    fs_sd = {k[len("first_stage_model."):]: v for k, v in sd.items() if k.startswith("first_stage_model.")}
    torch.save(fs_sd, tmp_path / "vqvae.pt")
    # (a) the first stage restores itself from ckpt_path inside the MAGE constructor
    p2 = {**params, "first_stage_config": {"target": params["first_stage_config"]["target"],
                                            "params": {**fs, "ckpt_path": str(tmp_path / "vqvae.pt")}}}
    from mage_b200.config import instantiate_from_config
    m2 = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": p2})
    m2.load_state_dict({k: v for k, v in sd.items() if not k.startswith("first_stage_model.")}, strict=False)
    m2 = m2.to("cuda").eval()
    m1 = _build(params, sd)
    batch = {k: v.to("cuda") for k, v in syn.make_batch(params, 2, seed=5, text_len=9).items()}
    noise = syn.make_noise(2, seed=1)
    want = m1.autoregressive_generate(batch, noise=noise)
    assert torch.equal(m2.autoregressive_generate(batch, noise=noise), want), "first stage loaded through ckpt_path differs"
    # (b) a child reload after the first generate: start from a different VQ-VAE, then restore the right one through the child
    other = syn.make_vqvae_state_dict(fs, seed=123)
    m3 = _build(params, {**sd, **{"first_stage_model." + k: v for k, v in other.items()}})
    assert not torch.equal(m3.autoregressive_generate(batch, noise=noise), want)
    m3.first_stage_model.init_from_ckpt(str(tmp_path / "vqvae.pt"))
    assert torch.equal(m3.autoregressive_generate(batch, noise=noise), want), "engine kept the stale first-stage weights"
    # (c) an in-place edit of a parameter is picked up as well
    with torch.no_grad():
        m3.generate_model.out.bias.add_(1.0)    # uniform shift of every logit: same argmax, new version counter
    assert not m3._engine_valid()
    assert torch.equal(m3.autoregressive_generate(batch, noise=noise), want)
    # standalone first stage through its own ckpt_path
    vq = VectorQuantizedVAE(**{**fs, "ckpt_path": str(tmp_path / "vqvae.pt")}).to("cuda")
    x = syn.structured_images(2, fs["input_dim"], 128, seed=4)
    assert torch.equal(vq.encode(x.to("cuda")), m1.first_stage_model.encode(x.to("cuda")))


def test_c5_full_batch_rows_equal_small_batch_rows():
    """BASELINE configs[4] at its full size (CATER-v2 128x128x32, 64 prompts): the first two prompts of the 64-prompt call are
    bit-identical to the same prompts generated alone (which test_full_length_c5_shape_vs_incremental_oracle holds to the oracle),
    every frame is finite and in tanh range, and the greedy tokens are valid code indices."""
    params = syn.model_params("caterv2", frames_length=32)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, 64, seed=4321, text_len=20)
    noise = syn.make_noise(64, seed=5)
    model = _build(params, sd)
    cu = lambda d: {k: v.to("cuda") for k, v in d.items()}
    v64 = model.autoregressive_generate(cu(batch), noise=noise)
    t64 = model.last_tokens.clone()
    assert tuple(v64.shape) == (64, 32, 3, 128, 128) and torch.isfinite(v64).all() and v64[:, 1:].abs().max() <= 1.0
    assert t64.min() >= 0 and t64.max() < params["codebook_size"]
    v2 = model.autoregressive_generate(cu({k: v[:2] for k, v in batch.items()}), noise=noise[:2])
    assert torch.equal(model.last_tokens, t64[:2]) and torch.equal(v2, v64[:2])


@pytest.mark.parametrize("name", MAGE_CASES)
def test_teacher_forced_tokens_vs_reference_golden(name):
    """with every step fed the REFERENCE's previous tokens, each position's greedy choice must equal the
    reference's unless the reference's own top1-top2 logit gap is below LOGIT_EPS -- checked at every position of every frame
    (no cascade, so nothing is skipped after a flip)."""
    params, sd, batch, noise, g = load_case(name)
    model = _build(params, sd)
    ref = torch.from_numpy(g["tokens"].astype(np.int64))
    tokens, logits = model.teacher_forced_tokens({k: v.to("cuda") for k, v in batch.items()}, ref, noise=noise)
    neq = tokens.cpu().numpy() != g["tokens"]
    gap = g["gap"].reshape(neq.shape)
    assert not (neq & (gap >= LOGIT_EPS)).any(), f"{int((neq & (gap >= LOGIT_EPS)).sum())} teacher-forced mismatches away from ties"
    # the recorded logits reproduce the recorded choice and match the reference's last-iteration logits (sampled in the golden)
    assert torch.equal(logits.argmax(-1).view_as(tokens), tokens)
    B, F = tokens.shape[:2]
    got = logits.view(B, F, 16, 16, -1)[:, :, ::5, ::5, ::16].cpu().numpy()
    err = np.abs(got - g["logits_sample"]).max()
    print(f"[parity] teacher-forced golden {name}: {neq.size} positions, {int(neq.sum())} excused, {int((gap < LOGIT_EPS).sum())} reference "
          f"near-ties, logit err {err:.2e} at |logit| <= {np.abs(g['logits_sample']).max():.2f}")
    assert err <= 5e-5 * max(1.0, np.abs(g["logits_sample"]).max()), f"teacher-forced logits differ from the reference by {err:.3e}"


@pytest.mark.parametrize("text_len,with_speed", [(3, True), (38, True), (20, False)])
def test_caption_length_extremes_and_missing_speed(text_len, with_speed):
    """Edge inputs of the batch-dict contract (dataload.py:260,370): the shortest caption ([CLS] word [SEP]), the longest the
    text encoder accepts (context_length 38 for CATER-v2), and a batch without 'speed' (mage_model.py:666 skips the embedding)."""
    from oracle import mage_oracle as orc
    params = syn.model_params("caterv2", frames_length=3)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, 2, seed=11, text_len=text_len, padded=text_len > 8, with_speed=with_speed)
    noise = syn.make_noise(2, seed=9)
    model = _build(params, sd)
    video = model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=noise)
    otr = {}
    want = orc.generate(sd, batch, noise, otr)
    assert np.array_equal(model.last_tok0.cpu().numpy(), otr["tok0"].numpy())
    parity_check(model, batch, noise, video, otr["tokens"].numpy(), otr["gap"].numpy(), want[:, 1:].numpy(),
                 label=f"caption T={text_len} speed={with_speed}")
    assert torch.equal(video[:, 0].cpu(), batch["images"][:, 0])


def test_too_long_caption_is_refused_like_the_reference():
    """A caption longer than `context_length` indexes past the learned position table: the reference raises (IndexError from
    nn.Embedding, mage_model.py:228-231); here the call must fail loudly as well instead of reading out of bounds."""
    params = syn.model_params("caterv2", frames_length=3)
    sd = syn.make_mage_state_dict(params)
    model = _build(params, sd)
    batch = syn.make_batch(params, 1, seed=3, text_len=20)
    batch["text"] = torch.cat([batch["text"], torch.full((1, 30), 5, dtype=torch.long)], 1)   # 50 > 38 positions
    with pytest.raises(Exception):
        model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=syn.make_noise(1))
        torch.cuda.synchronize()


@pytest.mark.parametrize("family,L,B", [("caterv2", 2, 1), ("mnist", 2, 3), ("caterv1", 5, 5)])
def test_small_and_odd_shapes_vs_oracle(family, L, B):
    """Shortest clip (one generated frame), odd batch sizes (odd row-tile counts: no CTA pairing on some layers)."""
    from oracle import mage_oracle as orc
    params = syn.model_params(family, frames_length=L)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, B, seed=17, text_len=9)
    noise = syn.make_noise(B, seed=4) if params["randomness"] else None
    model = _build(params, sd)
    video = model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=noise)
    otr = {}
    want = orc.generate(sd, batch, noise, otr)
    assert tuple(video.shape) == tuple(want.shape)
    assert np.array_equal(model.last_tok0.cpu().numpy(), otr["tok0"].numpy())
    parity_check(model, batch, noise, video, otr["tokens"].numpy(), otr["gap"].numpy(), want[:, 1:].numpy(),
                 label=f"{family} L={L} B={B}")


def test_main_mage_seeded_clips_do_not_depend_on_batch_size(tmp_path):
    """`--seed`: a prompt's AdaIN noise is a function of (seed, global prompt index), so the clips written by the entry are the same
    for --batch-size 1 and 3 (and therefore for any prompt-shard over GPUs)."""
    import yaml

    import main_mage
    params = syn.model_params("caterv2", frames_length=3)
    (tmp_path / "config.yaml").write_text(yaml.safe_dump({"model": {"target": "modules.mage_model.MAGE", "params": params},
                                                          "data": {"target": "dataload.CATER", "params": {}}}))
    torch.save({"state_dict": syn.make_mage_state_dict(params)}, tmp_path / "model_best.pth")
    clips = {}
    for bs in (1, 3):
        out = tmp_path / f"clips{bs}"
        opt = main_mage.parser.parse_args(["--split", "test", "--test_model", str(tmp_path / "model_best.pth"), "--synthetic", "4",
                                           "--batch-size", str(bs), "--seed", "5", "--out", str(out)])
        assert main_mage.sampling(opt) == 4 * 2
        clips[bs] = {p.name: np.load(p) for p in sorted(out.glob("*.npy"))}
    assert len(clips[1]) == 4 and clips[1].keys() == clips[3].keys()
    for k in clips[1]:
        assert np.array_equal(clips[1][k], clips[3][k]), k


@pytest.mark.parametrize("name", FORWARD_PLUS_CASES)
def test_forward_loss_mage_plus_vs_reference_golden(name, backend):
    """MAGE.forward for use_cids=False (MAGE+: Linear latent embed, full-sequence decoder, GroupNorm(32) -> SiLU -> 1x1x1 conv head,
    MSE; mage_model.py:583,621) in eval mode against the unmodified reference's goldens -- shipped line 92 and the documented
    line-93 edit, fixed beta and the shipped objective's PID-controlled beta (auto_beta, v_kl = 100): MSE to 2e-5 relative, KL and
    final loss to 1e-4, beta equal."""
    if backend != "tc":
        pytest.skip("the objective's forward pass is built on the tensor-core back end")
    params, sd, batch, eps, test_flag, g = load_forward_plus_case(name)
    params = dict(params, with_posterior=True)
    params["ma_config"] = {"target": params["ma_config"]["target"], "params": dict(params["ma_config"]["params"], ln_qkv=bool(g["ma_ln"]))}
    from mage_b200.config import instantiate_from_config
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": params})
    model.load_state_dict(sd)
    model = model.to("cuda").eval()
    final, loss_dict = model({k: v.to("cuda") for k, v in batch.items()}, test_flag=test_flag, eps=eps)
    print(f"[parity] forward {name}: " + ", ".join(f"{k} {v:.7g}" for k, v in loss_dict.items())
          + f"; reference prediction {float(g['prediction']):.7g} kl {float(g['kl_loss']):.7g} final {float(g['final_loss']):.7g}")
    for key in ("prediction", "kl_loss", "final_loss"):
        got, want = loss_dict["val/" + key], float(g[key])
        assert abs(got - want) <= (2e-5 if key == "prediction" else 1e-4) * abs(want), (key, got, want)
    if bool(g["auto_beta"]):
        assert loss_dict["val/beta"] == float(g["beta"])
    else:
        assert "val/beta" not in loss_dict


def test_main_mage_split_val_is_the_reference_validation_loss(tmp_path, backend):
    """`--split val` (additive): the reference's periodic validation (main_mage.py:163-182) for one checkpoint -- MAGE.forward in
    eval mode per batch, mean over the batches -- against the oracle's forward_loss on the same clips and draws."""
    import yaml

    import main_mage
    from mage_b200 import shard
    from oracle import mage_oracle as orc
    if backend != "tc":
        pytest.skip("the objective's forward pass is built on the tensor-core back end")
    params = syn.model_params("caterv2", frames_length=4)
    (tmp_path / "config.yaml").write_text(yaml.safe_dump({"model": {"target": "modules.mage_model.MAGE", "params": params},
                                                          "data": {"target": "dataload.CATER", "params": {}}}))
    sd = syn.make_mage_state_dict(params, posterior=True)
    torch.save({"state_dict": sd}, tmp_path / "model_best.pth")
    opt = main_mage.parser.parse_args(["--split", "val", "--test_model", str(tmp_path / "model_best.pth"), "--synthetic", "3",
                                       "--batch-size", "2", "--seed", "9"])
    got = main_mage.validation(opt)
    clips = syn.make_batch(params, 3, seed=9, text_len=20, frames=4)
    want = []
    for lo, hi in ((0, 2), (2, 3)):
        b = {k: v[lo:hi] for k, v in clips.items()}
        eps = shard.noise_for_prompts(9, range(lo, hi), params["image_resolution"])
        want.append(orc.forward_loss(sd, b, eps, randomness=True, beta=params["beta"], alpha=params["alpha"])["final_loss"])
    want = sum(want) / len(want)
    print(f"[parity] --split val: test_loss {got:.6f}, oracle {want:.6f}")
    assert abs(got - want) <= 1e-4 * abs(want)


def test_main_mage_caption_and_image_prompt(tmp_path):
    """Additive entry: one clip from a caption (the dataset's word-level vocabulary) and a first-frame image file."""
    import yaml
    from PIL import Image

    import main_mage
    params = syn.model_params("caterv2", frames_length=3)
    (tmp_path / "config.yaml").write_text(yaml.safe_dump({"model": {"target": "modules.mage_model.MAGE", "params": params},
                                                          "data": {"target": "dataload.CATER", "params": {"dataset": "caterv2"}}}))
    torch.save({"state_dict": syn.make_mage_state_dict(params)}, tmp_path / "model_best.pth")
    img = (syn.structured_images(1, 3, 128, seed=3)[0].permute(1, 2, 0).numpy() * 0.5 + 0.5) * 255
    Image.fromarray(img.astype(np.uint8)).save(tmp_path / "first.png")
    opt = main_mage.parser.parse_args(["--split", "test", "--test_model", str(tmp_path / "model_best.pth"), "--caption",
                                       "the small blue rubber sphere is picked up and placed to (2, -3).", "--image",
                                       str(tmp_path / "first.png"), "--out", str(tmp_path / "o"), "--gifs"])
    assert main_mage.sampling(opt) == 2
    clip = np.load(tmp_path / "o" / "caption_0.npy")
    assert clip.shape == (3, 3, 128, 128) and np.isfinite(clip).all()
    assert (tmp_path / "videos" / "caption_0.gif").exists()


@pytest.mark.parametrize("name", PLUS_CASES)
def test_mage_plus_branch_vs_reference_golden(name, backend):
    """SURVEY.md §8f N1 -- the MAGE+ transformer branch (use_cids=False: Linear(4->512) embed, GroupNorm -> SiLU -> 1x1x1 Conv3d
    head over all temporal slots, continuous autoregression; `ln_qkv` = the reference's documented line-93 edit as a config switch)
    against the reference's own latents and pixels.  The first stage is the stand-in PatchLatentAE run as a plain torch module
    (the shipped AutoencoderKL is not vendored: parity unpinned for it), everything between its two calls runs on the CUDA path."""
    if backend != "tc":
        pytest.skip("the MAGE+ branch runs on the tensor-core back end")
    params, sd, batch, noise, g = load_plus_case(name)
    params["ma_config"]["params"]["ln_qkv"] = bool(g["ma_ln"])
    model = _build(params, sd)
    video = model.autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=noise)
    B, L = int(g["batch"]), int(g["frames_length"])
    assert tuple(video.shape) == (B, L, 3, 128, 128) and video.is_contiguous()
    assert torch.equal(video[:, 0].cpu(), batch["images"][:, 0]), "frame 0 must be the raw input frame (mage_model.py:691)"
    lat = model.last_latents.cpu().numpy()
    err = np.abs(lat - g["latents"]).max()
    rel, mx = pix_check(video[:, 1:][..., ::4, ::4].cpu().numpy(), g["pixels"], f"MAGE+ {name}")
    print(f"[parity] MAGE+ {name}: latents max-abs err {err:.2e} at |latent| <= {np.abs(g['latents']).max():.2f} over {L - 1} autoregressive "
          f"slots, pixels rel-L2 {rel:.2e} max-abs {mx:.2e}")
    assert err <= 2e-4 * max(1.0, np.abs(g["latents"]).max())
    # the switch matters: the other TransformerBlock line gives a different clip
    params2 = dict(params, ma_config={"target": params["ma_config"]["target"], "params": dict(params["ma_config"]["params"], ln_qkv=not bool(g["ma_ln"]))})
    other = _build(params2, sd).autoregressive_generate({k: v.to("cuda") for k, v in batch.items()}, noise=noise)
    assert (other - video).abs().max() > 1e-3


def test_mage_plus_full_length_teacher_forced_and_free_running(backend):
    """MAGE+ at the frames_length BASELINE configs[4] names (L = 32).  Continuous autoregression has no argmax to absorb rounding
    noise: every iteration injects ~1e-5 into the latents and the loop amplifies it (the oracle itself moves by 2e-5 when frame 0's
    latents are perturbed by 1e-6), so the free-running clip is held to a loose bound and the tight check is TEACHER-FORCED: every
    iteration fed the oracle's own step predictions, so each iteration's prediction -- and the final latents of all 31 slots -- are
    compared under identical inputs."""
    if backend != "tc":
        pytest.skip("the MAGE+ branch runs on the tensor-core back end")
    from oracle import mage_oracle as orc
    params = syn.model_params("caterv2plus", frames_length=32)
    sd = syn.make_mage_state_dict(params)
    batch = syn.make_batch(params, 1, seed=1234, text_len=20)
    noise = syn.make_noise(1, seed=99)
    ae = syn.PatchLatentAE(**params["first_stage_config"]["params"])
    z0 = ae.encode(batch["images"][:, 0])
    otr = {}
    want = orc.generate_continuous(sd, z0, batch["text"], batch["speed"], noise, trace=otr)           # [1, 31, 4, 16, 16]
    model = _build(params, sd)
    eng = model.engine()
    cu = lambda t: t.to("cuda")
    free = eng.generate_continuous(cu(z0), cu(batch["text"]), cu(batch["speed"]), cu(noise)).cpu()
    tr = {"force_step_pred": cu(otr["step_pred"])}
    forced = eng.generate_continuous(cu(z0), cu(batch["text"]), cu(batch["speed"]), cu(noise), trace=tr).cpu()
    scale = want.abs().max().item()
    e_step = (tr["step_pred"].cpu() - otr["step_pred"]).abs().flatten(2).max(-1)[0][0]                 # per iteration
    e_forced = (forced - want).abs().max().item()
    e_free = (free - want).abs().max().item()
    print(f"[parity] MAGE+ L=32: teacher-forced per-iteration prediction err max {e_step.max():.2e} (|latent| <= {scale:.2f}), final latents "
          f"(31 slots) {e_forced:.2e}; free-running {e_free:.2e}")
    assert e_step.max().item() <= 3e-5 * max(1.0, scale) and e_forced <= 3e-5 * max(1.0, scale)
    assert e_free <= 5e-3 * max(1.0, scale)


def test_mage_plus_batch_invariance(backend):
    """The MAGE+ branch through the public call: a prompt's clip does not depend on the batch it is generated in (bit for bit),
    including the stand-in first stage (patchify + fp32 matmul: the same function on every device and batch size)."""
    if backend != "tc":
        pytest.skip("the MAGE+ branch runs on the tensor-core back end")
    params = syn.model_params("caterv2plus", frames_length=5)
    sd = syn.make_mage_state_dict(params)
    model = _build(params, sd)
    big = syn.make_batch(params, 6, seed=21, text_len=11)
    noise = syn.make_noise(6, seed=3)
    cu = lambda d: {k: v.to("cuda") for k, v in d.items()}
    v6 = model.autoregressive_generate(cu(big), noise=noise)
    l6 = model.last_latents.clone()
    v2 = model.autoregressive_generate(cu({k: v[2:4] for k, v in big.items()}), noise=noise[2:4])
    assert torch.equal(model.last_latents, l6[2:4]) and torch.equal(v2, v6[2:4])
    host = model.autoregressive_generate({k: v for k, v in big.items()}, noise=noise, to_host=True)
    assert not host.is_cuda and torch.equal(host, v6.cpu())


def test_mage_plus_shipped_config_names_an_external_first_stage():
    """config/mage+_caterv2.yaml keeps the reference's first stage target (latent-diffusion's AutoencoderKL, not vendored): the
    drop-in must fail loudly and helpfully at construction, not later."""
    from mage_b200.config import instantiate_from_config, load_yaml
    cfg = load_yaml(os.path.join(os.path.dirname(GOLDEN_DIR), "..", "config", "mage+_caterv2.yaml"))
    assert cfg["model"]["params"]["use_cids"] is False and cfg["model"]["params"]["ma_config"]["params"]["ln_qkv"] is True
    with pytest.raises(ImportError, match="first stage"):
        instantiate_from_config(cfg["model"])
