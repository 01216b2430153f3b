#!/usr/bin/env python
"""Benchmark of the MAGE sampling path (BASELINE.json metric: generated frames/sec, CATER-v2 128x128x32, batch 64,
on 1/2/4/8 B200 next to the reference's CPU path).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own implementation on the host CPU

A "step" is one `autoregressive_generate` call over one synthetic batch: VQ-VAE encode of frame 0, text / motion-anchor
prelude, L-1 greedy decode steps through the 6 axial blocks, VQ-VAE decode of every generated frame.

Multi-GPU: one process per GPU (torchrun); the GLOBAL prompt batch of the workload (64 for C5) is drawn once
and cut into contiguous per-rank slices -- STRONG scaling, 64/N prompts per GPU, which is the split BASELINE.json quotes
("b64, 1/2/4/8 B200", configs[4] "batch 64, 8xB200 batch-shard").  Prompts are independent, so there is no collective on the
data path; NCCL carries the barrier and the max-over-ranks of the timings only.  `--scaling weak` keeps 64 prompts per GPU;
at N > 1 the strong line also carries a short weak-scaling measurement under "weak_scaling".

Printed by rank 0: ONE JSON line on stdout (NCCL's INFO log goes to stderr).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 for every rank; the only CPU-heavy work here is rank 0's checker / reference legs (the oracle),
# which should use the host's cores (measured: 49 s instead of 3 s for the parity leg with one thread)
if os.environ.get("OMP_NUM_THREADS") == "1" and os.environ.get("RANK", "0") == "0":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import torch  # noqa: E402

UNIT = "frames/s"
TEXT_LEN = 20
# BASELINE.json configs[1..4]; gflop = algorithmic (incremental) FLOPs per generated frame, SURVEY.md §8(d) / BASELINE.md §2
WORKLOADS = {
    "c5": dict(family="caterv2", frames=32, batch=64, gflop=23.58, name="CATER-GEN-v2 128x128x32", cfg="BASELINE.json configs[4]",
               metric="generated frames/sec, CATER-v2 128x128x32 b64"),
    "c4": dict(family="caterv1", frames=16, batch=16, gflop=24.67, name="CATER-GEN-v1 128x128x16", cfg="BASELINE.json configs[3]",
               metric="generated frames/sec, CATER-v1 128x128x16 b16"),
    "c3": dict(family="mnist", frames=20, batch=32, gflop=13.41, name="Double Moving MNIST 64x64x20", cfg="BASELINE.json configs[2]",
               metric="generated frames/sec, Moving-MNIST 64x64x20 b32"),
    "c2": dict(family="mnist", frames=16, batch=1, gflop=13.67, name="Single Moving MNIST 64x64x16", cfg="BASELINE.json configs[1]",
               metric="generated frames/sec, Moving-MNIST 64x64x16 b1"),
    # BASELINE.json configs[4] names the MAGE+ yaml: the same shape through the use_cids=False branch (continuous 4-channel latents,
    # GroupNorm head over all slots => suffix re-evaluation, (L-1)L/2 position passes; stand-in first stage)
    "c5plus": dict(family="caterv2plus", frames=32, batch=64, gflop=None, name="CATER-GEN-v2 MAGE+ 128x128x32 (use_cids=False)",
                   cfg="BASELINE.json configs[4], mage+_caterv2.yaml branch", metric="generated frames/sec, CATER-v2 MAGE+ 128x128x32 b64"),
    # SURVEY.md §8 row N2, forward half: the reference's validation loss (MAGE.forward in eval mode, main_mage.py:163-176) on the C4 shape
    "c4val": dict(family="caterv1", frames=16, batch=16, gflop=None, name="CATER-GEN-v1 128x128x16, MAGE.forward (validation loss, eval mode)",
                  cfg="BASELINE.json configs[3] shape; the stage-2 objective's forward half", val=True,
                  metric="scored frames/sec, MAGE.forward validation loss, CATER-v1 128x128x16 b16"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        smax = int(float(self.rows[0][1])) if self.rows and self.rows[0][1].replace(".", "").isdigit() else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def _workload_inputs(wl, batch, seed=1234, noise_seed=99):
    """Seeded synthetic checkpoint + GLOBAL prompt batch of a workload (every rank builds the same tensors and slices its rows)."""
    from mage_b200 import synthetic as syn
    params = syn.model_params(wl["family"], frames_length=wl["frames"])
    fs = params["first_stage_config"]["params"]
    cb = os.path.join(syn.GOLDEN_DIR, "codebook_f%d.npy" % fs["down_ratio"])
    sd = syn.make_mage_state_dict(params, conditioned=os.path.isfile(cb))
    b = syn.make_batch(params, batch, seed=seed, text_len=TEXT_LEN)
    noise = syn.make_noise(batch, seed=noise_seed) if params["randomness"] else None
    return params, sd, b, noise


# ---------------------------------------------------------------------------------------------- reference arms (CPU / eager GPU)
def _reference_model(params, sd, device="cpu"):
    """The UNMODIFIED reference (oracle/_ref: a git-ignored copy of /root/reference/{modules,utils} made by oracle/build_ref.py,
    or /root/reference itself) behind three import shims, or None when neither is present."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        return None
    m = ref_shims.build_reference_mage(params, sd)
    return m.to(device)


def cpu_reference_sample(wl, threads: int, ar_iters: int = 0, budget_s: float = 60.0):
    """Time the reference's own CPU implementation of the path on ONE prompt of the workload (CPU throughput is flat in batch,
    BASELINE.md §2): `MAGE.autoregressive_generate` of the unmodified reference when a copy is present (kind "reference"),
    else the oracle's restatement in the reference's evaluation order (kind "port").  The whole call is timed -- prelude, all
    L-1 autoregressive iterations (each re-runs the conv and the 6 blocks over all L positions, mage_model.py:673-684), VQ-VAE
    decode -- unless one iteration of the port predicts more than `budget_s` (slow host) or `ar_iters` > 0 is forced: then
    prelude + `ar_iters` identical-cost iterations + decode are timed through the port and extrapolated."""
    from mage_b200 import synthetic as syn
    from oracle import mage_oracle as orc

    torch.set_num_threads(threads)
    L = wl["frames"]
    params, sd, batch, noise = _workload_inputs(wl, 1)
    if not params["use_cids"]:
        ref = _reference_model(params, sd)
        with torch.no_grad():
            t0 = time.perf_counter()
            if ref is not None:
                torch.manual_seed(0)
                ref.autoregressive_generate(batch)
                kind = "reference"
            else:
                ae = syn.PatchLatentAE(**params["first_stage_config"]["params"])
                ae.decode(orc.generate_continuous(sd, ae.encode(batch["images"][:, 0]), batch["text"], batch.get("speed"), noise)[0])
                kind = "port"
            total = time.perf_counter() - t0
        return (L - 1) / total, {"kind": kind, "measured_s": total,
                                 "sample": f"1 prompt of {wl['name']}: one full MAGE.autoregressive_generate call of the "
                                           f"{'unmodified reference' if kind == 'reference' else 'oracle restatement'} (use_cids=False, "
                                           f"stand-in first stage), {total:.1f}s on {threads} threads"}
    fsd = {k[len("first_stage_model."):]: v for k, v in sd.items() if k.startswith("first_stage_model.")}
    with torch.no_grad():
        # one warm iteration of the port: predicts the cost of the full call
        tok0 = orc.vqvae_encode(fsd, batch["images"][:, 0])
        anchor = orc.motion_anchor(sd, tok0, batch["text"], batch.get("speed"), noise)
        inp = orc.embed_tokens(sd, tok0).unsqueeze(1).repeat(1, L - 1, 1, 1, 1)
        orc.flat_axial_decoder(sd, anchor, orc.token_features(sd, inp))
        t0 = time.perf_counter()
        orc.flat_axial_decoder(sd, anchor, orc.token_features(sd, inp))
        t_iter_est = time.perf_counter() - t0
        if ar_iters <= 0 and t_iter_est * (L - 1) <= budget_s:
            ref = _reference_model(params, sd)
            t0 = time.perf_counter()
            if ref is not None:
                if noise is not None:
                    torch.manual_seed(0)
                ref.autoregressive_generate(batch)
                kind = "reference"
            else:
                orc.generate(sd, batch, noise)
                kind = "port"
            total = time.perf_counter() - t0
            what = ("the unmodified reference's MAGE.autoregressive_generate" if kind == "reference" else
                    "the oracle's restatement of MAGE.autoregressive_generate (reference evaluation order)")
            return (L - 1) / total, {"kind": kind, "measured_s": total,
                                     "sample": f"1 prompt of {wl['name']}: one full call of {what} -- prelude, all {L - 1} O(L) "
                                               f"autoregressive iterations, VQ-VAE decode -- {total:.1f}s on {threads} threads"}
        k = max(ar_iters, 3)
        t0 = time.perf_counter()
        tok0 = orc.vqvae_encode(fsd, batch["images"][:, 0])
        anchor = orc.motion_anchor(sd, tok0, batch["text"], batch.get("speed"), noise)
        inp = orc.embed_tokens(sd, tok0).unsqueeze(1).repeat(1, L - 1, 1, 1, 1)
        t_pre = time.perf_counter() - t0
        iters, pred = [], None
        for i in range(k):
            t0 = time.perf_counter()
            pred = orc.flat_axial_decoder(sd, anchor, orc.token_features(sd, inp))
            ids = torch.max(pred, -1)[1]
            inp[:, i + 1] = orc.embed_tokens(sd, ids[:, i])
            iters.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        toks = torch.max(pred, -1)[1]
        orc.vqvae_decode(fsd, toks.view(-1, *toks.shape[-2:]))
        t_dec = time.perf_counter() - t0
    t_iter = sum(iters) / len(iters)
    total = t_pre + (L - 1) * t_iter + t_dec
    return (L - 1) / total, {
        "kind": "port", "measured_s": t_pre + sum(iters) + t_dec,
        "sample": f"1 prompt of {wl['name']}, oracle restatement in the reference's evaluation order: prelude {t_pre:.2f}s + {k} of "
                  f"{L - 1} identical-cost AR iterations (mean {t_iter:.2f}s) + full VQ-VAE decode {t_dec:.2f}s, extrapolated to "
                  f"{total:.1f}s per prompt on {threads} threads (a full call would exceed the bench's time budget on this host)"}


def run_reference(args, wl, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals, secs, info = [], [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, info = cpu_reference_sample(wl, threads, ar_iters=args.ref_iters)
        if i >= args.warmup:
            vals.append(v)
            secs.append(time.perf_counter() - t0)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": wl["metric"], "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * sum(secs) / len(secs), 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl['name']}, global batch {wl['batch']} ({wl['cfg']}); timed on 1 prompt, see cpu_baseline.sample",
                       "family": wl["family"], "frames_length": wl["frames"], "global_batch": wl["batch"], "text_len": TEXT_LEN},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def eager_gpu_baseline(wl, dev, batch_n):
    """The >= 10x target's denominator (BASELINE.json north_star): the reference's eager-PyTorch path on THIS GPU with PyTorch's
    default precision flags (TF32 cuDNN convs, fp32 GEMMs), at a reduced batch (the reference algorithm is O(L^2) per prompt).
    Unmodified reference when a copy is present, else the oracle's restatement of the same algorithm (torch eager ops)."""
    from oracle import mage_oracle as orc
    L = wl["frames"]
    params, sd, batch, noise = _workload_inputs(wl, batch_n, seed=4321, noise_seed=7)
    eb = {k: v.to(dev) for k, v in batch.items()}
    ref = _reference_model(params, sd, dev)
    with torch.no_grad():
        if ref is not None:
            one = {k: v[:1] for k, v in eb.items()}
            ref.autoregressive_generate(one)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref.autoregressive_generate(eb)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            kind = "the unmodified reference (MAGE.autoregressive_generate, torch eager, PyTorch default precision flags) on the same GPU"
        elif not params["use_cids"]:
            raise RuntimeError("no reference copy (oracle/_ref) for the MAGE+ eager-GPU leg")
        else:
            sd_d = {k: v.to(dev) for k, v in sd.items()}
            en = noise.to(dev) if noise is not None else None
            orc.generate(sd_d, {k: v[:1] for k, v in eb.items()}, en[:1] if en is not None else None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            orc.generate(sd_d, eb, en)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            kind = "port of the reference algorithm (every step recomputes all L positions), torch eager on the same GPU"
    del ref
    torch.cuda.empty_cache()
    return {"value": round(batch_n * (L - 1) / dt, 2), "unit": UNIT, "batch": batch_n, "seconds": round(dt, 3), "kind": kind}


# ---------------------------------------------------------------------------------------------- validation loss (MAGE.forward, eval mode)
def _val_inputs(wl, batch, seed=1234, eps_seed=99):
    from mage_b200 import synthetic as syn
    params = syn.model_params(wl["family"], frames_length=wl["frames"])
    fs = params["first_stage_config"]["params"]
    cb = os.path.join(syn.GOLDEN_DIR, "codebook_f%d.npy" % fs["down_ratio"])
    sd = syn.make_mage_state_dict(params, conditioned=os.path.isfile(cb), posterior=True)
    b = syn.make_batch(params, batch, seed=seed, text_len=TEXT_LEN, frames=wl["frames"])
    eps = syn.make_noise(batch, seed=eps_seed) if params["randomness"] else None
    return params, sd, b, eps


def reference_forward_sample(wl, device, batch_n, threads=None):
    """The reference's own MAGE.forward in eval mode, no gradients (what its periodic validation runs, main_mage.py:166-176), on
    `batch_n` clips: frames scored per second.  Unmodified reference when a copy is present, else the oracle's restatement."""
    from oracle import mage_oracle as orc
    if threads:
        torch.set_num_threads(threads)
    L = wl["frames"]
    params, sd, batch, eps = _val_inputs(wl, batch_n, seed=4321, eps_seed=7)
    ref = _reference_model(params, sd, device)
    cuda = torch.device(device).type == "cuda"
    with torch.no_grad():
        if ref is not None:
            eb = {k: v.to(device) for k, v in batch.items()}
            if cuda:
                ref({k: v[:1] for k, v in eb.items()})
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref(eb)
            kind = "reference"
        else:
            t0 = time.perf_counter()
            orc.forward_loss(sd, batch, eps, randomness=params["randomness"], beta=params.get("beta", 1.0), alpha=params.get("alpha", 0.0))
            kind = "port"
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    del ref
    return batch_n * (L - 1) / dt, dt, kind


def run_reference_val(args, wl, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals, secs, kind = [], [], None
    for i in range(args.warmup + args.steps):
        v, dt, kind = reference_forward_sample(wl, "cpu", 1, threads)
        if i >= args.warmup:
            vals.append(v)
            secs.append(dt)
    value = sum(vals) / len(vals)
    sample = (f"1 clip of {wl['name']}: one MAGE.forward call of the {'unmodified reference' if kind == 'reference' else 'oracle restatement'} "
              f"in eval mode under no_grad, {sum(secs) / len(secs):.2f}s on {threads} threads")
    line = {"impl": "reference", "metric": wl["metric"], "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * sum(secs) / len(secs), 1), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl['name']}, global batch {wl['batch']} ({wl['cfg']}); timed on 1 clip", "family": wl["family"],
                       "frames_length": wl["frames"], "global_batch": wl["batch"], "text_len": TEXT_LEN},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_cuda_val(args, wl, rank, world, local_rank):
    """`--workload c4val`: a step is one MAGE.forward call (eval mode) over the rank's clips -- VQ-VAE encode of all L frames, 3-D conv
    posterior, motion anchor, teacher-forced decoder in full-sequence form, cross-entropy / KL -- returning the reference's
    (final_loss, loss_dict).  Clips shard over the GPUs like prompts; the loss is then the all_reduce of main_mage.py:177-180."""
    import torch.distributed as dist

    from mage_b200 import ops, shard
    from mage_b200.config import instantiate_from_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = os.environ.get("MAGE_NCCL_DEBUG", "INFO")
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    L = wl["frames"]
    G = args.batch if args.batch > 0 else wl["batch"]
    params, sd, gbatch, geps = _val_inputs(wl, G)
    lo, hi = shard.shard_bounds(G, world, rank)
    B = hi - lo
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": dict(params, with_posterior=True)})
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    host = {k: v[lo:hi].contiguous().pin_memory() for k, v in gbatch.items()}
    eps_h = geps[lo:hi].contiguous().pin_memory() if geps is not None else None
    dbatch = {k: v.to(dev) for k, v in host.items()}
    deps = eps_h.to(dev) if eps_h is not None else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allred(x, op):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        model(dbatch, eps=deps)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        final, loss_dict = model(dbatch, eps=deps)
    e1.record()
    barrier()
    ms_max = allred(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
    clocks = sampler.stop() if rank == 0 else None
    launches = ops.launch_count() - n0
    frames_step = allred(B * (L - 1), dist.ReduceOp.SUM if world > 1 else None)
    value = frames_step * args.steps / (ms_max / 1e3)
    test_loss = shard.validation_loss(lambda b: model(b, eps=deps), [dbatch], dev)     # per-rank mean -> all_reduce / world
    # end to end: the clips start in pinned host memory, the losses come back as Python floats
    model(host, eps=eps_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model(host, eps=eps_h)
    barrier()
    e2e_value = frames_step * args.steps / allred(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + (eps_h.numel() * 4 if eps_h is not None else 0)
    if rank == 0:
        ops.PROFILE = []
        model(dbatch, eps=deps)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for kind, flops, a, b in prof:
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += flops; d[1] += a.elapsed_time(b); d[2] += 1
        peaks, how = _peaks()
        peak = peaks["bf16_tflops_sustained"]
        fl = sum(agg.get(k, [0.0, 0.0, 0])[0] for k in ("gemm", "conv"))
        ms = sum(agg.get(k, [0.0, 0.0, 0])[1] for k in ("gemm", "conv"))
        ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "tensor-core classes of the pass: full-sequence decoder GEMMs (tc_gemm_kernel, fused QKV + axial "
                                             "attention) and the posterior's 3x3x3 convolutions as implicit GEMMs over frame triples (tc_conv_halo_kernel)",
                "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": None,
                "peak_source": how + ", dense bf16 sustained; 3 fp16 MMAs per product (fp32-grade)", "own_ceiling_frac": round(3 * ach / peak, 4),
                "breakdown_ms_per_step": {k: {"ms_per_step": round(v[1], 2), "launches": v[2], "tflops": round(v[0] / max(v[1], 1e-9) / 1e9, 1)}
                                          for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
                "how": "sum of algorithmic FLOPs / sum of CUDA-event durations over every launch of the two classes in one step of rank 0"}
        parity = eager = cpu = None
        if not args.no_parity:
            from oracle import mage_oracle as orc
            torch.set_num_threads(max(1, (os.cpu_count() or 1) - (world - 1)))
            t0 = time.perf_counter()
            sub = {k: v[:1] for k, v in host.items()}
            got = model({k: v.to(dev) for k, v in sub.items()}, eps=deps[:1] if deps is not None else None)[1]
            want = orc.forward_loss(sd, sub, eps_h[:1] if eps_h is not None else None, randomness=params["randomness"],
                                    beta=params.get("beta", 1.0), alpha=params.get("alpha", 0.0))
            parity = {"rows": 1, "got": {k.split("/")[1]: v for k, v in got.items()}, "oracle": want,
                      "rel_err": {k: abs(got["val/" + k] - want[k]) / abs(want[k]) for k in want},
                      "oracle_seconds": round(time.perf_counter() - t0, 1),
                      "what": "MAGE.forward of clip 0 of the timed batch alone vs oracle.forward_loss (CPU fp32, same draw)"}
        if args.eager_gpu > 0:
            model.invalidate()
            torch.cuda.empty_cache()
            try:
                v, dt, kind = reference_forward_sample(wl, dev, G)
                eager = {"value": round(v, 2), "unit": UNIT, "batch": G, "seconds": round(dt, 3),
                         "kind": ("the unmodified reference" if kind == "reference" else "oracle restatement") +
                                 " (MAGE.forward, eval, no_grad, torch eager, PyTorch default precision flags) on the same GPU",
                         "ratio_resident_per_gpu": round(value / world / v, 2)}
            except Exception as e:
                eager = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_cpu:
            v, dt, kind = reference_forward_sample(wl, "cpu", 1, os.cpu_count() or 1)
            cpu = {"value": round(v, 4), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                   "sample": f"1 clip of {wl['name']}: one MAGE.forward call (eval, no_grad), {dt:.2f}s on {os.cpu_count() or 1} threads"}
        line = {"metric": wl["metric"], "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{wl['name']}, global batch {G} ({wl['cfg']})", "family": wl["family"], "frames_length": L,
                           "global_batch": G, "batch_per_gpu": B, "text_len": TEXT_LEN,
                           "parallelism": f"clip-shard x{world}; one all_reduce of the per-rank mean loss", "cuda_graph": False,
                           "l2": "working set >> L2 (full-sequence activations %.1f GB per tensor); no explicit flush" % (L * B * 256 * 2048 * 4 / 1e9)},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12},
                "gpu_launches": int(launches), "kernels_per_step": int(launches // args.steps), "roofline": roof, "cpu_baseline": cpu,
                "clocks": clocks, "parity": parity, "test_loss": test_loss, "loss_dict": loss_dict}
        if eager is not None:
            line["eager_gpu_baseline"] = eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- this repo's CUDA path
def _time_resident(eng, dev_in, steps, barrier):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        eng.generate(*dev_in)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def run_cuda(args, wl, rank, world, local_rank):
    import torch.distributed as dist

    from mage_b200 import ops, shard
    from mage_b200.config import instantiate_from_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's INFO log (ranks, transports) goes to stderr; stdout carries the one JSON line and nothing else (NCCL prints to
        # stdout unless a debug file is named, and the image presets NCCL_DEBUG=VERSION)
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = os.environ.get("MAGE_NCCL_DEBUG", "INFO")
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    os.environ["MAGE_BACKEND"] = args.backend
    L = wl["frames"]
    G = args.batch if args.batch > 0 else wl["batch"]                      # global batch
    weak = args.scaling == "weak"
    params, sd, gbatch, gnoise = _workload_inputs(wl, G * world if weak else G)
    lo, hi = shard.shard_bounds(G * world if weak else G, world, rank)
    B = hi - lo
    assert B > 0, f"rank {rank} has no prompts (global batch {G} over {world} GPUs)"
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": params})
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    eng = model.engine()
    batch = {k: v[lo:hi] for k, v in gbatch.items()}
    noise = gnoise[lo:hi] if gnoise is not None else None
    host = {k: v.contiguous().pin_memory() for k, v in batch.items()}
    host["images"] = batch["images"][:, 0:1].contiguous().pin_memory()
    noise_h = noise.contiguous().pin_memory() if noise is not None else None
    dev_in = (host["images"][:, 0].to(dev), host["text"].to(dev), host["speed"].to(dev), noise_h.to(dev) if noise_h is not None else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        eng.generate(*dev_in)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_max = max_ranks(_time_resident(eng, dev_in, args.steps, barrier))
    clocks = sampler.stop() if rank == 0 else None
    kernels_per_step = eng.kernels_per_generate
    frames_step = sum_ranks(B * (L - 1))                   # generated frames of one step, all ranks
    value = frames_step * args.steps / (ms_max / 1e3)
    plan = eng._plan(B)
    parity_tokens = None
    if rank == 0 and not args.no_parity:
        _, ptok, ptok0 = eng.generate(*dev_in)
        parity_tokens = (ptok[:min(2, B)].cpu(), ptok0[:min(2, B)].cpu())

    # ---- end to end through the public API: pinned host inputs -> H2D -> generate -> D2H of the video
    hb = {"images": host["images"], "text": host["text"], "speed": host["speed"]}
    out_shape = (B, L, *host["images"].shape[2:])

    def step_e2e():
        # public API: pinned host inputs -> H2D -> generate -> the clip back on the host (frames stream out as they are decoded)
        video = model.autoregressive_generate(hb, noise=noise_h, to_host=True)
        assert not video.is_cuda and tuple(video.shape) == out_shape

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_value = frames_step * args.steps / max_ranks(time.perf_counter() - t0)
    h2d = sum(v.numel() * v.element_size() for v in hb.values()) + (noise_h.numel() * 4 if noise_h is not None else 0)
    d2h = 4
    for d in out_shape:
        d2h *= d

    # ---- at N > 1 on the strong split: the weak-scaling figure (global batch per GPU) as an extra key
    weak_extra = None
    if world > 1 and not weak and not args.no_weak:
        wparams, _, wb, wn = _workload_inputs(wl, G, seed=1234 + rank, noise_seed=99 + rank)
        w_in = (wb["images"][:, 0].to(dev), wb["text"].to(dev), wb["speed"].to(dev), wn.to(dev) if wn is not None else None)
        for _ in range(2):
            eng.generate(*w_in)
        wms = max_ranks(_time_resident(eng, w_in, 3, barrier))
        weak_extra = {"value": round(world * G * (L - 1) * 3 / (wms / 1e3), 2), "unit": UNIT, "batch_per_gpu": G, "steps": 3}

    # ---- roofline of the dominant kernel class (dense GEMM of the axial blocks), CUDA events around
    #      every launch of one extra eager step on the launching stream
    roof = None
    if rank == 0:
        eng.use_cuda_graph = False
        saved = (eng.overlap_decode, eng.n_streams)
        eng.overlap_decode, eng.n_streams = False, 1   # per-kernel durations are taken with the launches serialised on one stream
        ops.PROFILE = []
        eng.generate(*dev_in)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        eng.use_cuda_graph = True
        eng.overlap_decode, eng.n_streams = saved
        agg = {}
        for kind, flops, a, b in prof:
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += flops
            d[1] += a.elapsed_time(b)
            d[2] += 1
        breakdown = {k: {"ms_per_step": round(v[1], 2), "launches": v[2]} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        ta = agg.get("temporal_attn")
        if ta:
            breakdown["temporal_attn"]["achieved_gbs"] = round(ta[0] / (ta[1] * 1e-3) / 1e9, 1)
        peaks, how = _peaks()
        peak = peaks["bf16_tflops_sustained"]
        g = agg.get("gemm", [0.0, 1.0, 1])
        c = agg.get("conv", [0.0, 1.0, 1])
        ach = g[0] / (g[1] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get("gemm_bytes_per_launch")
        per_gpu_fps = value / world
        roof = {"bound": "tensor", "kernel": "dense GEMM (axial-block linears: QKV, out-proj, MLP, head) -- " +
                ("tc_gemm_kernel: CTA-pair tiles (tcgen05 cta_group::2), kind::f16 MMAs on fp16 hi/lo split operands, "
                 "3 MMAs per product (fp32-grade), TMA loads + TMA-store epilogue" if args.backend == "tc" else "fp32 FFMA"),
                "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic,
                "peak_source": how + ", dense bf16 sustained; token parity needs fp32-grade GEMMs: 3 fp16 MMAs per product, so the kernel's own ceiling is peak/3",
                "launches_per_step": g[2], "gflop_per_launch_avg": round(g[0] / g[2] / 1e9, 3), "ms_in_kernel_per_step": round(g[1], 2),
                "conv_implicit_gemm": {"achieved": round(c[0] / (c[1] * 1e-3) / 1e12, 2), "launches_per_step": c[2], "ms_per_step": round(c[1], 2)},
                "whole_step_algorithmic": {"achieved": round(per_gpu_fps * wl["gflop"] / 1e3, 2), "unit": "TFLOP/s",
                                           "frac": round(per_gpu_fps * wl["gflop"] / 1e3 / peak, 4),
                                           "note": "SURVEY.md §8(d) FLOPs per generated frame; the token 3x3 conv + in_linear (1.34 GFLOP "
                                                   "per frame, 5.7 %) are counted although they run as per-code table lookups here"},
                "own_ceiling_frac": round(3.0 * ach / peak, 4) if args.backend == "tc" else None,
                "breakdown_ms_per_step": breakdown,
                "how": "sum of algorithmic FLOPs / sum of CUDA-event durations over every launch of the kernel class in one eager "
                       "single-stream step of rank 0; own_ceiling_frac counts the 3 MMAs issued per product"}

    eager = None
    if rank == 0 and args.eager_gpu > 0:
        del eng
        model.invalidate()
        torch.cuda.empty_cache()
        try:
            eager = eager_gpu_baseline(wl, dev, args.eager_gpu)
            eager["ratio_resident_per_gpu"] = round(value / world / eager["value"], 1)
            eager["ratio_e2e_per_gpu"] = round(e2e_value / world / eager["value"], 1)
        except Exception as e:   # context only: never lose the bench line over it
            eager = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- parity evidence for the timed configuration: rows 0-1 of this rank's shard against the CPU oracle
    # (runs after every timed region: the other ranks are past their last barrier, so they do not spin next to the oracle's threads)
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import mage_oracle as orc
        n = min(2, B)
        tokens, tok0 = parity_tokens
        otr = {}
        t0 = time.perf_counter()
        torch.set_num_threads(max(1, (os.cpu_count() or 1) - (world - 1)))
        orc.generate_incremental(sd, {k: v[:n] for k, v in batch.items()}, noise[:n] if noise is not None else None, otr)
        neq = tokens != otr["tokens"]
        gap = otr["gap"].reshape(neq.shape)
        first_bad = [int(torch.nonzero(neq[b].flatten(1).any(1))[0]) if neq[b].any() else None for b in range(n)]
        parity = {"rows": n, "positions": int(neq.numel()), "token_mismatches": int(neq.sum()),
                  "mismatches_at_reference_gap_ge_5e-5": int((neq & (gap >= 5e-5)).sum()),
                  "first_frame_vq_index_mismatches": int((tok0 != otr["tok0"].reshape(tok0.shape)).sum()),
                  "first_diverging_frame_per_row": first_bad, "oracle_seconds": round(time.perf_counter() - t0, 1),
                  "what": "free-running greedy tokens of rows 0..%d of the timed batch vs oracle.generate_incremental (CPU fp32); a flip at a "
                          "reference near-tie cascades into the later frames of that row (tests/ re-check those teacher-forced)" % (n - 1)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, info = cpu_reference_sample(wl, os.cpu_count() or 1, ar_iters=args.ref_iters)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": info["kind"], "sample": info["sample"]}

    if rank == 0:
        S, Gd = plan
        line = {"metric": wl["metric"], "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{wl['name']}, global batch {G * world if weak else G} ({wl['cfg']})", "family": wl["family"],
                           "frames_length": L, "global_batch": G * world if weak else G, "batch_per_gpu": B, "text_len": TEXT_LEN,
                           "parallelism": f"prompt-shard x{world}, no data-path collective", "backend": args.backend,
                           "cuda_graph": True, "chunk_streams": S, "decode_group_frames": Gd,
                           "l2": "working set >> L2 (K/V cache %.1f GB, decoder activations > 126 MB per layer); no explicit flush" %
                                 (2 * 2 * B * 256 * L * 512 * 4 / 1e9)},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(kernels_per_step * args.steps), "kernels_per_step": int(kernels_per_step),
                "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "parity": parity}
        if weak_extra is not None:
            line["weak_scaling"] = weak_extra
        if eager is not None:
            line["eager_gpu_baseline"] = eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cuda_plus(args, wl, rank, world, local_rank):
    """MAGE+ branch (use_cids=False): the public call is the step -- first-stage encode (plain torch module), the CUDA path between
    the two first-stage calls (`SamplerEngine.generate_continuous`), first-stage decode.  Same line format as run_cuda."""
    import torch.distributed as dist

    from mage_b200 import ops, shard
    from mage_b200.config import instantiate_from_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = os.environ.get("MAGE_NCCL_DEBUG", "INFO")
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    L = wl["frames"]
    G = args.batch if args.batch > 0 else wl["batch"]
    params, sd, gbatch, gnoise = _workload_inputs(wl, G)
    lo, hi = shard.shard_bounds(G, world, rank)
    B = hi - lo
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": params})
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    batch = {k: v[lo:hi] for k, v in gbatch.items()}
    batch["images"] = batch["images"][:, 0:1].contiguous()
    noise = gnoise[lo:hi].contiguous()
    host = {k: v.contiguous().pin_memory() for k, v in batch.items()}
    noise_h = noise.pin_memory()
    dbatch = {k: v.to(dev) for k, v in host.items()}
    dnoise = noise_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allred(x, op):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        model.autoregressive_generate(dbatch, noise=dnoise)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        video = model.autoregressive_generate(dbatch, noise=dnoise)
    e1.record()
    barrier()
    ms_max = allred(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
    clocks = sampler.stop() if rank == 0 else None
    launches = ops.launch_count() - n0
    frames_step = allred(B * (L - 1), dist.ReduceOp.SUM if world > 1 else None)
    value = frames_step * args.steps / (ms_max / 1e3)
    lat = model.last_latents[:1].cpu()
    # end to end: pinned host inputs -> H2D -> generate -> the clip in pinned host memory
    model.autoregressive_generate(host, noise=noise_h, to_host=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = model.autoregressive_generate(host, noise=noise_h, to_host=True)
        assert not out.is_cuda
    barrier()
    e2e_value = frames_step * args.steps / allred(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + noise_h.numel() * 4
    d2h = out.numel() * 4
    roof = eager = cpu = parity = None
    if rank == 0:
        ops.PROFILE = []
        model.autoregressive_generate(dbatch, noise=dnoise)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for kind, flops, a, b in prof:
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += flops; d[1] += a.elapsed_time(b); d[2] += 1
        peaks, how = _peaks()
        peak = peaks["bf16_tflops_sustained"]
        g = agg.get("gemm", [0.0, 1.0, 1])
        ach = g[0] / (g[1] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "dense GEMM class of the full-sequence block path (tc_gemm_kernel, fused QKV + axial attention)",
                "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": None,
                "peak_source": how + ", dense bf16 sustained; 3 fp16 MMAs per product (fp32-grade)", "own_ceiling_frac": round(3 * ach / peak, 4),
                "launches_per_step": g[2], "breakdown_ms_per_step": {k: {"ms_per_step": round(v[1], 2), "launches": v[2]}
                                                                      for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
                "how": "sum of algorithmic FLOPs / sum of CUDA-event durations over every launch of the class in one step of rank 0"}
        if not args.no_parity:
            from oracle import mage_oracle as orc
            from mage_b200 import synthetic as syn
            torch.set_num_threads(max(1, (os.cpu_count() or 1) - (world - 1)))
            t0 = time.perf_counter()
            ae = syn.PatchLatentAE(**params["first_stage_config"]["params"])
            want = orc.generate_continuous(sd, ae.encode(batch["images"][:1, 0]), batch["text"][:1], batch["speed"][:1], noise[:1])
            err = float((lat - want).abs().max())
            parity = {"rows": 1, "latent_max_abs_err": err, "latent_max_abs": float(want.abs().max()), "slots": L - 1,
                      "oracle_seconds": round(time.perf_counter() - t0, 1),
                      "what": "predicted latents of row 0 of the timed batch vs oracle.generate_continuous (CPU fp32, reference evaluation order)"}
        if args.eager_gpu > 0:
            model.invalidate()
            torch.cuda.empty_cache()
            try:
                eager = eager_gpu_baseline(wl, dev, args.eager_gpu)
                eager["ratio_resident_per_gpu"] = round(value / world / eager["value"], 1)
            except Exception as e:
                eager = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_cpu:
            v, info = cpu_reference_sample(wl, os.cpu_count() or 1)
            cpu = {"value": round(v, 4), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": info["kind"], "sample": info["sample"]}
        line = {"metric": wl["metric"], "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{wl['name']}, global batch {G} ({wl['cfg']})", "family": wl["family"], "frames_length": L,
                           "global_batch": G, "batch_per_gpu": B, "text_len": TEXT_LEN, "parallelism": f"prompt-shard x{world}, no data-path collective",
                           "first_stage": "mage_b200.synthetic.PatchLatentAE (stand-in torch module; the shipped AutoencoderKL is not vendored)",
                           "position_passes": (L - 1) * L // 2, "cuda_graph": False,
                           "l2": "working set >> L2 (suffix activations up to %.1f GB per tensor); no explicit flush" % ((L - 1) * B * 256 * 2048 * 4 / 1e9)},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "kernels_per_step": int(launches // args.steps), "roofline": roof, "cpu_baseline": cpu,
                "clocks": clocks, "parity": parity}
        if eager is not None:
            line["eager_gpu_baseline"] = eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS), help="BASELINE.json configs[1..4] (default c5 = the metric's)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the workload's global batch is sharded over the GPUs (BASELINE's split); weak: that batch per GPU")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's global batch (strong) / per-GPU batch (weak)")
    ap.add_argument("--backend", default=os.environ.get("MAGE_BACKEND", "tc"), choices=["tc", "simt"],
                    help="tc: tcgen05 tensor cores on split-fp16 operands (fp32-grade); simt: fp32 FFMA kernels")
    ap.add_argument("--ref-iters", type=int, default=0, help="force the extrapolated CPU sample with this many AR iterations (0 = full call)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle token check of the timed batch's first rows")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling measurement at N > 1")
    ap.add_argument("--eager-gpu", type=int, default=8, metavar="B",
                    help="time the reference's eager-PyTorch path on this GPU at batch B (the >=10x target's denominator; 0 = skip)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        (run_reference_val if wl.get("val") else run_reference)(args, wl, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if wl.get("val"):
        run_cuda_val(args, wl, rank, world, local_rank)
    elif wl["family"] == "caterv2plus":
        run_cuda_plus(args, wl, rank, world, local_rank)
    else:
        run_cuda(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
