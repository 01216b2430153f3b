#!/usr/bin/env python
"""Benchmark of the MAGE sampling path (BASELINE.json metric: generated frames/sec,
CATER-v2 128x128x32, batch 64 per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

A "step" is one `autoregressive_generate` call over one synthetic batch: VQ-VAE encode of frame 0,
text / motion-anchor prelude, L-1 greedy decode steps through the 6 axial blocks, VQ-VAE decode of
every generated frame.  One process per GPU (torchrun for N > 1); prompts are independent, so the
batch shards with no collective on the data path (NCCL only for the barrier / max-over-ranks of the
timings) -- weak scaling, 64 prompts per GPU.

Printed by rank 0: ONE JSON line (see README of the contract in DESIGN.md §Measurement).
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

import torch  # noqa: E402

METRIC = "generated frames/sec, CATER-v2 128x128x32 b64"
UNIT = "frames/s"
FAMILY, FRAMES, BATCH, TEXT_LEN = "caterv2", 32, 64, 20
GFLOP_PER_FRAME = 23.58  # algorithmic (incremental) FLOPs per generated frame at C5, SURVEY.md §8(d)


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


def cpu_reference_sample(threads: int, ar_iters: int = 3):
    """Time the reference algorithm (oracle/mage_oracle.py, reference evaluation order -- every step
    re-runs the conv and the 6 blocks over all L positions, mage_model.py:673-684) on the host CPU
    for ONE prompt of the C5 workload.  Bounded sample: prelude + `ar_iters` of the L-1 (identical-cost)
    autoregressive iterations + the full VQ-VAE decode; frames/s = (L-1) / (prelude + (L-1)*mean_iter + decode)."""
    from mage_b200 import synthetic as syn
    from oracle import mage_oracle as orc

    torch.set_num_threads(threads)
    params = syn.model_params(FAMILY, frames_length=FRAMES)
    sd = syn.make_mage_state_dict(params, conditioned=os.path.isfile(os.path.join(syn.GOLDEN_DIR, "codebook_f8.npy")))
    batch = syn.make_batch(params, 1, seed=1234, text_len=TEXT_LEN)
    noise = syn.make_noise(1)
    fsd = {k[len("first_stage_model."):]: v for k, v in sd.items() if k.startswith("first_stage_model.")}
    L = FRAMES
    with torch.no_grad():
        t0 = time.perf_counter()
        tok0 = orc.vqvae_encode(fsd, batch["images"][:, 0])
        anchor = orc.motion_anchor(sd, tok0, batch["text"], batch.get("speed"), noise)
        inp = orc.embed_tokens(sd, tok0).unsqueeze(1).repeat(1, L - 1, 1, 1, 1)
        t_pre = time.perf_counter() - t0
        iters = []
        pred = None
        for i in range(ar_iters):
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
        "sample": f"1 prompt of the C5 workload (CATER-v2 128x128x32), reference evaluation order: prelude {t_pre:.2f}s + "
                  f"{ar_iters} of {L - 1} identical-cost AR iterations (mean {t_iter:.2f}s) + full 31-frame VQ-VAE decode {t_dec:.2f}s; "
                  f"extrapolated to {total:.1f}s per prompt (CPU throughput is flat in batch, BASELINE.md)",
        "measured_s": t_pre + sum(iters) + t_dec}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, info = cpu_reference_sample(threads, ar_iters=args.ref_iters)
        if i >= args.warmup:
            vals.append(v)
            secs.append(time.perf_counter() - t0)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * sum(secs) / len(secs), 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"CATER-GEN-v2 128x128x{FRAMES}, batch {BATCH} per GPU (timed on 1 prompt, see cpu_baseline.sample)",
                       "family": FAMILY, "frames_length": FRAMES, "batch_per_gpu": BATCH, "text_len": TEXT_LEN},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": info["sample"]},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_cuda(args, rank, world, local_rank):
    import torch.distributed as dist

    from mage_b200 import ops, synthetic as syn
    from mage_b200.config import instantiate_from_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("MAGE_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    os.environ["MAGE_BACKEND"] = args.backend
    B, L = args.batch, args.frames
    params = syn.model_params(FAMILY, frames_length=L)
    sd = syn.make_mage_state_dict(params, conditioned=os.path.isfile(os.path.join(syn.GOLDEN_DIR, "codebook_f8.npy")))
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": params})
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    eng = model.engine()
    # every rank gets its own shard of the global prompt batch (seeded by rank): no data-path collective
    batch = syn.make_batch(params, B, seed=1234 + rank, text_len=TEXT_LEN)
    noise = syn.make_noise(B, seed=99 + rank)
    host = {k: v.pin_memory() for k, v in batch.items()}
    host["images"] = batch["images"][:, 0:1].contiguous().pin_memory()
    noise_h = noise.pin_memory()
    d_img, d_txt, d_spd, d_noise = host["images"][:, 0].to(dev), host["text"].to(dev), host["speed"].to(dev), noise_h.to(dev)

    def step_resident():
        return eng.generate(d_img, d_txt, d_spd, d_noise)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    kernels_per_step = eng.kernels_per_generate
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    frames_total = world * B * (L - 1) * args.steps
    value = frames_total / (ms_max / 1e3)

    # ---- end to end through the public API: pinned host inputs -> H2D -> generate -> D2H of the video
    out_host = torch.empty(B, L, *host["images"].shape[2:], dtype=torch.float32).pin_memory()
    hb = {"images": host["images"], "text": host["text"], "speed": host["speed"]}

    def step_e2e():
        # public API: pinned host inputs -> H2D -> generate -> the clip back on the host (frames stream out as they are decoded)
        video = model.autoregressive_generate(hb, noise=noise_h, to_host=True)
        assert not video.is_cuda and tuple(video.shape) == tuple(out_host.shape)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = frames_total / float(e2e_s.item())
    h2d = sum(v.numel() * v.element_size() for v in hb.values()) + noise_h.numel() * 4
    d2h = out_host.numel() * 4

    # ---- roofline of the dominant kernel class (dense GEMM of the axial blocks), CUDA events around
    #      every launch of one extra eager step on the launching stream
    roof = None
    if rank == 0:
        eng.use_cuda_graph = False
        overlap, eng.overlap_decode = eng.overlap_decode, False   # per-kernel durations are taken with the launches serialised
        ops.PROFILE = []
        step_resident()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        eng.use_cuda_graph = True
        eng.overlap_decode = overlap
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
        roof = {"bound": "tensor", "kernel": "dense GEMM (axial-block linears: QKV, out-proj, MLP, head) -- " +
                ("tc_gemm_kernel<128,2>: CTA-pair 256x128 tiles (tcgen05 cta_group::2), kind::f16 MMAs on fp16 hi/lo split operands, "
                 "3 MMAs per product (fp32-grade), TMA loads + TMA-store epilogue" if args.backend == "tc" else "fp32 FFMA"),
                "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic,
                "peak_source": how + ", dense bf16 sustained; token parity needs fp32-grade GEMMs: 3 fp16 MMAs per product, so the kernel's own ceiling is peak/3",
                "launches_per_step": g[2], "gflop_per_launch_avg": round(g[0] / g[2] / 1e9, 3), "ms_in_kernel_per_step": round(g[1], 2),
                "conv_implicit_gemm": {"achieved": round(c[0] / (c[1] * 1e-3) / 1e12, 2), "launches_per_step": c[2], "ms_per_step": round(c[1], 2)},
                "whole_step_algorithmic": {"achieved": round(value / world * GFLOP_PER_FRAME / 1e3, 2), "unit": "TFLOP/s",
                                           "frac": round(value / world * GFLOP_PER_FRAME / 1e3 / peak, 4)},
                "own_ceiling_frac": round(3.0 * ach / peak, 4) if args.backend == "tc" else None,
                "breakdown_ms_per_step": breakdown,
                "how": "sum of algorithmic FLOPs / sum of CUDA-event durations over every launch of the kernel class in one eager step; "
                       "own_ceiling_frac counts the 3 MMAs issued per product"}

    eager = None
    overlap_flag = bool(eng.overlap_decode)
    if rank == 0 and args.eager_gpu > 0:
        # context for the >=10x target of BASELINE.json: the reference algorithm (reference evaluation order, no KV cache) as
        # plain PyTorch eager ops on this same GPU, PyTorch's default precision flags, at a reduced batch (it is O(L^2))
        from oracle import mage_oracle as orc
        overlap_flag = bool(eng.overlap_decode)
        del eng
        model._engine = None
        torch.cuda.empty_cache()
        Be = args.eager_gpu
        sd_d = {k: v.to(dev) for k, v in sd.items()}
        eb = {k: v.to(dev) for k, v in syn.make_batch(params, Be, seed=4321, text_len=TEXT_LEN).items()}
        en = syn.make_noise(Be, seed=7).to(dev)
        with torch.no_grad():
            orc.generate(sd_d, {k: v[:1] for k, v in eb.items()}, en[:1])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            orc.generate(sd_d, eb, en)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        eager = {"value": round(Be * (L - 1) / dt, 2), "unit": UNIT, "batch": Be, "seconds": round(dt, 3),
                 "kind": "port of the reference algorithm (every step recomputes all L positions), torch eager on the same GPU"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, info = cpu_reference_sample(os.cpu_count() or 1, ar_iters=args.ref_iters)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": info["sample"]}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"CATER-GEN-v2 128x128x{L}, batch {B} per GPU (BASELINE.json configs[4])", "family": FAMILY,
                           "frames_length": L, "batch_per_gpu": B, "global_batch": B * world, "text_len": TEXT_LEN,
                           "parallelism": f"prompt-shard x{world}, no data-path collective", "backend": args.backend,
                           "cuda_graph": True, "decode_overlap_stream": overlap_flag,
                           "l2": "working set >> L2 (K/V cache %.1f GB, decoder activations >1 GB per tensor); no explicit flush" %
                                 (2 * 2 * B * 256 * L * 512 * 4 / 1e9)},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(kernels_per_step * args.steps), "kernels_per_step": int(kernels_per_step),
                "roofline": roof, "cpu_baseline": cpu, "clocks": clocks}
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
    ap.add_argument("--batch", type=int, default=BATCH, help="prompts per GPU")
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--backend", default=os.environ.get("MAGE_BACKEND", "tc"), choices=["tc", "simt"],
                    help="tc: tcgen05 tensor cores on split-fp16 operands (fp32-grade); simt: fp32 FFMA kernels")
    ap.add_argument("--ref-iters", type=int, default=3, help="AR iterations timed per CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--eager-gpu", type=int, default=0, metavar="B",
                    help="also time the reference algorithm as torch eager ops on this GPU at batch B (context only, off by default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
