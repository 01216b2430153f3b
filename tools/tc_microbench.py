#!/usr/bin/env python
"""Micro-benchmark of the tensor-core GEMM / conv kernel on the shapes of the C5 workload
(B=64: M = 16384 token rows per decode step; VQ-VAE decoder maps for 64 frames).
Prints achieved fp32-grade TFLOP/s per shape (CUDA events, L2 flushed between iterations)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mage_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--nsplit", type=int, default=1, help="ops.tc_nsplit mode (0 off, 1 auto, 2 force)")
    ap.add_argument("--scan", action="store_true", help="K / N scans that separate per-tile from per-k-block cost")
    ap.add_argument("--cfgs", default="0,-1", help="';'-separated bn,pair tile overrides (ops.tc_tuning) to sweep")
    ap.add_argument("--passes", type=int, default=3, help="convolutions: 3 = fp32-grade products, 1 = single-pass fp16 (decoder budget)")
    ap.add_argument("--rows", type=int, default=16384, help="M of the GEMM cases (16384 = 64 prompts, 2048 = 8 prompts)")
    args = ap.parse_args()
    flush = None if args.no_flush else torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
    M = args.rows
    g = torch.Generator(device="cpu").manual_seed(0)
    rows = []

    def gemm_case(name, M, N, K, mode):
        a = ops.split(torch.randn(M, K, generator=g).to(DEV))
        w = ops.split((torch.randn(N, K, generator=g) * K ** -0.5).to(DEV))
        b = torch.randn(N, generator=g).to(DEV)
        x = torch.randn(M, N, generator=g).to(DEV)
        out = torch.empty(M, N, device=DEV)
        sp = torch.empty(2, M, N, device=DEV, dtype=torch.float16)
        if mode == "f32":
            fn = lambda: ops.gemm_tc(a, w, b, out=out)
        elif mode == "split":
            fn = lambda: ops.gemm_tc(a, w, b, want=(), out_split=sp, act=2)
        elif mode == "res":
            fn = lambda: ops.gemm_tc(a, w, b, residual=x, out=x)
        ms = timeit(fn, args.iters, flush)
        rows.append((name, 2.0 * M * N * K / ms / 1e9, ms))

    def conv_case(name, n, H, Cin, Cout, k, mode="split"):
        x = ops.split(torch.randn(n, H, H, Cin, generator=g).to(DEV))
        w = ops.split((torch.randn(Cout, k, k, Cin, generator=g) * (k * k * Cin) ** -0.5).to(DEV))
        b = torch.randn(Cout, generator=g).to(DEV)
        sp = torch.empty(2, n, H, H, Cout, device=DEV, dtype=torch.float16)
        out = torch.empty(n, H, H, Cout, device=DEV) if mode == "f32" else None
        fn = (lambda: ops.conv2d_tc(x, w, b, pad=(k // 2, k // 2), act=1, want=(), out_split=sp, passes=args.passes)) if mode == "split" else \
             (lambda: ops.conv2d_tc(x, w, b, pad=(k // 2, k // 2), act=1, out=out, passes=args.passes))
        ms = timeit(fn, args.iters, flush)
        rows.append((name, 2.0 * n * H * H * Cout * k * k * Cin / ms / 1e9, ms))

    def head_case(name, n, H):
        x = ops.split(torch.randn(n, H, H, 64, generator=g).to(DEV))
        w = ops.split((torch.randn(256, 3, 3, 64, generator=g) * 576 ** -0.5).to(DEV))
        b = torch.randn(256, generator=g).to(DEV)
        r = torch.randn(n, H // 2, H // 2, 256, generator=g).to(DEV)
        hw, hb = (torch.randn(3, 256, generator=g) / 16).to(DEV), torch.randn(3, generator=g).to(DEV)
        out = torch.empty(n, 3, H, H, device=DEV)
        fn = lambda: ops.conv2d_tc_pixel_head(x, w, b, pad=(1, 1), residual=r, res_mode=2, head_w=hw, head_b=hb, out=out,
                                              out_img_stride=3 * H * H, passes=args.passes)
        ms = timeit(fn, args.iters, flush)
        rows.append((name, 2.0 * n * H * H * 256 * 576 / ms / 1e9, ms))

    cases = [
        ("pixel head 128x128 64->256->3", lambda: head_case("pixel head 128x128 64->256->3", 64, 128)),
        (f"qkv {M}x1536x512 f32", lambda: gemm_case(f"qkv {M}x1536x512 f32", M, 1536, 512, "f32")),
        (f"outproj {M}x512x512 res", lambda: gemm_case(f"outproj {M}x512x512 res", M, 512, 512, "res")),
        (f"fc {M}x2048x512 split", lambda: gemm_case(f"fc {M}x2048x512 split", M, 2048, 512, "split")),
        (f"proj {M}x512x2048 res", lambda: gemm_case(f"proj {M}x512x2048 res", M, 512, 2048, "res")),
        (f"head {M}x512x512 f32", lambda: gemm_case(f"head {M}x512x512 f32", M, 512, 512, "f32")),
        ("conv3x3 tok 64x16x16 512->512", lambda: conv_case("conv3x3 tok 64x16x16 512->512", 64, 16, 512, 512, 3)),
        ("dec 16x16 128->128", lambda: conv_case("dec 16x16 128->128", 64, 16, 128, 128, 3)),
        ("dec 16x16 128->512", lambda: conv_case("dec 16x16 128->512", 64, 16, 128, 512, 3, "f32")),
        ("dec 32x32 64->64", lambda: conv_case("dec 32x32 64->64", 64, 32, 64, 64, 3)),
        ("dec 64x64 64->64", lambda: conv_case("dec 64x64 64->64", 64, 64, 64, 64, 3)),
        ("dec 64x64 64->256", lambda: conv_case("dec 64x64 64->256", 64, 64, 64, 256, 3, "f32")),
        ("dec 128x128 64->64", lambda: conv_case("dec 128x128 64->64", 64, 128, 64, 64, 3)),
        ("dec 128x128 64->256", lambda: conv_case("dec 128x128 64->256", 64, 128, 64, 256, 3, "f32")),
    ]
    cases += [
        ("mainloop 16384x512x8192 f32", lambda: gemm_case("mainloop 16384x512x8192 f32", M, 512, 8192, "f32")),
        ("mainloop 16384x2048x4096 f32", lambda: gemm_case("mainloop 16384x2048x4096 f32", M, 2048, 4096, "f32")),
    ]
    if args.scan:
        cases = []
        for cin in (64, 128, 256):
            for cout in (64, 128, 256):
                nm = f"scan 128x128 {cin}->{cout}"
                cases.append((nm, (lambda nm=nm, cin=cin, cout=cout: conv_case(nm, 64, 128, cin, cout, 3))))
        for k in (64, 512, 1024, 2048):
            for n in (64, 128, 512):
                nm = f"scan gemm 16384x{n}x{k} f32"
                cases.append((nm, (lambda nm=nm, n=n, k=k: gemm_case(nm, M, n, k, "f32"))))
    ops.tc_nsplit(args.nsplit)
    table = {}
    cfgs = [tuple(int(v) for v in c.split(",")) for c in args.cfgs.split(";")]
    for bn, pair in cfgs:
        ops.tc_tuning(bn, pair)
        rows.clear()
        g.manual_seed(0)
        for name, fn in cases:
            if args.only and args.only not in name:
                continue
            fn()
        for name, tf, ms in rows:
            table.setdefault(name, []).append((tf, ms))
    ops.tc_tuning(0, -1)
    print(f"{'shape':36s} " + " ".join(f"{f'bn{b},pair{p_}':>18s}" for b, p_ in cfgs) + "   (TFLOP/s fp32-grade | ms)")
    for name, vals in table.items():
        print(f"{name:36s} " + " ".join(f"{tf:9.1f} |{ms:7.3f}" for tf, ms in vals))


if __name__ == "__main__":
    main()
