"""Timing of the fused SeparableConv2d kernel against the two-kernel path it replaces, at the C2 sizes (384 frames).

    python tools/sep_bench.py [--iters 10] [--frames 384]
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("2023-tifs-istvt_b200")
ops = pkg.ops

# (name, side, c_in, c_out, relu_in, act)
LAYERS = [("b1_sep1", 147, 64, 128, False, 1), ("b1_sep2", 147, 128, 128, False, 0),
          ("b2_sep1", 74, 128, 256, True, 1), ("b2_sep2", 74, 256, 256, False, 0)]


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--frames", type=int, default=384)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = "cuda"
    tot_f = tot_u = 0.0
    for name, side, c, n_out, relu_in, act in LAYERS:
        if args.only and name not in args.only.split(","):
            continue
        xs = [torch.randn(args.frames, side, side, c, device=dev).to(torch.bfloat16) for _ in range(2)]
        dw = torch.randn(3, 3, c, device=dev) * 0.3
        pw = (torch.randn(n_out, c, device=dev) * c ** -0.5).to(torch.bfloat16)
        bias = torch.randn(n_out, device=dev) * 0.1
        i = [0]

        def unfused():
            i[0] ^= 1
            return ops.gemm(ops.dwconv3x3(xs[i[0]], dw, relu_in=relu_in), pw, bias=bias, act=act)
        ms_u = timed(unfused, args.iters)
        by = (xs[0].numel() + args.frames * side * side * n_out) * 2
        line = f"{name} {side}x{side} {c:3d}->{n_out:3d}  two kernels {ms_u:6.3f} ms"
        tot_u += ms_u
        if ops.sepconv_fused_supported(c, n_out, side):
            def fused():
                i[0] ^= 1
                return ops.sepconv_fused(xs[i[0]], dw, pw, bias, relu_in, act)
            ms_f = timed(fused, args.iters)
            tot_f += ms_f
            line += f"   fused {ms_f:6.3f} ms = {by / ms_f / 1e6:5.0f} GB/s (in + out once)"
        else:
            tot_f += ms_u
            line += "   (not supported by the fused kernel)"
        print(line, flush=True)
        del xs
    print(f"sum: two kernels {tot_u:.3f} ms, fused where supported {tot_f:.3f} ms")


if __name__ == "__main__":
    main()
