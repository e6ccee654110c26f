"""Per-layer timing of the HBM-bound entry-flow kernels at the C2 sizes (384 frames): depthwise 3x3, maxpool+add,
stem.  Prints ms and algorithmic GB/s (in + out bytes) per call site.

    python tools/entry_bench.py [--frames 384] [--iters 10]
    ISTVT_DW_TILES=1 python tools/entry_bench.py      # the older 16x16-tile depthwise kernel, for A/B
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

DW = [(147, 64, False), (147, 128, True), (74, 128, True), (74, 256, True), (37, 256, True), (37, 728, True)]
POOL = [(147, 128), (74, 256), (37, 728)]


def timed(fn, iters):
    for _ in range(3):
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
    ap.add_argument("--frames", type=int, default=384)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    n = args.frames
    dev = "cuda"
    tot_ms = tot_b = 0.0
    for (hw, c, relu) in DW:
        xs = [torch.randn(n, hw, hw, c, device=dev).to(torch.bfloat16) for _ in range(2)]   # > L2 for every layer
        w = torch.randn(3, 3, c, device=dev)
        i = [0]

        def f():
            i[0] ^= 1
            ops.dwconv3x3(xs[i[0]], w, relu)
        ms = timed(f, args.iters)
        by = 2 * xs[0].numel() * 2
        tot_ms += ms
        tot_b += by
        print(f"dwconv {hw:3d}x{hw:<3d} c={c:3d}  {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s", flush=True)
        del xs
    print(f"dwconv sum {tot_ms:7.3f} ms  {tot_b / tot_ms / 1e6:7.0f} GB/s")
    tot_ms = tot_b = 0.0
    for (hw, c) in POOL:
        x = torch.randn(n, hw, hw, c, device=dev).to(torch.bfloat16)
        ho = (hw - 1) // 2 + 1
        skip = torch.randn(n, ho, ho, c, device=dev).to(torch.bfloat16)
        ms = timed(lambda: ops.pool_add(x, skip), args.iters)
        by = (x.numel() + 2 * skip.numel()) * 2
        tot_ms += ms
        tot_b += by
        print(f"pool_add {hw:3d}x{hw:<3d} c={c:3d}  {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s", flush=True)
        del x, skip
    print(f"pool_add sum {tot_ms:7.3f} ms  {tot_b / tot_ms / 1e6:7.0f} GB/s")
    x = torch.rand(n, 3, 300, 300, device=dev)
    wt = torch.randn(32, 3, 3, 3, device=dev) * 0.2
    b = torch.zeros(32, device=dev)
    ms = timed(lambda: ops.conv_stem(x, wt, b, torch.bfloat16), args.iters)
    by = x.numel() * 4 + n * 149 * 149 * 32 * 2
    print(f"conv_stem            {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
