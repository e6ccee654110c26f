"""Timing of the LayerNorm backward at the training size (64 clips: 162 176 rows x 728, bf16 operands at the aligned pitch).

    python tools/ln_bench.py [--iters 20]
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
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = "cuda"
    b, f, p, d = 64, 7, 362, 728
    rows = b * f * p
    dy = ops.pad_rows((torch.randn(rows, d, device=dev) * 0.1).to(torch.bfloat16))
    dy2 = ops.pad_rows((torch.randn(rows, d, device=dev) * 0.1).to(torch.bfloat16))
    x32 = torch.randn(rows, d, device=dev)
    xbf = ops.pad_rows(torch.randn(rows, d, device=dev).to(torch.bfloat16))
    gamma = torch.randn(d, device=dev)
    dg, db, cs = (torch.zeros(d, device=dev) for _ in range(3))
    g = torch.zeros(rows, d, device=dev)
    gbf = ops.empty_rows((rows, d), torch.bfloat16, dev)
    cases = {
        "LN3/LN1-style: x fp32, g += dx, bf16 copy, colsum": lambda: ops.layernorm_bwd(dy, x32, gamma, dg, db, g_accum=g, g_bf16=gbf, out_colsum=cs),
        "LN1 with the frame-difference backward (dy2)": lambda: ops.layernorm_bwd(dy, x32, gamma, dg, db, g_accum=g, g_bf16=gbf, dy2=dy2, frames=f, tokens_per_frame=p, out_colsum=cs),
        "LN2-style: x bf16, dx out, colsum": lambda: ops.layernorm_bwd(dy, xbf, gamma, dg, db, out_colsum=cs),
    }
    for name, fn in cases.items():
        ms = timed(fn, args.iters)
        if "LN2" in name:
            by = rows * d * (2 + 2 + 2)
        else:
            by = rows * d * (2 + 4 + 8 + 2) + (rows * d * 2 * 2 if "dy2" in name else 0)
        print(f"{name:60s} {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
