"""Timing of the spatial / temporal attention kernels at the C2 sizes (64 clips: 448 frames x 362 tokens x 8 heads).

    python tools/attn_bench.py [--clips 64] [--iters 20]
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


def joint(args) -> None:
    heads, dev = 8, "cuda"
    for clips, n in ((args.clips, args.frames * 361 + 1), (args.clips * (args.frames + 1), 362)):
        qkvs = [torch.randn(clips * n, 1536, device=dev).to(torch.bfloat16) for _ in range(2)]
        i = [0]

        def run():
            i[0] ^= 1
            ops.attn_joint(qkvs[i[0]], clips, n, heads, 0.125)
        ms = timed(run, args.iters)
        fl = 4.0 * clips * heads * n * n * 64
        print(f"attn_joint   {clips} sequences x {n} tokens  {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  "
              f"{(qkvs[0].numel() * 2 + clips * n * 512 * 2) / ms / 1e6:7.0f} GB/s (qkv + out once)", flush=True)
        if n <= 384:
            ms = timed(lambda: ops.attn_spatial(qkvs[0], clips, n, heads, 0.125), args.iters)
            print(f"attn_spatial {clips} sequences x {n} tokens  {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--bwd", action="store_true", help="also time the spatial-attention backward (training)")
    ap.add_argument("--joint", action="store_true",
                    help="only the key-streaming joint attention at VanillaTr's sequence length (frames*361+1 tokens)")
    args = ap.parse_args()
    if args.joint:
        return joint(args)
    b, f, p, heads = args.clips, args.frames + 1, 362, 8
    dev = "cuda"
    qkvs = [torch.randn(b * f * p, 1536, device=dev).to(torch.bfloat16) for _ in range(2)]
    i = [0]

    def spatial():
        i[0] ^= 1
        ops.attn_spatial(qkvs[i[0]], b * f, p, heads, 0.125)
    ms = timed(spatial, args.iters)
    fl = 4.0 * b * f * heads * p * p * 64
    print(f"attn_spatial  {b * f} frames x {p} tokens  {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  "
          f"{(qkvs[0].numel() * 2 + b * f * p * 512 * 2) / ms / 1e6:7.0f} GB/s", flush=True)
    if args.bwd:
        o, lse = ops.attn_spatial_lse(qkvs[0], b * f, p, heads, 0.125)
        dout = torch.randn_like(o)
        scratch = torch.empty(b * f * p, 512, dtype=torch.float32, device=dev)
        ms = timed(lambda: ops.attn_spatial_bwd(qkvs[0], o, dout, lse, b * f, p, heads, 0.125, scratch=scratch), args.iters)
        print(f"attn_spatial_bwd {b * f} frames x {p} tokens  {ms:7.3f} ms  {2.5 * fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    qk = torch.randn(b * f * p, 1024, device=dev).to(torch.bfloat16)
    v = torch.randn(b * f * p, 512, device=dev).to(torch.bfloat16)
    ms = timed(lambda: ops.attn_temporal(qk, v, b, f, p, heads, 0.125), args.iters)
    by = (qk.numel() + 2 * v.numel()) * 2
    print(f"attn_temporal {b} clips x {f} frames x {p} positions  {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
