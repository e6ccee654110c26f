"""Timing of conv2 (3x3, 32 -> 64 channels, 384 frames of 149 x 149) under each of its constructions.

    python tools/conv_bench.py [pair taps strip gather]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("2023-tifs-istvt_b200")
ops = pkg.ops

x = [torch.randn(384, 149, 149, 32, device="cuda").to(torch.bfloat16) for _ in range(2)]
w = (torch.randn(64, 3, 3, 32, device="cuda") * 0.05).to(torch.bfloat16)
b = torch.zeros(64, device="cuda")
by = x[0].numel() * 2 + 384 * 147 * 147 * 64 * 2
for kern in (sys.argv[1:] or ["pair", "taps"]):
    os.environ["ISTVT_CONV2_KERNEL"] = kern
    for i in range(3):
        ops.conv3x3(x[i % 2], w, b, kernel=kern)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        ops.conv3x3(x[i % 2], w, b, kernel=kern)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"conv2 {kern:7s} {ms:7.3f} ms  {by / ms / 1e6:6.0f} GB/s (in + out once)", flush=True)
