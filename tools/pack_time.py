"""How long does weight packing take?  (SURVEY.md section 8(f) rank 4 proposes persisting the packed weights beside the
checkpoint.)  Times `load_state_dict` -> first forward (pack + forward) against a steady-state forward, batch 1.

    python tools/pack_time.py
"""
from __future__ import annotations

import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("2023-tifs-istvt_b200")


def main() -> None:
    torch.manual_seed(0)
    model = pkg.XceptionVidTr().cuda().eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.rand(1, 6, 3, 300, 300, device="cuda")
    model(x)
    torch.cuda.synchronize()
    steady = []
    for _ in range(5):
        t0 = time.perf_counter()
        model(x)
        torch.cuda.synchronize()
        steady.append(time.perf_counter() - t0)
    first = []
    for _ in range(3):
        model.load_state_dict(sd)                   # drops the packed-weight cache (a checkpoint load does the same)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model(x)
        torch.cuda.synchronize()
        first.append(time.perf_counter() - t0)
    s, f = min(steady) * 1e3, min(first) * 1e3
    print(f"steady-state forward (batch 1, wall clock): {s:.2f} ms; first forward after load_state_dict: {f:.2f} ms; "
          f"packing (BN folding, bf16 casts of 109 M parameters, on the GPU): {f - s:.2f} ms")


if __name__ == "__main__":
    main()
