"""How much do the per-launch CUDA events of bench.py's LaunchRecorder cost?  C2 step with and without the recorder.

    python tools/rec_overhead.py [--steps 15]
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("2023-tifs-istvt_b200")
ops = pkg.ops

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=15)
args = ap.parse_args()
torch.manual_seed(0)
model = pkg.XceptionVidTr(num_frames=6, precision="bf16").eval().cuda()
x = torch.rand(64, 6, 3, 300, 300, device="cuda")
with torch.no_grad():
    for _ in range(4):
        model(x)
    for rep in range(3):
        for rec_on in (False, True):
            ops.set_recorder(ops.LaunchRecorder() if rec_on else None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                model(x)
            e1.record()
            torch.cuda.synchronize()
            ops.set_recorder(None)
            print(f"recorder={'on ' if rec_on else 'off'}  {e0.elapsed_time(e1) / args.steps:7.3f} ms per step", flush=True)
