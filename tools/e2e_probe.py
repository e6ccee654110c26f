"""Where does the end-to-end loop (ClipStream: pinned host batch -> H2D on a side stream -> forward -> D2H) lose time against
the device-resident loop?  Times (1) model(x_dev), (2) ClipStream, (3) ClipStream without the H2D copies, and the H2D copy
alone / under load.

    python tools/e2e_probe.py [--steps 10]
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("2023-tifs-istvt_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
torch.manual_seed(0)
model = pkg.XceptionVidTr(num_frames=6, precision="bf16").eval().cuda()
x_host = torch.rand(64, 6, 3, 300, 300).pin_memory()
x_dev = x_host.cuda()
sink = torch.empty(64, 1)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


def plain():
    for _ in range(args.steps):
        model(x_dev)


def stream(feeder):
    for out in feeder.run([x_host] * args.steps):
        sink.copy_(out)


with torch.no_grad():
    for _ in range(4):
        model(x_dev)
    feeder = pkg.ClipStream(model)
    stream(feeder)
    for rep in range(2):
        print(f"device-resident loop      {timed(plain):7.2f} ms per step", flush=True)
        print(f"ClipStream                {timed(lambda: stream(feeder)):7.2f} ms per step", flush=True)
        orig = feeder._prefetch

        def no_copy(slot, host_batch, used_before):
            with torch.cuda.stream(feeder._copy):
                feeder._ready[slot].record(feeder._copy)
        feeder._prefetch = no_copy
        print(f"ClipStream, no H2D copies {timed(lambda: stream(feeder)):7.2f} ms per step", flush=True)
        feeder._prefetch = orig
    # the copy alone, and while the forward runs
    buf = torch.empty_like(x_dev)
    side = torch.cuda.Stream()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        c0.record(side); buf.copy_(x_host, non_blocking=True); c1.record(side)
    torch.cuda.synchronize()
    print(f"H2D 415 MB alone          {c0.elapsed_time(c1):7.2f} ms", flush=True)
    model(x_dev)
    with torch.cuda.stream(side):
        c0.record(side); buf.copy_(x_host, non_blocking=True); c1.record(side)
    model(x_dev)
    torch.cuda.synchronize()
    print(f"H2D 415 MB under load     {c0.elapsed_time(c1):7.2f} ms", flush=True)
