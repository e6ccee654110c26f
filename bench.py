#!/usr/bin/env python
"""ISTVT forward throughput on B200 (BASELINE.json metric: clips/sec forward; config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one forward of the hot path (Xception entry flow + 12 spatial-temporal blocks + head) over one
batch of 64 synthetic clips [64, 6, 3, 300, 300] per GPU, bf16 mode, random-init weights (seed 0).
Clips shard across GPUs with no collective on the data path (weak scaling: 64 clips per GPU);
torch.distributed (NCCL) is used for the barrier and the max-over-ranks of the device-timed duration only.

One JSON line on rank 0:
  value        whole-job clips/s with the input batch already resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through the public API (`ClipStream(model).run(batches)` -> `model(x)`) from PINNED
               HOST clips, H2D copy of every step's batch and D2H read of its logits inside the timed region;
               the copy of step i+1 overlaps the forward of step i (`serial_value`: no overlap, reference loop shape)
  roofline     the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of its launches / their summed
               CUDA-event durations, measured live in the timed region (ops.LaunchRecorder: events around every
               launch of every REC_STRIDE-th timed step — instrumenting all of them costs 0.5 ms per step), against
               the measured bf16 peak in MEASURED_PEAKS.json (or the profiling guide's fallback)
  cpu_baseline the reference forward on the host cores (fp32, all host threads) on a bounded sample of the same
               workload — rank 0, N=1 only.  kind "reference": the UNMODIFIED reference modules staged under
               oracle/_ref by oracle/make_ref.py (they travel with the gpurun snapshot); kind "port": the oracle
               restatement oracle/istvt_oracle.py when oracle/_ref is absent or the clip length is not 6
  gpu_eager_baseline  the same op sequence as eager PyTorch on the GPU (TF32 and bf16 autocast): SURVEY.md 8(d)'s
               "kernel to beat"; N=1 only, skipped with --no-eager-baseline
  kernels      per-kernel-family share of the step (launches, ms, TFLOP/s or GB/s) — explains `value`

`--impl reference` times the reference's own CPU implementation on the host cores (oracle/_ref when staged, else the
oracle port — /root/reference itself does not exist on the GPU box) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "2023-tifs-istvt_b200"

REC_STRIDE = 5      # per-launch CUDA events on every REC_STRIDE-th timed step (0.5 ms per instrumented step)
METRIC = "clips/sec forward"
UNIT = "clips/s"
GFLOP_PER_CLIP_T6 = 494.5            # 12 x 38.515 + 6 x 5.386
GFLOP_PER_CLIP_T32 = 2358.8          # 12 x 182.206 + 32 x 5.386
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            out = dict(FALLBACK_PEAKS)
            for k in out:
                if k in d and d[k]:
                    out[k] = float(d[k])
            return out, "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.lines:
            if ts < t0 or ts > t1:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w": round(sum(pw) / len(pw), 1),
                "samples": len(sm), "reasons": sorted(reasons)}


def use_all_host_threads() -> int:
    """The CPU legs use every host core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently turn the reference arm into a single-thread run; torch.set_num_threads overrides it."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


REF_DIR = os.path.join(ROOT, "oracle", "_ref")       # unmodified reference modules staged by oracle/make_ref.py


def cpu_forward_fn(frames: int):
    """(forward(x) on the host cores, kind, description).  kind "reference": the UNMODIFIED reference `XceptionVidTr`
    staged under oracle/_ref (oracle/make_ref.py; its constructor hard-codes 6 frames, vivit.py:201, so other clip
    lengths fall back); kind "port": the oracle restatement, pinned to the reference by tests/test_oracle.py."""
    import torch
    use_all_host_threads()
    if frames == 6 and os.path.isfile(os.path.join(REF_DIR, "network", "vivit", "vivit.py")):
        os.environ["ISTVT_REFERENCE_ROOT"] = REF_DIR
        from oracle import reference_shim as shim
        shim.REFERENCE_ROOT = REF_DIR
        product_network = {k: v for k, v in sys.modules.items() if k == "network" or k.startswith("network.")}
        try:
            model = shim.build_reference_model(seed=0)
        finally:      # the shim imports the reference's `network` package: put the product's aliases back
            for k in [k for k in sys.modules if k == "network" or k.startswith("network.")]:
                del sys.modules[k]
            sys.modules.update(product_network)
        return (lambda x: model(x)), "reference", "the unmodified reference XceptionVidTr (oracle/_ref via oracle/reference_shim.py)"
    from oracle import istvt_oracle as O
    pkg = importlib.import_module(PKG)
    torch.manual_seed(0)
    model = pkg.XceptionVidTr(num_frames=frames).eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    return (lambda x: O.forward(sd, x)), "port", "oracle/istvt_oracle.py"


def cpu_oracle_clips_per_s(batch: int, frames: int, budget_s: float, min_iters: int = 2):
    """Reference algorithm on the host cores, fp32, eval, no_grad (SURVEY.md §8d CPU baseline).
    Returns (clips/s, iterations, threads, kind, what)."""
    import torch
    fwd, kind, what = cpu_forward_fn(frames)
    x = torch.rand(batch, frames, 3, 300, 300, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        fwd(x)  # warm-up
        t0 = time.perf_counter()
        iters = 0
        while iters < min_iters or (time.perf_counter() - t0) < budget_s:
            fwd(x)
            iters += 1
            if time.perf_counter() - t0 > 3 * budget_s:
                break
        dt = time.perf_counter() - t0
    return batch * iters / dt, iters, torch.get_num_threads(), kind, what


def gpu_eager_clips_per_s(frames: int, dev, batch: int = 16, iters: int = 3):
    """The reference's algorithm as plain eager PyTorch ops ON THE SAME B200 ("the real kernel to
    beat"): the oracle port (same F.conv2d / F.linear / matmul / softmax / LayerNorm sequence, same permute copies and
    materialised score tensors as the reference) with its state_dict on the GPU, (a) fp32 with TF32 tensor cores allowed,
    (b) under torch.autocast(bfloat16).  cuDNN / cuBLAS / ATen do the work: library baselines, device-timed, inputs resident
    in HBM.  Reported next to cpu_baseline; never part of the product path."""
    import torch
    from oracle import istvt_oracle as O
    pkg = importlib.import_module(PKG)
    torch.manual_seed(0)
    model = pkg.XceptionVidTr(num_frames=frames).eval()
    sd = {k: v.detach().to(dev) for k, v in model.state_dict().items()}
    x = torch.rand(batch, frames, 3, 300, 300, generator=torch.Generator().manual_seed(1234)).to(dev)
    out = {}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        for name, ctx in (("fp32_tf32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            def step():
                with torch.no_grad():
                    if ctx is None:
                        return O.forward(sd, x)
                    with ctx:
                        return O.forward(sd, x)
            for _ in range(2):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            out[name] = batch * iters / (e0.elapsed_time(e1) / 1e3)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out, batch, iters


def gpu_eager_train_clips_per_s(frames: int, dev, batch: int = 16, iters: int = 2):
    """The reference's TRAINING iteration (train-mode forward with BatchNorm batch statistics, BCE-with-logits,
    autograd backward, AdamW on the 252 trained tensors; train_CNN.py:513-533) as eager PyTorch on the same B200: the
    oracle port (`loss_and_grads` + `adamw_update`) with its state_dict on the GPU, fp32 with TF32 allowed and under
    torch.autocast(bfloat16).  Device-timed, inputs resident.  A reported baseline only."""
    import torch
    from oracle import istvt_oracle as O
    pkg = importlib.import_module(PKG)
    out = {}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        for name in ("fp32_tf32", "bf16_autocast"):
            torch.manual_seed(0)
            model = pkg.XceptionVidTr(num_frames=frames)
            sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
            del model
            x = torch.rand(batch, frames, 3, 300, 300, generator=torch.Generator().manual_seed(1234)).to(dev)
            y = torch.randint(0, 2, (batch,), generator=torch.Generator().manual_seed(99)).to(dev)
            mom, var = {}, {}

            def step(i):
                if name == "bf16_autocast":
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        _, _, grads = O.loss_and_grads(sd, x, y)
                else:
                    _, _, grads = O.loss_and_grads(sd, x, y)
                with torch.no_grad():
                    for k, g in grads.items():
                        if k not in mom:
                            mom[k], var[k] = torch.zeros_like(sd[k]), torch.zeros_like(sd[k])
                        O.adamw_update(sd[k], g.float(), mom[k], var[k], i + 1, 1e-4)

            step(0)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                step(i + 1)
            e1.record()
            torch.cuda.synchronize(dev)
            out[name] = batch * iters / (e0.elapsed_time(e1) / 1e3)
            del sd, x, y, mom, var
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out, batch, iters


def run_reference(args) -> int:
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    fwd, kind, what = cpu_forward_fn(args.frames)
    sample = args.ref_clips
    x = torch.rand(sample, args.frames, 3, 300, 300, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd(x)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    threads = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, note=f"CPU arm: each step is a bounded sample of {sample} clip(s) of the workload"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{sample} clip(s) x {args.frames} frames x 300x300 per step, {args.steps} steps, "
                                   f"fp32, {what}, {threads} threads of {os.cpu_count()} logical CPUs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, note: str = "") -> dict:
    cfg = {
        "workload": f"{'C2' if args.frames == 6 else 'C5 (long clip)'}: ISTVT bf16 inference, {args.batch} clips x {args.frames} frames x 300x300 per GPU "
                    "(Xception entry flow + 12 spatial-temporal blocks + head), random-init weights seed 0",
        "batch_per_gpu": args.batch, "frames": args.frames, "image": 300, "precision": args.precision,
        "sharding": "clips sharded across GPUs, no data-path collective",
        "l2_policy": "no flush needed: the per-step input (%.0f MB) and every activation tensor exceed the 126 MB L2"
                     % (args.batch * args.frames * 3 * 300 * 300 * 4 / 1e6),
    }
    if note:
        cfg["note"] = note
    return cfg


def run_ours(args) -> int:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pkg = importlib.import_module(PKG)
    ops = pkg.ops
    torch.manual_seed(0)
    model = pkg.XceptionVidTr(num_frames=args.frames, precision=args.precision).eval().to(dev)

    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(args.batch, args.frames, 3, 300, 300, generator=gen).pin_memory()
    x_dev = x_host.to(dev)
    logits_host = torch.empty(args.batch, 1).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            model(x_dev)
        barrier()

        # ---------------- device-resident timed region ----------------
        sampler = ClockSampler(local) if rank == 0 else None
        time.sleep(0.25)
        # Per-launch CUDA events (the kernel breakdown and the roofline) are recorded INSIDE the timed region, but only on
        # every REC_STRIDE-th step: two events around each of ~170 launches cost 0.5 ms per step (1.1 %,
        # profiles/README.md r8j, tools/rec_overhead.py), which `value` should not carry on every step.
        rec = ops.LaunchRecorder()
        rec_steps = [i for i in range(args.steps) if i % REC_STRIDE == 0]
        n0 = pkg._lib.launch_count()
        barrier()
        t_wall0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            ops.set_recorder(rec if i % REC_STRIDE == 0 else None)
            logits = model(x_dev)
        e1.record()
        barrier()
        t_wall1 = time.time()
        launches = pkg._lib.launch_count() - n0
        ops.set_recorder(None)
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None

        # ---------------- end-to-end region: pinned host clips -> logits on the host ----------------
        # (a) serial: H2D, forward, D2H one after the other, as the reference's eval loop does (train_CNN.py:928-944)
        for _ in range(2):
            logits_host.copy_(model(x_host.to(dev, non_blocking=True)))
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            xd = x_host.to(dev, non_blocking=True)          # H2D of this step's clips
            logits_host.copy_(model(xd))                     # public API call + D2H of the step's result (syncs)
        f1.record()
        barrier()
        ms_e2e_serial = f0.elapsed_time(f1)
        del xd
        # (b) the package's feeder (ClipStream): the same per-step H2D + forward + D2H, with the copy of step i+1
        # issued on a side stream while step i computes.  Every step's copies are inside the timed region.
        feeder = pkg.ClipStream(model)
        sink = torch.empty(args.batch, 1)
        for out in feeder.run([x_host] * 2):
            sink.copy_(out)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for out in feeder.run([x_host] * args.steps):
            sink.copy_(out)
        g1.record()
        barrier()
        ms_e2e = g0.elapsed_time(g1)
        # (c) the same feeder on DECODED frames: uint8 [B, T, H, W, 3] pinned batches, normalisation folded into the
        # stem kernel — 4x fewer H2D bytes per step.  Reported beside `e2e`, which
        # stays on the reference-facing fp32 [B, T, 3, H, W] call.
        u8_host = (x_host * 255.0).round().to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous().pin_memory()
        for out in feeder.run([u8_host] * 2):
            sink.copy_(out)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for out in feeder.run([u8_host] * args.steps):
            sink.copy_(out)
        h1.record()
        barrier()
        ms_e2e_u8 = h0.elapsed_time(h1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e_serial, ms_e2e_u8], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_serial, ms_e2e_u8 = t.tolist()

    if rank == 0:
        total_clips = args.batch * world * args.steps
        value = total_clips / (ms / 1e3)
        e2e = total_clips / (ms_e2e / 1e3)
        peaks, peak_src = load_peaks()
        fam = rec.summary()
        kernels = {}
        n_rec = len(rec_steps)                      # steps whose launches carry events
        ms_rec = ms * n_rec / args.steps            # their share of the timed region
        for name, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            k = {"launches": d["launches"], "ms_per_step": d["ms"] / n_rec, "share": d["ms"] / ms_rec}
            if d["flops"] > 0:
                k["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12
            k["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            kernels[name] = k
        g = fam.get("gemm_bf16")
        roofline = None
        traffic, traffic_src = None, None
        try:   # DRAM bytes per GEMM launch from the committed ncu capture of this same command
            cands = sorted(n for n in os.listdir(os.path.join(ROOT, "profiles")) if n.endswith("_gemm_traffic.json"))
            with open(os.path.join(ROOT, "profiles", cands[-1])) as f:      # newest round's capture
                tj = json.load(f)
            traffic, traffic_src = tj["gemm_dram_bytes_per_launch"], tj["source"]
        except Exception:  # noqa: BLE001
            pass
        if g is not None and g["ms"] > 0:
            achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"]
            roofline = {"kernel": "gemm_tcgen05_kernel (istvt_gemm_fwd)", "bound": "tensor", "achieved": achieved,
                        "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                        "traffic_source": traffic_src,
                        "algorithmic_bytes_per_launch": g["bytes"] / g["launches"],
                        "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                        "launches_per_step": g["launches"] / n_rec,
                        "flops_per_step": g["flops"] / n_rec,
                        "share_of_step": g["ms"] / ms_rec,
                        "events": f"CUDA events around every launch of {n_rec} of the {args.steps} timed steps "
                                  f"(every {REC_STRIDE}th)"}
        flops_step = {6: GFLOP_PER_CLIP_T6, 32: GFLOP_PER_CLIP_T32}.get(args.frames)
        flops_step = flops_step * 1e9 * args.batch if flops_step else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "how": "ClipStream: pinned host batch -> H2D (side stream, overlapped with the previous step's "
                           "forward) -> model(x) -> D2H logits, every step",
                    "serial_value": total_clips / (ms_e2e_serial / 1e3),
                    "h2d_bytes_per_step": x_host.numel() * x_host.element_size() * world,
                    "d2h_bytes_per_step": logits_host.numel() * logits_host.element_size() * world},
            "e2e_uint8": {"value": total_clips / (ms_e2e_u8 / 1e3), "unit": UNIT, "ms_per_step": ms_e2e_u8 / args.steps,
                          "how": "ClipStream on decoded frames: pinned uint8 [B,T,H,W,3] -> H2D -> model(x_u8) (input "
                                 "normalisation folded into the stem kernel) -> D2H logits, every step",
                          "h2d_bytes_per_step": u8_host.numel() * world,
                          "d2h_bytes_per_step": logits_host.numel() * logits_host.element_size() * world},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
        }
        if flops_step is not None:
            # The reference's forward is 494.5 GFLOP per clip; the last block is pruned to the rows that can reach
            # the class token (engine.py), so the utilisation figures use the work ACTUALLY executed (sum of the
            # algorithmic flops of the launches in the timed region), never the skipped flops.
            executed = sum(d["flops"] for d in fam.values()) / args.steps
            line["whole_step"] = {"reference_algorithmic_tflop_per_step": flops_step / 1e12,
                                  "executed_tflop_per_step": executed / 1e12,
                                  "achieved_tflops_per_gpu": executed / (ms / args.steps * 1e-3) / 1e12,
                                  "frac_of_bf16_sustained": executed / (ms / args.steps * 1e-3) / 1e12
                                                            / peaks["bf16_tflops_sustained"]}
        if world == 1 and not args.no_cpu_baseline:
            v, iters, threads, kind, what = cpu_oracle_clips_per_s(1, args.frames, args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{iters} forwards of 1 clip x {args.frames} frames x 300x300 "
                                              f"({what}, fp32, {threads} torch threads of "
                                              f"{os.cpu_count()} logical CPUs)"}
        if world == 1 and not args.no_eager_baseline:
            try:
                vals, eb, ei = gpu_eager_clips_per_s(args.frames, dev, batch=min(args.batch, 64))
                line["gpu_eager_baseline"] = {
                    "fp32_tf32": vals["fp32_tf32"], "bf16_autocast": vals["bf16_autocast"], "unit": UNIT, "kind": "port",
                    "sample": f"{ei} forwards of {eb} clips x {args.frames} frames x 300x300, the oracle port of the "
                              "reference's op sequence run as eager PyTorch (cuDNN / cuBLAS / ATen) on this GPU, inputs "
                              "resident, CUDA events"}
            except Exception as e:  # noqa: BLE001 — a baseline must never take the bench line down
                line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_train(args) -> int:
    """--mode train: BASELINE.json config 3 — ISTVT bf16 training step (fwd + BCE + bwd + AdamW), data-parallel over
    the GPUs of one box with ONE NCCL all-reduce of the flat fp32 gradient buffer per step (pkg.Trainer)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    ops = pkg.ops
    torch.manual_seed(0)
    model = pkg.XceptionVidTr(num_frames=args.frames, precision="bf16").to(dev).train()
    trainer = pkg.Trainer(model, lr=5e-4, weight_decay=0.01)
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(args.batch, args.frames, 3, 300, 300, generator=gen).pin_memory()
    y_host = torch.randint(0, 2, (args.batch,), generator=gen).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        trainer.step(x_dev, y_dev)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25)
    rec = ops.LaunchRecorder()                  # events on every REC_STRIDE-th step only (see run_ours)
    rec_steps = [i for i in range(args.steps) if i % REC_STRIDE == 0]
    n0 = pkg._lib.launch_count()
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        ops.set_recorder(rec if i % REC_STRIDE == 0 else None)
        loss = trainer.step(x_dev, y_dev)
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = pkg._lib.launch_count() - n0
    ops.set_recorder(None)
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None
    peak_mem = torch.cuda.max_memory_allocated(dev) / 1e9

    # end to end: pinned host clips + labels -> H2D -> step -> loss value on the host, every step
    losses = []
    for _ in range(1):
        losses.append(float(trainer.step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True))))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        losses.append(float(trainer.step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True))))
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank == 0:
        total_clips = args.batch * world * args.steps
        peaks, peak_src = load_peaks()
        fam = rec.summary()
        kernels = {}
        n_rec = len(rec_steps)
        ms_rec = ms * n_rec / args.steps
        for name, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            k = {"launches": d["launches"], "ms_per_step": d["ms"] / n_rec, "share": d["ms"] / ms_rec}
            if d["flops"] > 0:
                k["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12
            k["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            kernels[name] = k
        gm = [fam[k] for k in ("gemm_bf16", "gemm_wgrad") if k in fam]
        g_ms = sum(d["ms"] for d in gm)
        g_fl = sum(d["flops"] for d in gm)
        peak = peaks["bf16_tflops_sustained"]
        roofline = {"kernel": "gemm_tcgen05 kernels (forward, data-gradient and split-K weight-gradient GEMMs)",
                    "bound": "tensor", "achieved": g_fl / (g_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "frac": g_fl / (g_ms * 1e-3) / 1e12 / peak, "traffic": None,
                    "peak_source": peak_src + ", sustained bf16", "share_of_step": g_ms / ms_rec,
                    "events": f"CUDA events around every launch of {n_rec} of the {args.steps} timed steps "
                              f"(every {REC_STRIDE}th)"}
        line = {
            "metric": "clips/sec training step (fwd+bwd+AdamW)", "value": total_clips / (ms / 1e3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C3: ISTVT bf16 training step, {args.batch} clips x {args.frames} frames x 300x300 per "
                                   "GPU, BCE-with-logits, AdamW, BatchNorm batch statistics, data-parallel with one NCCL "
                                   "all-reduce of the flat fp32 gradient buffer (357.9 MB) per step",
                       "batch_per_gpu": args.batch, "frames": args.frames, "peak_mem_gb": round(peak_mem, 1),
                       "l2_policy": "no flush needed: every activation tensor exceeds the 126 MB L2"},
            "e2e": {"value": total_clips / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": (x_host.numel() * 4 + y_host.numel() * 8) * world, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "final_loss": losses[-1],
            "whole_step": {"algorithmic_tflop_per_step": 3 * GFLOP_PER_CLIP_T6 * args.batch / 1e3,
                           "achieved_tflops_per_gpu": 3 * GFLOP_PER_CLIP_T6 * args.batch / 1e3 / (ms / args.steps * 1e-3)},
        }
        if world == 1 and not args.no_eager_baseline:
            try:
                vals = None
                for eb_try in (32, 16):      # the eager autograd graph of 64 fp32 clips does not fit 180 GB
                    try:
                        torch.cuda.empty_cache()
                        vals, eb, ei = gpu_eager_train_clips_per_s(args.frames, dev, batch=eb_try)
                        break
                    except torch.cuda.OutOfMemoryError:
                        continue
                if vals is None:
                    raise RuntimeError("out of memory at 16 clips")
                line["gpu_eager_baseline"] = {
                    "fp32_tf32": vals["fp32_tf32"], "bf16_autocast": vals["bf16_autocast"], "unit": UNIT, "kind": "port",
                    "sample": f"{ei} training iterations of {eb} clips x {args.frames} frames x 300x300, the oracle port of "
                              "the reference's train-mode forward + BCE + autograd backward + AdamW as eager PyTorch on "
                              "this GPU, inputs resident, CUDA events"}
            except Exception as e:  # noqa: BLE001 — a baseline must never take the bench line down
                line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_relevance(args) -> int:
    """--mode relevance: BASELINE.json config 4 — relevance pass (spatial + temporal maps), clips sharded over GPUs."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    ops = pkg.ops
    torch.manual_seed(0)
    model = pkg.XceptionVidTr(num_frames=args.frames, precision="bf16").to(dev).eval()
    x_host = torch.rand(args.batch, args.frames, 3, 300, 300, generator=torch.Generator().manual_seed(1234 + rank)).pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        pkg.relevance_maps(model, x_dev)
    barrier()
    rec = ops.LaunchRecorder()
    ops.set_recorder(rec)
    n0 = pkg._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cam_s, cam_t, _ = pkg.relevance_maps(model, x_dev)
    e1.record()
    barrier()
    launches = pkg._lib.launch_count() - n0
    ops.set_recorder(None)
    ms = e0.elapsed_time(e1)
    out_s = torch.empty(cam_s.shape).pin_memory()
    out_t = torch.empty(cam_t.shape).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        cs, ct, _ = pkg.relevance_maps(model, x_host.to(dev, non_blocking=True))
        out_s.copy_(cs); out_t.copy_(ct)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank == 0:
        total = args.batch * world * args.steps
        fam = rec.summary()
        kernels = {n: {"launches": d["launches"], "ms_per_step": d["ms"] / args.steps, "share": d["ms"] / ms}
                   for n, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        line = {"metric": "clips/sec relevance pass (spatial + temporal maps)", "value": total / (ms / 1e3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"C4: ISTVT relevance pass, {args.batch} clips x {args.frames} frames x 300x300 per GPU "
                                       "(forward keeping activations + activation-only backward + attention rollout); "
                                       "parity unpinned (reference implementation absent)",
                           "batch_per_gpu": args.batch},
                "e2e": {"value": total / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": x_host.numel() * 4 * world,
                        "d2h_bytes_per_step": (out_s.numel() + out_t.numel()) * 4 * world},
                "gpu_launches": int(launches), "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_ablation(args) -> int:
    """--mode ablation --variant vivit|vanilla: SURVEY.md section 8(f) rank 3 — the ablation transformers behind the same
    entry flow (`XceptionVidTr(variant=...)`), clips sharded over GPUs, no collective.  Not BASELINE.json's headline
    metric: a side line for the widened scope, same JSON shape."""
    import time

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    ops = pkg.ops
    torch.manual_seed(0)
    model = pkg.XceptionVidTr(num_frames=args.frames, precision=args.precision, variant=args.variant).eval()
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()} if rank == 0 and not args.no_cpu_baseline else None
    model = model.to(dev)
    x_host = torch.rand(args.batch, args.frames, 3, 300, 300, generator=torch.Generator().manual_seed(1234 + rank)).pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        model(x_dev)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25)
    rec = ops.LaunchRecorder()
    ops.set_recorder(rec)
    n0 = pkg._lib.launch_count()
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        logits = model(x_dev)
    e1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None
    launches = pkg._lib.launch_count() - n0
    ops.set_recorder(None)
    ms = e0.elapsed_time(e1)
    out = torch.empty(logits.shape).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        out.copy_(model(x_host.to(dev, non_blocking=True)))
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank == 0:
        total = args.batch * world * args.steps
        fam = rec.summary()
        peaks, peak_source = load_peaks()
        kernels = {}
        for n, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            k = {"launches": d["launches"], "ms_per_step": d["ms"] / args.steps, "share": d["ms"] / ms}
            if d["flops"]:
                k["tflops"] = d["flops"] / d["ms"] / 1e9
            if d["bytes"]:
                k["gbs"] = d["bytes"] / d["ms"] / 1e6
            kernels[n] = k
        top = "attn_joint" if "attn_joint" in fam else "gemm_bf16"
        d = fam.get(top)
        roof = None
        if d is not None and d["flops"]:
            ach = d["flops"] / d["ms"] / 1e9
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None,
                    "peak_source": peak_source + ", sustained bf16 (kernel timed inside a long step)",
                    "share_of_step": d["ms"] / ms}
        line = {"metric": f"clips/sec forward ({args.variant} ablation)", "value": total / (ms / 1e3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
                "data": "synthetic",
                "config": {"workload": f"8(f)-3 ablation: Xception entry flow + {type(model.vit).__name__} "
                                       f"({args.precision} inference), {args.batch} clips x {args.frames} frames x 300x300 "
                                       "per GPU, random-init weights seed 0",
                           "batch_per_gpu": args.batch, "variant": args.variant,
                           "l2_policy": "no flush needed: the per-step input and every activation tensor exceed the L2"},
                "e2e": {"value": total / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": x_host.numel() * 4 * world, "d2h_bytes_per_step": out.numel() * 4 * world},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels}
        if sd_cpu is not None:
            from oracle import ablation_oracle as A
            threads = use_all_host_threads()
            xs = x_host[:1].clone()
            with torch.no_grad():
                A.clip_forward(sd_cpu, xs, args.variant)
                n_it, t0 = 0, time.perf_counter()
                while n_it < 2 or time.perf_counter() - t0 < args.cpu_budget:
                    A.clip_forward(sd_cpu, xs, args.variant)
                    n_it += 1
                dt_s = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_it / dt_s, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_it} forwards of 1 clip (oracle/ablation_oracle.py clip_forward, fp32)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU")
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-clips", type=int, default=2, help="--impl reference: clips per CPU step")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip timing the reference's op sequence as eager PyTorch on the GPU (gpu_eager_baseline)")
    ap.add_argument("--variant", default="vanilla", choices=["vivit", "vanilla"], help="--mode ablation: which transformer")
    ap.add_argument("--mode", default="infer", choices=["infer", "train", "relevance", "ablation"],
                    help="infer: BASELINE.json's headline metric (C2, default; --frames 32 --batch 8 = C5); "
                         "train: the DP training step (C3); relevance: the relevance pass (C4, use --batch 32)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)
    if args.mode == "relevance":
        return run_relevance(args)
    if args.mode == "ablation":
        return run_ablation(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
