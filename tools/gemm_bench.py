"""Per-shape timing of the tcgen05 GEMMs of the C2 workload (M = 64 clips x 2534 tokens = 162176 rows).

    python tools/gemm_bench.py [--iters 20] [--m 162176] [--only ff1,ff2]

Prints one line per (call site): ms, TFLOP/s, effective GB/s.  Inputs are rotated over several buffers so that
consecutive launches do not hit in L2 more than the real schedule does.
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

# name: (N, K, bias, residual(fp32 in-place), act, out dtype)
SHAPES = {
    "to_qk": (1024, 728, False, False, 0, torch.bfloat16),
    "to_v": (512, 728, False, False, 0, torch.bfloat16),
    "t_out": (728, 512, True, False, 0, torch.bfloat16),
    "to_qkv": (1536, 728, False, False, 0, torch.bfloat16),
    "s_out": (728, 512, True, True, 0, torch.float32),
    "ff1": (2912, 728, True, False, 2, torch.bfloat16),
    "ff2": (728, 2912, True, True, 0, torch.float32),
}


# name: (pixels per frame, N, K) — 1x1 convolutions + folded BN (+ ReLU) as GEMMs over NHWC pixels
ENTRY = {
    "b1_pw1": (147 * 147, 128, 64), "b1_pw2": (147 * 147, 128, 128), "b1_skip": (74 * 74, 128, 64),
    "b2_pw1": (74 * 74, 256, 128), "b2_pw2": (74 * 74, 256, 256), "b2_skip": (37 * 37, 256, 128),
    "b3_pw1": (37 * 37, 728, 256), "b3_pw2": (37 * 37, 728, 728), "b3_skip": (19 * 19, 728, 256),
}


def entry_flow(args) -> None:
    dev = "cuda"
    frames = 384
    tot_ms = tot_by = 0.0
    for name, (px, n, k) in ENTRY.items():
        m = frames * px
        a = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(2)]
        w = (torch.randn(n, k, device=dev) * k ** -0.5).to(torch.bfloat16)
        b = torch.randn(n, device=dev)
        outs = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(2)]
        for i in range(3):
            ops.gemm(a[i % 2], w, bias=b, act=1, out=outs[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            ops.gemm(a[i % 2], w, bias=b, act=1, out=outs[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        by = (m * k + m * n + n * k) * 2
        tot_ms += ms
        tot_by += by
        print(f"{name:8s} M={m:8d} N={n:4d} K={k:4d} {ms:7.3f} ms  {2.0 * m * n * k / ms / 1e9:7.1f} TFLOP/s  "
              f"{by / ms / 1e6:6.0f} GB/s", flush=True)
        del a, outs
    print(f"entry-flow GEMM sum {tot_ms:7.3f} ms  {tot_by / tot_ms / 1e6:6.0f} GB/s")


def pitch_experiment(args) -> None:
    """Does the row pitch of the K-major operands matter?  K = 728 rows are 1456 B apart, so 7 of every 8 box rows
    (128 B of one operand row) straddle two 128-byte lines; with a pitch of 768 elements every box row is one line."""
    from importlib import import_module
    _lib = import_module("2023-tifs-istvt_b200._lib")
    lib = _lib.lib()
    dev = "cuda"
    m = args.m
    st = torch.cuda.current_stream().cuda_stream
    for (n, k) in ((1536, 728), (2912, 728), (1024, 728), (728, 2912)):
        for (lda, ldw, ldc) in ((k, k, n), (k + 40 if k == 728 else k + 32, k, n), (k, k + 40 if k == 728 else k + 32, n),
                                (k + 40 if k == 728 else k + 32, k + 40 if k == 728 else k + 32, n),
                                (k, k, (n + 63) // 64 * 64)):
            nbuf = 3
            a = [torch.randn(m, lda, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
            w = torch.randn(n, ldw, device=dev, dtype=torch.bfloat16) * k ** -0.5
            outs = [torch.zeros(m, ldc, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]

            def run(i):
                _lib.check(lib.istvt_gemm_fwd(a[i % nbuf].data_ptr(), lda, w.data_ptr(), ldw, outs[i % nbuf].data_ptr(), ldc,
                                              ops.BF16, m, n, k, None, None, n, 0, st),
                           "istvt_gemm_fwd")
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            ref = (a[(args.iters - 1) % nbuf][:4096, :k].float() @ w[:, :k].float().t())
            err = ((outs[(args.iters - 1) % nbuf][:4096, :n].float() - ref).abs().max() / ref.abs().max()).item()
            print(f"N={n:5d} K={k:5d} lda={lda:5d} ldw={ldw:5d} ldc={ldc:5d} {ms:7.3f} ms  {2.0 * m * n * k / ms / 1e9:7.1f} TFLOP/s"
                  f"  rel_err {err:.1e}", flush=True)
            del a, outs


def residual_pitch_experiment(args) -> None:
    """In-place fp32 residual update (s_out / ff2: TMA reduce-add of 128-byte box rows) with the residual stream at
    pitch 728 (2912 B: 3 box rows of 4 straddle two lines) vs 768 (3072 B)."""
    from importlib import import_module
    _lib = import_module("2023-tifs-istvt_b200._lib")
    lib = _lib.lib()
    dev = "cuda"
    m = args.m
    st = torch.cuda.current_stream().cuda_stream
    for (n, k) in ((728, 512), (728, 2912)):
        for ldc in (728, 768):
            nbuf = 3
            a = [torch.randn(m, k, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
            w = torch.randn(n, k, device=dev, dtype=torch.bfloat16) * k ** -0.5
            bias = torch.randn(n, device=dev)
            outs = [torch.zeros(m, ldc, device=dev, dtype=torch.float32) for _ in range(nbuf)]

            def run(i):
                o = outs[i % nbuf]
                _lib.check(lib.istvt_gemm_fwd(a[i % nbuf].data_ptr(), k, w.data_ptr(), k, o.data_ptr(), ldc, ops.F32, m, n, k,
                                              bias.data_ptr(), o.data_ptr(), ldc, 0, st), "istvt_gemm_fwd")
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            by = m * k * 2 + n * k * 2 + 2 * m * n * 4
            print(f"residual in place N={n:5d} K={k:5d} ldc={ldc:5d} {ms:7.3f} ms  {2.0 * m * n * k / ms / 1e9:7.1f} TFLOP/s"
                  f"  {by / ms / 1e6:6.0f} GB/s", flush=True)
            del a, outs


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--m", type=int, default=162176)
    ap.add_argument("--only", default="")
    ap.add_argument("--custom", default="", help="extra shapes 'N,K;N,K' (bf16 out, no bias)")
    ap.add_argument("--cublas", action="store_true", help="also time torch.matmul (cuBLAS) on the same operands: a "
                                                          "yardstick for what the shape can reach, not a product path")
    ap.add_argument("--entry", action="store_true", help="the 9 pointwise / skip GEMMs of the Xception entry flow at "
                                                         "the C2 size (384 frames) instead of the transformer's")
    ap.add_argument("--aligned", action="store_true", help="operands / bf16 outputs at the engine's aligned row pitch")
    ap.add_argument("--residual-pitch", action="store_true", help="s_out / ff2 in place with the fp32 stream at pitch 728 vs 768")
    ap.add_argument("--pitch", action="store_true", help="row-pitch experiment: K = 728 operands with pitch 728 vs 768")
    args = ap.parse_args()
    if args.entry:
        return entry_flow(args)
    if args.pitch:
        return pitch_experiment(args)
    if args.residual_pitch:
        return residual_pitch_experiment(args)
    for i, nk in enumerate(filter(None, args.custom.split(";"))):
        n_, k_ = (int(v) for v in nk.split(","))
        SHAPES[f"c{n_}x{k_}"] = (n_, k_, False, False, 0, torch.bfloat16)
    m = args.m
    dev = "cuda"
    tot_ms, tot_fl = 0.0, 0.0
    for name, (n, k, bias, res, act, odt) in SHAPES.items():
        if args.only and name not in args.only.split(","):
            continue
        nbuf = 3
        a = [torch.randn(m, k, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        w = torch.randn(n, k, device=dev, dtype=torch.bfloat16) * k ** -0.5
        if args.aligned:      # the engine's layout: K = 728 operands at the 128-byte aligned row pitch (768 elements)
            a = [ops.pad_rows(t) for t in a]
            w = ops.pad_rows(w)
        b = torch.randn(n, device=dev) if bias else None
        outs = [torch.zeros(m, n, device=dev, dtype=odt) for _ in range(nbuf)]
        if args.aligned and odt == torch.bfloat16:
            outs = [ops.pad_rows(t) for t in outs]
        for i in range(3):
            ops.gemm(a[i % nbuf], w, bias=b, residual=outs[i % nbuf] if res else None, act=act, out=outs[i % nbuf])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            ops.gemm(a[i % nbuf], w, bias=b, residual=outs[i % nbuf] if res else None, act=act, out=outs[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        fl = 2.0 * m * n * k
        by = m * k * 2 + n * k * 2 + m * n * outs[0].element_size() * (2 if res else 1)
        tot_ms += ms
        tot_fl += fl
        extra = ""
        if args.cublas:
            if args.aligned:      # cuBLAS on the dense operands (torch.matmul would copy a strided view anyway)
                a = [t.contiguous() for t in a]
                w = w.contiguous()
            wt = w.t()
            for i in range(3):
                torch.matmul(a[i % nbuf], wt)
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.iters):
                torch.matmul(a[i % nbuf], wt)
            e1.record()
            torch.cuda.synchronize()
            msc = e0.elapsed_time(e1) / args.iters
            extra = f"   | cuBLAS (plain A.Wt, bf16 out) {msc:7.3f} ms {fl / msc / 1e9:7.1f} TFLOP/s"
        print(f"{name:10s} N={n:5d} K={k:5d} {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  {by / ms / 1e6:7.0f} GB/s{extra}", flush=True)
        del a, outs
    if tot_ms:
        print(f"layer sum {tot_ms:7.3f} ms  {tot_fl / tot_ms / 1e9:7.1f} TFLOP/s  (x12 layers = {12 * tot_ms:.1f} ms)")


if __name__ == "__main__":
    main()
