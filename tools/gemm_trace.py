"""Debug: timeline of the first tiles of cluster 0 of the CTA-pair GEMM (globaltimer stamps recorded by the kernel).

Needs a debug build:   ISTVT_BUILD_DEFS=-DISTVT_GEMM_TRACE python 2023-tifs-istvt_b200/build.py --force
then (GPU box):        python tools/gemm_trace.py [--n 1536 --k 728]
Rebuild without the define afterwards (python 2023-tifs-istvt_b200/build.py --force): the product never ships the trace.
"""
import argparse
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("2023-tifs-istvt_b200")
ops = pkg.ops

NAMES = ["tma_first", "tma_last", "acc_free", "kb0_landed", "mma_done_issued", "epi_full_seen", "epi_tmem_released",
         "epi_stores_issued"]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=162176)
    ap.add_argument("--n", type=int, default=1536)
    ap.add_argument("--k", type=int, default=728)
    args = ap.parse_args()
    lib = pkg._lib.lib()
    if not hasattr(lib, "istvt_debug_gemm_trace"):
        raise SystemExit("not a trace build: ISTVT_BUILD_DEFS=-DISTVT_GEMM_TRACE python 2023-tifs-istvt_b200/build.py --force")
    dev = "cuda"
    a = torch.randn(args.m, args.k, device=dev).to(torch.bfloat16)
    w = (torch.randn(args.n, args.k, device=dev) * args.k ** -0.5).to(torch.bfloat16)
    out = torch.empty(args.m, args.n, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, out=out)
    buf = torch.zeros(24 * 8, dtype=torch.int64, device=dev)
    lib.istvt_debug_gemm_trace.argtypes = [ctypes.c_void_p]
    assert lib.istvt_debug_gemm_trace(buf.data_ptr()) == 0
    torch.cuda.synchronize()
    ops.gemm(a, w, out=out)
    torch.cuda.synchronize()
    assert lib.istvt_debug_gemm_trace(None) == 0
    t = buf.view(24, 8).cpu()
    t0 = int(t[0, 0])
    print("tile " + " ".join(f"{n:>18s}" for n in NAMES) + "   (us since the first TMA of tile 0)")
    for i in range(24):
        print(f"{i:4d} " + " ".join(f"{(int(v) - t0) / 1e3:18.2f}" if int(v) else f"{'-':>18s}" for v in t[i]))
    d = (t[1:, 4] - t[:-1, 4]).float() / 1e3
    print("tile period (mma_done to mma_done), us:", [round(float(x), 2) for x in d[:20]])


if __name__ == "__main__":
    main()
