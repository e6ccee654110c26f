"""Run the kernel parity checks (tests/kernel_checks.py) in ONE process, meant to be launched under compute-sanitizer:

    compute-sanitizer --tool memcheck  --kernel-regex kns=istvt --log-file gpurun_out/memcheck.log \
        python tools/sanitizer_run.py [--only a,b] [--skip c,d]
    compute-sanitizer --tool racecheck --kernel-regex kns=istvt --log-file gpurun_out/racecheck.log \
        python tools/sanitizer_run.py --only gemm_basic,layernorm,...

Only this library's kernels (namespace istvt) are instrumented; torch's own kernels run uninstrumented.  Every check runs
in try/except: a parity failure or a sanitizer-reported error does not stop the remaining checks.  Prints one line per
check and a JSON summary (also written to --out).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--skip", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--budget", type=float, default=1e9, help="stop starting new checks after this many seconds")
    args = ap.parse_args()
    import torch
    import kernel_checks
    names = list(kernel_checks.CHECKS.keys())
    if args.only:
        names = [n for n in names if n in args.only.split(",")]
    if args.skip:
        names = [n for n in names if n not in args.skip.split(",")]
    t_start = time.time()
    report = []
    for n in names:
        if time.time() - t_start > args.budget:
            report.append({"name": n, "ok": None, "error": "not started: time budget"})
            print(f"[skip] {n}: time budget", flush=True)
            continue
        t0 = time.time()
        try:
            kernel_checks.CHECKS[n]()
            torch.cuda.synchronize()
            rec = {"name": n, "ok": True}
        except Exception as e:  # noqa: BLE001
            rec = {"name": n, "ok": False, "error": f"{type(e).__name__}: {e}"[:2000], "trace": traceback.format_exc()[-1500:]}
        rec["sec"] = round(time.time() - t0, 1)
        report.append(rec)
        print(f"[{'ok  ' if rec['ok'] else 'FAIL'}] {n} ({rec['sec']}s)" + ("" if rec["ok"] else " " + rec["error"][:300]), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)
    bad = [r["name"] for r in report if r["ok"] is False]
    print(f"{sum(1 for r in report if r['ok'])}/{len(report)} checks ran clean" + (f"; failed: {bad}" if bad else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
