"""Run every kernel / model parity check in its own subprocess with a timeout (one faulting or hanging
kernel must not take the rest of a remote GPU session down) and write a JSON report.

    python tools/gpu_check.py [--out gpurun_out/checks.json] [--only name1,name2] [--timeout 300]
    python tools/gpu_check.py --one <name>        (internal: run one check in this process)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def all_checks():
    import kernel_checks
    import model_checks
    checks = dict(kernel_checks.CHECKS)
    checks["api"] = model_checks.run_api_checks
    for prec in ("fp32", "bf16"):
        checks[f"golden_default_{prec}"] = lambda p=prec: model_checks.run_golden_case("default_init_b1", p)
        checks[f"golden_sens_{prec}"] = lambda p=prec: model_checks.run_golden_case("sensitised_b2", p)
        checks[f"oracle_{prec}"] = lambda p=prec: model_checks.run_oracle_case(p)
    checks["golden_t32_bf16"] = lambda: model_checks.run_golden_case("sensitised_t32_b1", "bf16")
    checks["golden_t32_fp32"] = lambda: model_checks.run_golden_case("sensitised_t32_b1", "fp32")
    checks["batch64"] = model_checks.run_batch64_parity
    import torch
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        checks["data_parallel"] = model_checks.run_data_parallel_check
    checks["train_golden"] = model_checks.run_train_golden
    checks["train_t32_oracle"] = model_checks.run_train_t32_oracle
    checks["relevance"] = model_checks.run_relevance_check
    checks["relevance_t32"] = model_checks.run_relevance_t32_check
    checks["cuda_graph"] = model_checks.run_graph_check
    checks["pack_cache"] = model_checks.run_pack_cache_check
    checks["uint8_input"] = model_checks.run_uint8_input_check
    checks["xception_fp32"] = lambda: model_checks.run_xception_golden("fp32")
    checks["xception_bf16"] = lambda: model_checks.run_xception_golden("bf16")
    for prec in ("fp32", "bf16"):
        for case in ("vivit_d2_b2", "vivit_mean_d2_b2", "vivit_d12_b1", "vanilla_d2_b1", "vanilla_d12_b1"):
            checks[f"ablation_{case}_{prec}"] = lambda c=case, p=prec: model_checks.run_ablation_golden(c, p)
        checks[f"ablation_blocks_{prec}"] = lambda p=prec: model_checks.run_ablation_blocks(p)
    for variant in ("vivit", "vanilla"):
        checks[f"ablation_clip_{variant}"] = lambda v=variant: model_checks.run_ablation_clip_check(v)
    return checks


def run_one(name: str) -> int:
    import torch
    checks = all_checks()
    t0 = time.time()
    try:
        res = checks[name]()
        torch.cuda.synchronize()
        print("RESULT " + json.dumps({"name": name, "ok": True, "sec": round(time.time() - t0, 2),
                                      "metrics": {k: float(v) for k, v in (res or {}).items()}}))
        return 0
    except Exception as e:  # noqa: BLE001
        tb = traceback.format_exc()
        print("RESULT " + json.dumps({"name": name, "ok": False, "sec": round(time.time() - t0, 2),
                                      "error": f"{type(e).__name__}: {e}"[:6000], "trace": tb[-3000:]}))
        return 1


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--one")
    ap.add_argument("--only", default="")
    ap.add_argument("--skip", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "checks.json"))
    ap.add_argument("--timeout", type=int, default=420)
    args = ap.parse_args()
    if args.one:
        return run_one(args.one)
    names = list(all_checks().keys())
    if args.only:
        names = [n for n in names if n in args.only.split(",")]
    if args.skip:
        names = [n for n in names if n not in args.skip.split(",")]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    report = []
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True,
                               timeout=args.timeout)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                rec = json.loads(line[-1][7:])
            else:
                rec = {"name": n, "ok": False, "error": f"no result, rc={r.returncode}",
                       "stdout": r.stdout[-3000:], "stderr": r.stderr[-3000:]}
            if not rec.get("ok"):
                rec.setdefault("stderr", r.stderr[-2000:])
                rec.setdefault("stdout", r.stdout[-2000:])
        except subprocess.TimeoutExpired as e:
            rec = {"name": n, "ok": False, "error": f"timeout after {args.timeout}s",
                   "stdout": (e.stdout or b"")[-2000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-2000:]}
        rec["wall"] = round(time.time() - t0, 1)
        report.append(rec)
        status = "ok  " if rec.get("ok") else "FAIL"
        worst = ""
        if rec.get("ok") and rec.get("metrics"):
            k = max(rec["metrics"], key=lambda q: rec["metrics"][q])
            worst = f" worst {k}={rec['metrics'][k]:.2e}"
        print(f"[{status}] {n} ({rec['wall']}s){worst}" + ("" if rec.get("ok") else "\n    " + str(rec.get("error"))[:1500]), flush=True)
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)
    bad = [r["name"] for r in report if not r.get("ok")]
    print(f"{len(report) - len(bad)}/{len(report)} checks passed" + (f"; failed: {bad}" if bad else ""))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
