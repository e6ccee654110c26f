"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel family the launches, serialised device time and its share, and the DRAM traffic; writes the GEMM-family
traffic per launch to a JSON that bench.py quotes as `roofline.traffic`.

    python tools/launches_summary.py profiles/r3_launches.csv [--json profiles/r3_gemm_traffic.json]
"""
import argparse
import csv
import json
import re
from collections import defaultdict

FAMILIES = [("gemm", r"gemm_tcgen05"), ("attn_joint", r"attn_joint"), ("attn_spatial", r"attn_spatial"), ("attn_temporal", r"attn_temporal"),
            ("sepconv_fused", r"sepconv_fused"), ("dwconv3x3", r"dwconv3x3"), ("layernorm_diff", r"layernorm_diff"), ("layernorm", r"layernorm"),
            ("pool_add", r"pool_add"), ("conv_stem", r"conv_stem"), ("subsample2", r"subsample2"),
            ("token/head/gather", r"token_fill|token_build|mean_rows|head_kernel|gather_rows")]


def family(name: str) -> str:
    for fam, pat in FAMILIES:
        if re.search(pat, name):
            return fam
    return "other:" + name.split("(")[0][-40:]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    lines = [l for l in open(args.csv) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = defaultdict(lambda: defaultdict(float))     # launch id -> metric -> value
    names = {}
    for r in rows:
        per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        per[r["ID"]]["unit:" + r["Metric Name"]] = r["Metric Unit"]
        names[r["ID"]] = r["Kernel Name"]
    fam = defaultdict(lambda: {"launches": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    for i, m in per.items():
        f = fam[family(names[i])]
        f["launches"] += 1
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(m["unit:gpu__time_duration.sum"], 1.0)
        f["ns"] += m["gpu__time_duration.sum"] * scale
        f["rd"] += m.get("dram__bytes_read.sum", 0.0)
        f["wr"] += m.get("dram__bytes_write.sum", 0.0)
    tot = sum(f["ns"] for f in fam.values())
    print(f"# {len(per)} launches, {tot / 1e6:.2f} ms serialised")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"{k:28s} {f['launches']:4d} launches {f['ns'] / 1e6:8.3f} ms {100 * f['ns'] / tot:5.1f} %  "
              f"dram read {f['rd'] / 1e9:7.2f} GB  write {f['wr'] / 1e9:7.2f} GB")
    if args.json:
        g = fam["gemm"]
        out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one "
                         f"steady-state C2 step ({args.csv})",
               "gemm_launches": g["launches"], "gemm_dram_bytes_per_step": g["rd"] + g["wr"],
               "gemm_dram_bytes_per_launch": (g["rd"] + g["wr"]) / max(g["launches"], 1)}
        json.dump(out, open(args.json, "w"), indent=1)
        print("wrote", args.json)


if __name__ == "__main__":
    main()
