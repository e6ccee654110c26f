"""Read `ncu -i X.ncu-rep --page source --csv` output and print the instructions with the most warp-stall samples,
their top stall reasons and an opcode histogram of executed instructions (weighted by executions).

    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv; python tools/ncu_src_top.py /tmp/src.csv [N]
"""
import csv
import sys
from collections import Counter


def main() -> None:
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]       # first captured launch only
    sk, ex = ci["# Samples"], ci["Instructions Executed"]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[sk] or 0) for r in body)
    print(f"# kernel: {rows[0][1][:100]}")
    print(f"# total warp-stall samples {int(tot)}; instructions {len(body)}")
    agg = Counter()
    for r in body:
        for h in stall_cols:
            agg[h] += float(r[ci[h]] or 0)
    print("# stall reasons (all samples):", ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for h, v in agg.most_common(9)))
    ops = Counter()
    for r in body:
        op = r[ci["Source"]].split()
        op = [o for o in op if not o.startswith("@")]
        if op:
            ops[op[0].split(".")[0]] += float(r[ex] or 0)
    te = sum(ops.values())
    print("# executed warp-instructions by opcode:", ", ".join(f"{k} {100 * v / te:.1f}%" for k, v in ops.most_common(14)))
    for idx, r in sorted(enumerate(body), key=lambda kv: -float(kv[1][sk] or 0))[:topn]:
        st = sorted(((float(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{idx:5d} {int(float(r[sk])):6d} {100 * float(r[sk]) / tot:5.1f}% ex={r[ex]:>8s} {r[ci['Source']].strip()[:72]:72s} "
              f"{[(h, int(v)) for v, h in st if v > 0]}")


if __name__ == "__main__":
    main()
