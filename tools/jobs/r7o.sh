#!/bin/bash
# r7o: is the spatial relevance map sensitive to the P rounding of the new attention kernel?  relevance check under three kernels
set -u
mkdir -p gpurun_out
for cfg in "pipe x" "pp x" "pp rne"; do
set -- $cfg
echo "== kernel=$1 round=$2"
ISTVT_SA_KERNEL=$1 ISTVT_SA_ROUND=$2 timeout 200 python tools/gpu_check.py --only relevance,relevance_t32 --out gpurun_out/r7o_$1_$2.json --timeout 120 2>&1 | grep -E "ok  |FAIL|Assert" | cut -c1-250
python - <<PY
import json
for e in json.load(open('gpurun_out/r7o_$1_$2.json')):
    print(e['name'], e.get('metrics') or e.get('error','')[:200])
PY
done
ISTVT_SA_ROUND=rne timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
