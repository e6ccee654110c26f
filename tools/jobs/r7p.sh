#!/bin/bash
# r7p: new attention kernel with the normalisation of the truncating variant fixed: kernel error profile and relevance A/B
set -u
mkdir -p gpurun_out
for cfg in "pp x" "pp rne"; do
set -- $cfg
echo "== kernel=$1 round=$2"
ISTVT_SA_KERNEL=$1 ISTVT_SA_ROUND=$2 timeout 300 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky,attn_spatial_bwd,relevance,relevance_t32 --out gpurun_out/r7p_$1_$2.json --timeout 120 2>&1 | grep -E "ok  |FAIL|Assert" | cut -c1-250
python - <<PY
import json
for e in json.load(open('gpurun_out/r7p_$1_$2.json')):
    m = e.get('metrics') or {}
    if e['name'] == 'attn_spatial_bf16':
        print('pipe_*', {k: round(v, 5) for k, v in m.items() if k.startswith('pipe_') and 'map' not in k})
    elif e['name'].startswith('relevance'):
        print(e['name'], {k: round(v, 4) for k, v in m.items()} or e.get('error', '')[:200])
PY
done
