#!/bin/bash
# r6b: parity of the frame-tiled temporal attention kernels (fwd/bwd), the new model-level checks, C5 bench
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only attn_temporal,attn_temporal_bwd,batch64,train_golden,train_t32_oracle,relevance,relevance_t32,golden_t32_bf16,golden_sens_bf16,api --out gpurun_out/r6b_checks.json --timeout 400 > gpurun_out/r6b_checks.log 2>&1
python bench.py --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6b_longclip_t32_b8.json 2> gpurun_out/r6b_longclip.err
tail -15 gpurun_out/r6b_checks.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r6b_longclip_t32_b8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k: (round(v['ms_per_step'],3), round(v.get('gbs',0))) for k, v in d['kernels'].items()})
PY
