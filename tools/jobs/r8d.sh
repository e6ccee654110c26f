#!/bin/bash
# r8d: + im2col^T kernel rewritten (256 pixels x 9 taps per CTA, conflict-free transposition)
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only entry_train_kernels,train_golden --out gpurun_out/r8d_checks.json --timeout 200 > gpurun_out/r8d_checks.log 2>&1
tail -4 gpurun_out/r8d_checks.log | cut -c1-250
timeout 300 python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8d_train_b64.json 2> gpurun_out/r8d_train.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r8d_train_b64.json').read().strip().splitlines()[-1])
print('train', round(d['value'],1), round(d['ms_per_step'],2), {k: (round(v['ms_per_step'],2), round(v.get('gbs') or 0)) for k, v in d['kernels'].items() if k in ('pool_bwd','pool_add','im2col_t','bn_bwd','layernorm_bwd','attn_spatial','attn_spatial_bwd')})
PY
