#!/bin/bash
# r7e: bias gradients taken by the kernels that produce the output gradients (layernorm_bwd out_colsum, gelu_bwd + column
# sums) instead of stand-alone column-sum passes: parity, C3 bench; engine fingerprint cache (batch-1 latency)
set -u
mkdir -p gpurun_out
timeout 500 python tools/gpu_check.py --only layernorm_bwd,gelu_cast_transpose,train_golden,train_t32_oracle,relevance,cuda_graph,api --out gpurun_out/r7e_checks.json --timeout 300 > gpurun_out/r7e_checks.log 2>&1
tail -9 gpurun_out/r7e_checks.log
python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7e_train_b64.json 2> gpurun_out/r7e_train.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r7e_train_b64.json').read().strip().splitlines()[-1])
print('train', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],2), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','gemm_wgrad','layernorm_bwd','attn_spatial_bwd','gelu','gelu_bwd','colsum')})
PY
