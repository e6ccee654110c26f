#!/bin/bash
# r6s: spatial-attention backward as a PERSISTENT kernel (one CTA per SM walks (frame, head, key chunk) items, next
# (32 key columns per thread) instead of 8: parity, C3 and C4 benches
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only attn_spatial_bwd,train_golden,train_t32_oracle,relevance,relevance_t32 --out gpurun_out/r6s_checks.json --timeout 400 > gpurun_out/r6s_checks.log 2>&1
tail -8 gpurun_out/r6s_checks.log
python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6s_train_b64.json 2> gpurun_out/r6s_train.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6s_train_b64.json').read().strip().splitlines()[-1])
print('train', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],2), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','gemm_wgrad','layernorm_bwd','attn_spatial_bwd','gelu','gelu_bwd')})
PY
python bench.py --mode relevance --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r6s_relevance_b32.json 2>> gpurun_out/r6s_train.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6s_relevance_b32.json').read().strip().splitlines()[-1])
print('relevance', round(d['value'],1), round(d['ms_per_step'],2), {k: (round(v['ms_per_step'],2)) for k, v in d['kernels'].items() if k in ('gemm_bf16','attn_spatial_bwd','layernorm_bwd')})
PY
