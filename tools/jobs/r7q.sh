#!/bin/bash
# r7q: online-softmax spatial attention as the inference default (round-to-nearest P), exact-max kernel kept for the lse modes
set -u
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky,attn_spatial_bwd,golden_sens_bf16,oracle_bf16,batch64,golden_t32_bf16,train_golden,relevance,relevance_t32,cuda_graph --out gpurun_out/r7q_checks.json --timeout 200 > gpurun_out/r7q_checks.log 2>&1
tail -13 gpurun_out/r7q_checks.log
timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7q_bench.json 2> gpurun_out/r7q_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r7q_bench.json').read().strip().splitlines()[-1])
print('C2', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('tflops') or 0)) for k, v in d['kernels'].items() if k in ('gemm_bf16','attn_spatial')})
PY
