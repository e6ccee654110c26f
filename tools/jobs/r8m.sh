#!/bin/bash
# r8m: 1-CTA GEMM epilogue with the bias loads hoisted in front of the TMEM load: parity + conv2 timing + entry-flow bench
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only conv3x3,gemm_basic,gemm_shapes,xception_bf16,golden_sens_bf16,train_golden --out gpurun_out/r8m_checks.json --timeout 200 > gpurun_out/r8m_checks.log 2>&1
tail -8 gpurun_out/r8m_checks.log | cut -c1-300
timeout 120 python tools/conv_bench.py pair taps 2>&1 | tail -3
timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8m_bench.json 2> gpurun_out/r8m_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r8m_bench.json').read().strip().splitlines()[-1])
print('C2', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('gbs') or 0)) for k, v in d['kernels'].items() if k in ('gemm_bf16','conv3x3','conv_stem')})
PY
