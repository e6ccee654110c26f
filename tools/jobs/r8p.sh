#!/bin/bash
# r8p: ff1 epilogue with the GELU on packed fp32 pairs: parity (bit-identical canary: golden logits 5.92e-03), ff1 timing, C2 bench
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only gemm_basic,gemm_shapes,gemm_mlp_fusions,golden_sens_bf16,train_golden --out gpurun_out/r8p_checks.json --timeout 200 > gpurun_out/r8p_checks.log 2>&1
tail -7 gpurun_out/r8p_checks.log | cut -c1-300
timeout 200 python tools/gemm_bench.py --aligned --iters 30 2>&1 | tail -9
timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8p_bench.json 2> gpurun_out/r8p_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r8p_bench.json').read().strip().splitlines()[-1])
print('C2', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('tflops') or 0)) for k, v in d['kernels'].items() if k in ('gemm_bf16','attn_spatial')})
PY
