#!/bin/bash
# r6i: LayerNorm-2 fold with per-group (sum, M2) statistics (no atomics): parity + determinism, C2 bench A/B
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only gemm_lnfold,gemm_shapes,golden_sens_bf16,golden_default_bf16,golden_t32_bf16,oracle_bf16,batch64,uint8_input,cuda_graph,train_golden --out gpurun_out/r6i_checks.json --timeout 400 > gpurun_out/r6i_checks.log 2>&1
tail -13 gpurun_out/r6i_checks.log
for d in 0 1 0 1; do
ISTVT_LN2_FOLD=$d python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6i_bench_fold${d}.json 2> gpurun_out/r6i_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6i_bench_fold${d}.json').read().strip().splitlines()[-1])
print('fold=$d', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3)) for k, v in d['kernels'].items() if k in ('gemm_bf16','layernorm','attn_spatial')})
PY
done
