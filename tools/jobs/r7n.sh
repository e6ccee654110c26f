#!/bin/bash
# r7n: spatial attention forward v3 (exp loop restructured) in the model: kernel + model parity, C2 bench A/B, training step
set -u
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky,attn_spatial_bwd,golden_sens_bf16,oracle_bf16,batch64,golden_t32_bf16,train_golden,train_t32_oracle,relevance --out gpurun_out/r7n_checks.json --timeout 200 > gpurun_out/r7n_checks.log 2>&1
tail -12 gpurun_out/r7n_checks.log
ISTVT_SA_KERNEL=pp timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
for k in pp pipe; do
ISTVT_SA_KERNEL=$k timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7n_bench_$k.json 2> gpurun_out/r7n_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r7n_bench_$k.json').read().strip().splitlines()[-1])
print('sa=$k', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('tflops') or 0)) for k, v in d['kernels'].items() if k in ('gemm_bf16','attn_spatial')})
PY
done
