#!/bin/bash
# r6e: bf16 GEMM epilogue with TMA box stores (ISTVT_G2_TMASTORE=1, default) vs per-lane LDS + STG (=0)
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only gemm_basic,gemm_shapes,golden_sens_bf16,golden_default_bf16,train_golden,xception_bf16 --out gpurun_out/r6e_checks.json --timeout 400 > gpurun_out/r6e_checks.log 2>&1
tail -8 gpurun_out/r6e_checks.log
{
for d in 0 1 0 1; do
  echo "== ISTVT_G2_TMASTORE=$d"
  ISTVT_G2_TMASTORE=$d python tools/gemm_bench.py --iters 30
done
echo "== entry flow GEMMs, TMASTORE=0 / 1"
ISTVT_G2_TMASTORE=0 python tools/gemm_bench.py --entry
ISTVT_G2_TMASTORE=1 python tools/gemm_bench.py --entry
} > gpurun_out/r6e_gemm.log 2>&1
cat gpurun_out/r6e_gemm.log
for d in 0 1; do
ISTVT_G2_TMASTORE=$d python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6e_bench_tma$d.json 2> gpurun_out/r6e_bench_tma$d.err
done
python - <<'PY'
import json
for f in ('gpurun_out/r6e_bench_tma0.json', 'gpurun_out/r6e_bench_tma1.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d.get('clocks'), {k: (round(v['ms_per_step'],3), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','attn_spatial')})
PY
