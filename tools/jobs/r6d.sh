#!/bin/bash
# r6d: bf16 GEMM epilogue storing straight from registers (ISTVT_G2_DIRECT=1, default) vs through the shared-memory
# transpose (=0): parity (gemm_basic / gemm_shapes / goldens), per-shape timing with the clocks seen, C2 bench
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only gemm_basic,gemm_shapes,golden_sens_bf16,golden_default_bf16 --out gpurun_out/r6d_checks.json --timeout 400 > gpurun_out/r6d_checks.log 2>&1
tail -6 gpurun_out/r6d_checks.log
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > gpurun_out/r6d_clocks.csv &
SMI=$!
{
for d in 0 1 0 1; do
  echo "== ISTVT_G2_DIRECT=$d $(date +%s.%N)"
  ISTVT_G2_DIRECT=$d python tools/gemm_bench.py --iters 30
done
echo "== cublas $(date +%s.%N)"
python tools/gemm_bench.py --iters 30 --cublas --only to_qkv,ff1,ff2
} > gpurun_out/r6d_gemm.log 2>&1
kill $SMI
cat gpurun_out/r6d_gemm.log
for d in 0 1; do
ISTVT_G2_DIRECT=$d python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6d_bench_direct$d.json 2> gpurun_out/r6d_bench_direct$d.err
done
python - <<'PY'
import json
for f in ('gpurun_out/r6d_bench_direct0.json', 'gpurun_out/r6d_bench_direct1.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d.get('clocks'), {k: (round(v['ms_per_step'],3), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','attn_spatial')})
PY
