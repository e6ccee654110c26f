#!/bin/bash
# r6l: conv2 with a SIMT-gathered A operand (ISTVT_CONV2_TC=1 default) parity + A/B; TMA-store epilogue on 8 vs 16
# epilogue warps; full GPU suite on the resulting tree
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only conv3x3,entry_train_kernels,golden_sens_bf16,golden_default_bf16,train_golden,xception_bf16 --out gpurun_out/r6l_checks.json --timeout 400 > gpurun_out/r6l_checks.log 2>&1
tail -8 gpurun_out/r6l_checks.log
for tc in 0 1; do echo "ISTVT_CONV2_TC=$tc"; ISTVT_CONV2_TC=$tc python tools/conv_bench.py 2>&1 | tail -1; done
{
for ew in 16 8 16 8; do
  echo "== ISTVT_G2_EPI_WARPS=$ew"
  ISTVT_G2_EPI_WARPS=$ew python tools/gemm_bench.py --iters 30 --only to_qk,to_v,t_out,to_qkv,ff1
done
} > gpurun_out/r6l_gemm_ew.log 2>&1
cat gpurun_out/r6l_gemm_ew.log
for tc in 0 1 1; do
ISTVT_CONV2_TC=$tc python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6l_bench_conv2tc$tc.json 2> gpurun_out/r6l_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6l_bench_conv2tc$tc.json').read().strip().splitlines()[-1])
print('conv2_tc=$tc', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3)) for k, v in d['kernels'].items() if k in ('gemm_bf16','layernorm','conv_stem','conv3x3')})
PY
done
python -m pytest tests/ -x -q -m gpu > gpurun_out/r6l_pytest_gpu.log 2>&1
tail -4 gpurun_out/r6l_pytest_gpu.log
