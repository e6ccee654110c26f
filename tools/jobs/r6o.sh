#!/bin/bash
# r6o: aligned row pitch (768) for the K = 728 GEMM operands in the engine, single-thread roles in the highest warp ids
# (GEMM pair kernel, spatial attention): parity, A/B per kernel and on the whole C2 step; reduce-add with a 768 stream
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only layernorm,layernorm_diff,gemm_basic,gemm_shapes,gemm_lnfold,attn_spatial_bf16,attn_spatial_spiky,golden_sens_bf16,golden_default_bf16,golden_sens_fp32,batch64,oracle_bf16,cuda_graph,golden_t32_bf16 --out gpurun_out/r6o_checks.json --timeout 400 > gpurun_out/r6o_checks.log 2>&1
tail -16 gpurun_out/r6o_checks.log
{
for r in 1 0 1 0; do
  echo "== ISTVT_G2_ROLES_HIGH=$r"
  ISTVT_G2_ROLES_HIGH=$r python tools/gemm_bench.py --iters 30
done
for r in 1 0 1 0; do
  echo "== ISTVT_SA_ROLES_HIGH=$r"
  ISTVT_SA_ROLES_HIGH=$r python tools/attn_bench.py --iters 30 2>&1 | grep -i spatial
done
python tools/gemm_bench.py --residual-pitch --iters 30
} > gpurun_out/r6o_roles.log 2>&1
cat gpurun_out/r6o_roles.log
for v in "1 1" "0 1" "1 0" "1 1"; do
set -- $v
ISTVT_ROW_PITCH=$1 ISTVT_G2_ROLES_HIGH=$2 ISTVT_SA_ROLES_HIGH=$2 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6o_bench_pitch$1_roles$2.json 2> gpurun_out/r6o_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6o_bench_pitch$1_roles$2.json').read().strip().splitlines()[-1])
print('pitch=$1 roles_high=$2', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3)) for k, v in d['kernels'].items() if k in ('gemm_bf16','layernorm','layernorm_diff','attn_spatial')})
PY
done
python -m pytest tests/ -x -q -m gpu > gpurun_out/r6o_pytest_gpu.log 2>&1
tail -4 gpurun_out/r6o_pytest_gpu.log
