#!/bin/bash
# r6q: MLP fusions of the training step — GELU in ff1's epilogue with two outputs (istvt_gemm_act_dual_fwd), gelu' in the
# ff2 data-gradient GEMM's epilogue (istvt_gemm_dgelu_fwd): parity, C3 bench A/B (ISTVT_MLP_FUSE 1 / 0), full suite
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only gemm_mlp_fusions,gemm_basic,gemm_shapes,gemm_lnfold,train_golden,train_t32_oracle,relevance,golden_sens_bf16 --out gpurun_out/r6q_checks.json --timeout 400 > gpurun_out/r6q_checks.log 2>&1
tail -10 gpurun_out/r6q_checks.log
for v in 1 0 1; do
ISTVT_MLP_FUSE=$v python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6q_train_fuse$v.json 2> gpurun_out/r6q_train.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6q_train_fuse$v.json').read().strip().splitlines()[-1])
print('mlp_fuse=$v', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],2), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','gemm_wgrad','layernorm_bwd','gelu','gelu_bwd','colsum')})
PY
done
python -m pytest tests/ -x -q -m gpu > gpurun_out/r6q_pytest_gpu.log 2>&1
tail -4 gpurun_out/r6q_pytest_gpu.log
