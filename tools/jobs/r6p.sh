#!/bin/bash
# r6p: training side — 128-byte aligned row pitch for every bf16 [rows, 728] operand of the training step (activations,
# gradients, weights, transposed weights; ISTVT_ROW_PITCH 1 / 0), layernorm_bwd with L2 prefetch of the next row
# (ISTVT_LNB_PREFETCH 1 / 0), pool_bwd with a 3-D grid instead of 64-bit index divisions; persisted pack cache check
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only layernorm_bwd,gelu_cast_transpose,gemm_wgrad,entry_train_kernels,train_golden,train_t32_oracle,relevance,relevance_t32,pack_cache,api --out gpurun_out/r6p_checks.json --timeout 400 > gpurun_out/r6p_checks.log 2>&1
tail -12 gpurun_out/r6p_checks.log
for v in "1 1" "1 0" "0 1"; do
set -- $v
ISTVT_ROW_PITCH=$1 ISTVT_LNB_PREFETCH=$2 python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6p_train_pitch$1_pf$2.json 2> gpurun_out/r6p_train.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6p_train_pitch$1_pf$2.json').read().strip().splitlines()[-1])
print('pitch=$1 prefetch=$2', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],2), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','gemm_wgrad','layernorm_bwd','pool_bwd','attn_spatial_bwd','colsum','cast','transpose')})
PY
done
ISTVT_ROW_PITCH=1 python bench.py --mode relevance --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r6p_relevance_b32.json 2>> gpurun_out/r6p_train.err
tail -c 600 gpurun_out/r6p_relevance_b32.json
