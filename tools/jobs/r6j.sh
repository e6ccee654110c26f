#!/bin/bash
# r6j: stem conv1 as a TF32 implicit GEMM on tcgen05 (ISTVT_STEM_TC=1 default) + LayerNorm-2 fold v3 (finalize kernel)
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only conv_stem,conv_stem_u8,gemm_lnfold,golden_sens_bf16,golden_default_bf16,golden_t32_bf16,oracle_bf16,batch64,uint8_input,cuda_graph,train_golden,xception_bf16 --out gpurun_out/r6j_checks.json --timeout 400 > gpurun_out/r6j_checks.log 2>&1
tail -14 gpurun_out/r6j_checks.log
for tc in 0 1; do echo "ISTVT_STEM_TC=$tc"; ISTVT_STEM_TC=$tc python tools/entry_bench.py 2>&1 | grep conv_stem; done
for cfg in "0 0" "1 0" "0 1" "1 1" "1 1"; do
set -- $cfg
ISTVT_STEM_TC=$1 ISTVT_LN2_FOLD=$2 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6j_bench_tc$1_fold$2.json 2> gpurun_out/r6j_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6j_bench_tc$1_fold$2.json').read().strip().splitlines()[-1])
print('stem_tc=$1 fold=$2', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('e2e_uint8',{}).get('value'), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3)) for k, v in d['kernels'].items() if k in ('gemm_bf16','layernorm','conv_stem')})
PY
done
