#!/bin/bash
# r7j: packed-fp32 multiply-adds (FFMA2) in the stand-alone depthwise strip kernel as well: parity (forward + training), C2 bench,
# training step
set -u
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py --only dwconv,sepconv_fused,entry_train_kernels,golden_sens_bf16,batch64,xception_bf16,train_golden --out gpurun_out/r7j_checks.json --timeout 200 > gpurun_out/r7j_checks.log 2>&1
tail -9 gpurun_out/r7j_checks.log
for v in 1 0; do
ISTVT_SEP_FUSE=$v timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7j_bench_fuse$v.json 2> gpurun_out/r7j_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r7j_bench_fuse$v.json').read().strip().splitlines()[-1])
print('sep_fuse=$v', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','sepconv_fused','dwconv3x3','pool_add','conv3x3')})
PY
done
timeout 300 python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7j_train_b64.json 2> gpurun_out/r7j_train.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r7j_train_b64.json').read().strip().splitlines()[-1])
print('train', round(d['value'],1), round(d['ms_per_step'],2), {k: (round(v['ms_per_step'],2), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('dwconv3x3','sepconv_fused','gemm_bf16','gemm_wgrad')})
PY
