#!/bin/bash
# r7i: fused SeparableConv2d, fifth version (depthwise multiply-adds as packed fp32 FFMA2: two per issue slot
#): parity, per-layer timing, C2 bench A/B
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only sepconv_fused,golden_sens_bf16,batch64,xception_bf16 --out gpurun_out/r7i_checks.json --timeout 200 > gpurun_out/r7i_checks.log 2>&1
tail -6 gpurun_out/r7i_checks.log
python tools/sep_bench.py > gpurun_out/r7i_sep_bench.log 2>&1
cat gpurun_out/r7i_sep_bench.log
for v in 1 0; do
ISTVT_SEP_FUSE=$v timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7i_bench_fuse$v.json 2> gpurun_out/r7i_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r7i_bench_fuse$v.json').read().strip().splitlines()[-1])
print('sep_fuse=$v', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','sepconv_fused','dwconv3x3','pool_add','conv3x3')})
PY
done
