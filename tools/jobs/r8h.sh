#!/bin/bash
# r8h: conv2 on pixel pairs (128-byte TMA rows, 7 k-blocks of 64 for two output pixels): parity, timing, C2 bench A/B
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only conv3x3,golden_sens_bf16,batch64,xception_bf16,train_golden --out gpurun_out/r8h_checks.json --timeout 200 > gpurun_out/r8h_checks.log 2>&1
tail -12 gpurun_out/r8h_checks.log | cut -c1-300
timeout 120 python tools/conv_bench.py pair taps 2>&1 | tail -3
for k in pair taps; do
ISTVT_CONV2_KERNEL=$k timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8h_bench_$k.json 2> gpurun_out/r8h_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r8h_bench_$k.json').read().strip().splitlines()[-1])
print('conv2=$k', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('gbs') or 0)) for k, v in d['kernels'].items() if k in ('gemm_bf16','conv3x3','conv_stem')})
PY
done
