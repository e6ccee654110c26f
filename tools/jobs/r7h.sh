#!/bin/bash
# r7h: conv2 on the strip pipeline (TMA slab -> shared-memory im2col -> tcgen05 -> 4-D TMA store): parity of the three
# conv2 kernels, stand-alone timing, C2 bench A/B (ISTVT_CONV2_KERNEL strip / taps)
set -u
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py --only conv3x3,entry_train_kernels,golden_sens_bf16,golden_default_bf16,batch64,xception_bf16,train_golden --out gpurun_out/r7h_checks.json --timeout 300 > gpurun_out/r7h_checks.log 2>&1
tail -9 gpurun_out/r7h_checks.log
for k in strip taps strip taps; do echo "ISTVT_CONV2_KERNEL=$k"; ISTVT_CONV2_KERNEL=$k timeout 120 python tools/conv_bench.py 2>&1 | tail -1; done
for k in strip taps; do
ISTVT_CONV2_KERNEL=$k timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7h_bench_$k.json 2> gpurun_out/r7h_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r7h_bench_$k.json').read().strip().splitlines()[-1])
print('conv2=$k', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','conv3x3','conv_stem','sepconv_fused')})
PY
done
