#!/bin/bash
# r6w: fused SeparableConv2d kernel (depthwise result produced into the pointwise GEMM's shared-memory A tiles) for
# entry-flow blocks 1 / 2: parity, C2 bench A/B (ISTVT_SEP_FUSE 1 / 0); spatial-attention backward after the D / lse
# prefetch
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only sepconv_fused --out gpurun_out/r6w_check_kernel.json --timeout 200 > gpurun_out/r6w_check_kernel.log 2>&1
tail -12 gpurun_out/r6w_check_kernel.log
timeout 600 python tools/gpu_check.py --only golden_sens_bf16,golden_default_bf16,batch64,oracle_bf16,xception_bf16,uint8_input,cuda_graph,golden_t32_bf16 --out gpurun_out/r6w_checks.json --timeout 300 > gpurun_out/r6w_checks.log 2>&1
tail -10 gpurun_out/r6w_checks.log
for v in 1 0 1; do
ISTVT_SEP_FUSE=$v timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6w_bench_fuse$v.json 2> gpurun_out/r6w_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6w_bench_fuse$v.json').read().strip().splitlines()[-1])
print('sep_fuse=$v', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('gemm_bf16','sepconv_fused','dwconv3x3','pool_add','conv3x3')})
PY
done
python tools/attn_bench.py --bwd --iters 30 2>&1 | grep -i bwd
