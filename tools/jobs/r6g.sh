#!/bin/bash
# r6g: LayerNorm-2 fold (row statistics from to_out's epilogue, gamma/beta/mu/rstd applied in to_qkv's epilogue):
# parity, C2 bench A/B (ISTVT_LN2_FOLD=0/1); temporal tail kernel at 4 CTAs/SM A/B (ISTVT_TA_OCC4) on C5
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only gemm_lnfold,gemm_shapes,golden_sens_bf16,golden_default_bf16,golden_t32_bf16,oracle_bf16,batch64,uint8_input,cuda_graph,relevance,attn_temporal --out gpurun_out/r6g_checks.json --timeout 400 > gpurun_out/r6g_checks.log 2>&1
tail -14 gpurun_out/r6g_checks.log
for d in 0 1 0 1; do
ISTVT_LN2_FOLD=$d python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6g_bench_fold${d}.json 2> gpurun_out/r6g_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6g_bench_fold${d}.json').read().strip().splitlines()[-1])
print('fold=$d', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],3)) for k, v in d['kernels'].items() if k in ('gemm_bf16','layernorm','attn_spatial')})
PY
done
for d in 0 1; do
ISTVT_TA_OCC4=$d python bench.py --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6g_c5_occ${d}.json 2>> gpurun_out/r6g_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r6g_c5_occ${d}.json').read().strip().splitlines()[-1])
print('occ4=$d', d['value'], d['ms_per_step'], {k: (round(v['ms_per_step'],3), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('attn_temporal','gemm_bf16','layernorm')})
PY
done
