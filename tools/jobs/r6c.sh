#!/bin/bash
# r6c: re-check the frame-tiled temporal attention (fp16 P.V, tail variant), C5 bench, then GEMM timing experiments in a
# debug build (-DISTVT_GEMM_TRACE): which operand's TMA fill / how much of the epilogue is in the K = 728 tile time
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only attn_temporal,golden_t32_bf16,golden_t32_fp32,train_t32_oracle,relevance_t32 --out gpurun_out/r6c_checks.json --timeout 400 > gpurun_out/r6c_checks.log 2>&1
tail -8 gpurun_out/r6c_checks.log
python bench.py --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6c_longclip_t32_b8.json 2> gpurun_out/r6c_longclip.err
ISTVT_TA_TAIL=0 python bench.py --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6c_longclip_t32_b8_notail.json 2>> gpurun_out/r6c_longclip.err
python - <<'PY'
import json
for f in ('gpurun_out/r6c_longclip_t32_b8.json', 'gpurun_out/r6c_longclip_t32_b8_notail.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], {k: (round(v['ms_per_step'],3), round(v.get('gbs',0))) for k, v in d['kernels'].items() if k in ('attn_temporal','gemm_bf16','attn_spatial')})
PY
ISTVT_BUILD_DEFS=-DISTVT_GEMM_TRACE python 2023-tifs-istvt_b200/build.py --force > gpurun_out/r6c_build.log 2>&1
{
for v in "0 0" "1 0" "2 0" "3 0" "0 1" "1 1"; do
  set -- $v
  echo "== ISTVT_TRACE_NOTMA=$1 ISTVT_TRACE_NOEPI=$2"
  ISTVT_TRACE_NOTMA=$1 ISTVT_TRACE_NOEPI=$2 python tools/gemm_bench.py --only to_qkv,ff1,ff2,to_v
done
} > gpurun_out/r6c_gemm_experiments.log 2>&1
cat gpurun_out/r6c_gemm_experiments.log
