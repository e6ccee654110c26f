#!/bin/bash
# r6n: bf16 epilogue with 64-column (128-byte) TMA store boxes (ISTVT_G2_STORE128, default 1) — parity, A/B per shape;
# row-pitch experiment (K = 728 operands at pitch 728 vs 768); conv3x3 check with both conv2 kernels
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only gemm_basic,gemm_shapes,gemm_lnfold,conv3x3,golden_sens_bf16,golden_default_bf16,batch64 --out gpurun_out/r6n_checks.json --timeout 400 > gpurun_out/r6n_checks.log 2>&1
tail -9 gpurun_out/r6n_checks.log
{
for s in 1 0 1 0; do
  echo "== ISTVT_G2_STORE128=$s"
  ISTVT_G2_STORE128=$s python tools/gemm_bench.py --iters 30 --only to_qk,to_v,t_out,to_qkv,ff1
done
echo "== entry flow, STORE128=1 / 0"
ISTVT_G2_STORE128=1 python tools/gemm_bench.py --entry
ISTVT_G2_STORE128=0 python tools/gemm_bench.py --entry
} > gpurun_out/r6n_gemm_store128.log 2>&1
cat gpurun_out/r6n_gemm_store128.log
python tools/gemm_bench.py --pitch --iters 30 > gpurun_out/r6n_gemm_pitch.log 2>&1
cat gpurun_out/r6n_gemm_pitch.log
