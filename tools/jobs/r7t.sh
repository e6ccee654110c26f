#!/bin/bash
# r7t: pp kernel, whole 128-key chunk per pass (one quarter of the scores parked in local memory by the compiler): parity + timing
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky --out gpurun_out/r7t_checks.json --timeout 120 > gpurun_out/r7t_checks.log 2>&1
tail -4 gpurun_out/r7t_checks.log | cut -c1-300
for k in pp pp3 pipe; do
echo "== ISTVT_SA_KERNEL=$k"
ISTVT_SA_KERNEL=$k timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
done
