#!/bin/bash
# r7r: spatial attention with three query tiles in flight (attn_spatial_pp3.cuh): parity, timing A/B, ncu source
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky --out gpurun_out/r7r_checks.json --timeout 120 > gpurun_out/r7r_checks.log 2>&1
tail -6 gpurun_out/r7r_checks.log | cut -c1-300
for k in pp3 pp pipe; do
echo "== ISTVT_SA_KERNEL=$k"
ISTVT_SA_KERNEL=$k timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial_pp3 -s 2 -c 1 -o gpurun_out/r7r_attn_spatial_pp3 \
  python tools/attn_bench.py --iters 2 > gpurun_out/r7r_ncu.log 2>&1
ls -la gpurun_out/r7r_attn_spatial_pp3.ncu-rep
