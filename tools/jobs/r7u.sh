#!/bin/bash
# r7u: pp (two halves per 128-key chunk, pairwise exp helper) vs pp3 (early P stores): parity + timing, ncu of pp3
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky --out gpurun_out/r7u_checks.json --timeout 120 > gpurun_out/r7u_checks.log 2>&1
tail -4 gpurun_out/r7u_checks.log | cut -c1-300
for k in pp pp3 pipe pp pp3; do
echo "== ISTVT_SA_KERNEL=$k"
ISTVT_SA_KERNEL=$k timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
done
ISTVT_SA_KERNEL=pp3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial_pp3 -s 2 -c 1 -o gpurun_out/r7u_attn_spatial_pp3 \
  python tools/attn_bench.py --iters 2 > gpurun_out/r7u_ncu.log 2>&1
ls -la gpurun_out/r7u_attn_spatial_pp3.ncu-rep
