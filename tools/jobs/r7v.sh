#!/bin/bash
# r7v: pp3 with the MUFU rotation (exponential phases of the three groups issued in turn): parity + timing A/B
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky --out gpurun_out/r7v_checks.json --timeout 120 > gpurun_out/r7v_checks.log 2>&1
tail -4 gpurun_out/r7v_checks.log | cut -c1-300
for cfg in "pp3 1" "pp3 0" "pp 1"; do
set -- $cfg
echo "== ISTVT_SA_KERNEL=$1 rotate=$2"
ISTVT_SA_KERNEL=$1 ISTVT_SA_ROTATE=$2 timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
done
ISTVT_SA_KERNEL=pp3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial_pp3 -s 2 -c 1 -o gpurun_out/r7v_attn_spatial_pp3 \
  python tools/attn_bench.py --iters 2 > gpurun_out/r7v_ncu.log 2>&1
