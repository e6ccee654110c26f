#!/bin/bash
# r6t: persistent spatial-attention backward — item order (one (frame, head) per CTA at a time vs round-robin) and what
# the red.global accumulation of dQ costs (ISTVT_SAB_NORED=1: timing only, wrong dQ)
set -u
mkdir -p gpurun_out
python tools/gpu_check.py --only attn_spatial_bwd,train_golden,relevance --out gpurun_out/r6t_checks.json --timeout 400 > gpurun_out/r6t_checks.log 2>&1
tail -5 gpurun_out/r6t_checks.log
{
for v in "1 0" "0 0" "1 1" "0 1" "1 0"; do
  set -- $v
  echo "== ISTVT_SAB_GROUPED=$1 ISTVT_SAB_NORED=$2"
  ISTVT_SAB_GROUPED=$1 ISTVT_SAB_NORED=$2 python tools/attn_bench.py --bwd --iters 30 2>&1 | grep -i "bwd"
done
} > gpurun_out/r6t_attn_bwd.log 2>&1
cat gpurun_out/r6t_attn_bwd.log
