#!/bin/bash
# r7l: spatial attention forward, new kernel (two tiles in flight, thread = row, online softmax with lazy rescale): parity + timing A/B
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky,attn_spatial_bwd --out gpurun_out/r7l_checks.json --timeout 120 > gpurun_out/r7l_checks.log 2>&1
tail -8 gpurun_out/r7l_checks.log
for k in pp pipe; do
echo "== ISTVT_SA_KERNEL=$k"
ISTVT_SA_KERNEL=$k timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
done
