#!/bin/bash
# r7m: spatial attention forward, new kernel v2 (PV wait behind the exponentials, rare rescale before them): parity, A/B, ncu source
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only attn_spatial_bf16,attn_spatial_spiky,attn_spatial_bwd --out gpurun_out/r7m_checks.json --timeout 120 > gpurun_out/r7m_checks.log 2>&1
tail -5 gpurun_out/r7m_checks.log
for k in pp pipe; do
echo "== ISTVT_SA_KERNEL=$k"
ISTVT_SA_KERNEL=$k timeout 120 python tools/attn_bench.py --iters 30 2>&1 | grep attn_spatial
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial_pp -s 2 -c 1 -o gpurun_out/r7m_attn_spatial_pp \
  python tools/attn_bench.py --iters 2 > gpurun_out/r7m_ncu.log 2>&1
ls -la gpurun_out/r7m_attn_spatial_pp.ncu-rep
