#!/bin/bash
# r6x: per-layer timing of the fused SeparableConv2d kernel vs the two-kernel path; ncu --set full of the 128 -> 128 layer
set -u
mkdir -p gpurun_out
python tools/sep_bench.py > gpurun_out/r6x_sep_bench.log 2>&1
cat gpurun_out/r6x_sep_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sepconv_fused -s 2 -c 1 -o gpurun_out/r6x_sepconv_fused \
    python tools/sep_bench.py --only b1_sep2 --iters 1 > gpurun_out/r6x_ncu.log 2>&1
python tools/attn_bench.py --bwd --iters 30 2>&1 | grep -i bwd
ls -la gpurun_out/r6x_sepconv_fused.ncu-rep
