#!/bin/bash
# r7a: ncu --set full of the fused SeparableConv2d kernel on the 64 -> 128 layer (epilogue-paced)
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sepconv_fused -s 2 -c 1 -o gpurun_out/r7a_sepconv_fused_b1s1 \
    python tools/sep_bench.py --only b1_sep1 --iters 1 > gpurun_out/r7a_ncu.log 2>&1
ls -la gpurun_out/r7a_sepconv_fused_b1s1.ncu-rep
