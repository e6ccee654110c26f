#!/bin/bash
# r8b: ncu of the stand-alone depthwise kernel (after FFMA2) at the two layer shapes that stay unfused in C2
set -u
mkdir -p gpurun_out
python tools/entry_bench.py --iters 10 2>&1 | tee gpurun_out/r8b_entry_bench.log | head -20
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_strip -s 42 -c 1 -o gpurun_out/r8b_dw_74x256 python tools/entry_bench.py --iters 10 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_strip -s 68 -c 1 -o gpurun_out/r8b_dw_37x728 python tools/entry_bench.py --iters 10 > /dev/null 2>&1
ls -la gpurun_out/r8b_*.ncu-rep
