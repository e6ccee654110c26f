#!/bin/bash
# r8l: ncu of the pixel-pair conv2 kernel, and the launch list + DRAM traffic of one C2 step on the final tree
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 4 -c 1 -o gpurun_out/r8l_conv2_pair python tools/conv_bench.py pair > gpurun_out/r8l_ncu_conv.log 2>&1
timeout 400 ncu --kernel-name-base demangled -k regex:istvt:: -s 510 -c 170 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r8l_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8l_ncu_bench.log 2>&1
python tools/launches_summary.py gpurun_out/r8l_launches.csv --json gpurun_out/r8l_gemm_traffic.json > gpurun_out/r8l_launches_summary.txt 2>&1
head -16 gpurun_out/r8l_launches_summary.txt
ls -la gpurun_out/r8l_conv2_pair.ncu-rep
