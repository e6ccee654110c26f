#!/bin/bash
# r6u: evidence pass — per-shape GEMM table at the engine's operand pitch vs cuBLAS; ncu launch list of one steady-state C2
# step with DRAM traffic; ncu --set full (source) of the CTA-pair GEMM (to_qkv, aligned pitch), the persistent
# spatial-attention backward and the spatial-attention forward
set -u
mkdir -p gpurun_out
python tools/gemm_bench.py --aligned --cublas --iters 30 > gpurun_out/r6u_gemm_bench_aligned.log 2>&1
cat gpurun_out/r6u_gemm_bench_aligned.log
timeout 600 ncu --kernel-name-base demangled -k regex:istvt:: -s 516 -c 172 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r6u_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6u_ncu_bench.log 2>&1
python tools/launches_summary.py gpurun_out/r6u_launches.csv --json gpurun_out/r6u_gemm_traffic.json | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 4 -c 1 -o gpurun_out/r6u_gemm_to_qkv \
    python tools/gemm_bench.py --aligned --only to_qkv --iters 2 > gpurun_out/r6u_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial_bwd_tc -s 2 -c 1 -o gpurun_out/r6u_attn_spatial_bwd \
    python tools/attn_bench.py --bwd --iters 2 > gpurun_out/r6u_ncu_attn_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial_pipe -s 2 -c 1 -o gpurun_out/r6u_attn_spatial \
    python tools/attn_bench.py --iters 2 > gpurun_out/r6u_ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
