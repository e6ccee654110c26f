#!/bin/bash
# r6a: safety evidence on the round-1 code (compute-sanitizer memcheck + racecheck over tests/kernel_checks.py),
# ncu source capture of the stem kernels, and the per-shape GEMM yardstick (ours vs cuBLAS) on this box.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r6a_smi.txt 2>&1
python tools/gemm_bench.py --cublas > gpurun_out/r6a_gemm_bench.log 2>&1
python tools/entry_bench.py > gpurun_out/r6a_entry_bench.log 2>&1
python tools/conv_bench.py >> gpurun_out/r6a_entry_bench.log 2>&1
timeout 420 compute-sanitizer --tool memcheck --kernel-regex kns=istvt --log-file gpurun_out/r6a_memcheck.log \
    python tools/sanitizer_run.py --budget 360 --out gpurun_out/r6a_memcheck_checks.json > gpurun_out/r6a_memcheck_stdout.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r6a_memcheck_stdout.log
timeout 330 compute-sanitizer --tool racecheck --kernel-regex kns=istvt --log-file gpurun_out/r6a_racecheck.log \
    python tools/sanitizer_run.py --budget 270 --only layernorm,layernorm_diff,gemm_basic,dwconv,pool_subsample_tokens,attn_temporal,attn_spatial_bf16,attn_joint,conv3x3,conv_stem,head,layernorm_bwd,attn_spatial_bwd,gemm_wgrad,entry_train_kernels \
    --out gpurun_out/r6a_racecheck_checks.json > gpurun_out/r6a_racecheck_stdout.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r6a_racecheck_stdout.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv_stem -c 2 -o gpurun_out/r6a_stem \
    python tools/entry_bench.py --iters 1 > gpurun_out/r6a_ncu_stem.log 2>&1
tail -5 gpurun_out/r6a_gemm_bench.log gpurun_out/r6a_memcheck_stdout.log gpurun_out/r6a_racecheck_stdout.log
