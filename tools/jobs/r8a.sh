#!/bin/bash
# r8a: evidence for the final tree — compute-sanitizer (memcheck, racecheck) over the kernels added or changed since r7c
# (online-softmax spatial attention pp / pp3, FFMA2 depthwise kernels), and the ncu launch list + DRAM traffic of one C2 step
set -u
mkdir -p gpurun_out
CHK=attn_spatial_bf16,attn_spatial_spiky,sepconv_fused,dwconv,entry_train_kernels
timeout 330 compute-sanitizer --tool memcheck --kernel-regex kns=istvt --log-file gpurun_out/r8a_memcheck.log \
    python tools/sanitizer_run.py --budget 280 --only $CHK --out gpurun_out/r8a_memcheck_checks.json > gpurun_out/r8a_memcheck_stdout.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r8a_memcheck_stdout.log
tail -7 gpurun_out/r8a_memcheck_stdout.log; tail -3 gpurun_out/r8a_memcheck.log
timeout 330 compute-sanitizer --tool racecheck --kernel-regex kns=istvt --log-file gpurun_out/r8a_racecheck.log \
    python tools/sanitizer_run.py --budget 280 --only $CHK --out gpurun_out/r8a_racecheck_checks.json > gpurun_out/r8a_racecheck_stdout.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r8a_racecheck_stdout.log
tail -7 gpurun_out/r8a_racecheck_stdout.log; tail -3 gpurun_out/r8a_racecheck.log
timeout 400 ncu --kernel-name-base demangled -k regex:istvt:: -s 507 -c 169 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r8a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8a_ncu_bench.log 2>&1
python tools/launches_summary.py gpurun_out/r8a_launches.csv --json gpurun_out/r8a_gemm_traffic.json | head -30
