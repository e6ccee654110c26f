#!/bin/bash
# r8t: compute-sanitizer over the conv2 kernels added after r8a (pixel pairs, resident weights, hoisted bias loads)
set -u
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --kernel-regex kns=istvt --log-file gpurun_out/r8t_memcheck.log \
    python tools/sanitizer_run.py --budget 120 --only conv3x3,gemm_basic --out gpurun_out/r8t_memcheck_checks.json 2>&1 | tail -3
tail -2 gpurun_out/r8t_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --kernel-regex kns=istvt --log-file gpurun_out/r8t_racecheck.log \
    python tools/sanitizer_run.py --budget 170 --only conv3x3 --out gpurun_out/r8t_racecheck_checks.json 2>&1 | tail -2
tail -3 gpurun_out/r8t_racecheck.log
grep -c "hazard" gpurun_out/r8t_racecheck.log
