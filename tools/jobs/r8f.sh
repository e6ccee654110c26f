#!/bin/bash
# r8f: LayerNorm backward at the training size: timing per call form + ncu of the fp32-accumulate form; memcheck of the
# rewritten training entry-flow kernels
set -u
mkdir -p gpurun_out
python tools/ln_bench.py 2>&1 | tee gpurun_out/r8f_ln_bench.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd_kernel -s 4 -c 1 -o gpurun_out/r8f_ln_bwd python tools/ln_bench.py --iters 3 > /dev/null 2>&1
timeout 200 compute-sanitizer --tool memcheck --kernel-regex kns=istvt --log-file gpurun_out/r8f_memcheck.log \
    python tools/sanitizer_run.py --budget 150 --only entry_train_kernels --out gpurun_out/r8f_memcheck_checks.json 2>&1 | tail -2
tail -2 gpurun_out/r8f_memcheck.log
