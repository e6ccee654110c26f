#!/bin/bash
# r7c: compute-sanitizer on the round-2 kernels (memcheck over every kernel check; racecheck over the kernels that were
# added or restructured this round plus the round-1 set)
set -u
mkdir -p gpurun_out
timeout 560 compute-sanitizer --tool memcheck --kernel-regex kns=istvt --log-file gpurun_out/r7c_memcheck.log \
    python tools/sanitizer_run.py --budget 500 --out gpurun_out/r7c_memcheck_checks.json > gpurun_out/r7c_memcheck_stdout.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r7c_memcheck_stdout.log
tail -4 gpurun_out/r7c_memcheck_stdout.log; tail -3 gpurun_out/r7c_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --kernel-regex kns=istvt --log-file gpurun_out/r7c_racecheck.log \
    python tools/sanitizer_run.py --budget 360 --only sepconv_fused,attn_spatial_bwd,gemm_basic,gemm_lnfold,gemm_mlp_fusions,layernorm,layernorm_diff,layernorm_bwd,conv_stem,conv3x3,attn_temporal,attn_temporal_bwd,attn_spatial_bf16,dwconv,entry_train_kernels \
    --out gpurun_out/r7c_racecheck_checks.json > gpurun_out/r7c_racecheck_stdout.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r7c_racecheck_stdout.log
tail -4 gpurun_out/r7c_racecheck_stdout.log
grep -c "Race reported\|Error:" gpurun_out/r7c_racecheck.log
grep "RACECHECK SUMMARY\|ERROR SUMMARY" gpurun_out/r7c_racecheck.log gpurun_out/r7c_memcheck.log
grep -h "Race reported between\|and .* access at" gpurun_out/r7c_racecheck.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -20 > gpurun_out/r7c_racecheck_summary.txt
cat gpurun_out/r7c_racecheck_summary.txt
