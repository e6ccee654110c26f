#!/bin/bash
# r6m: what the bf16 TMA-store epilogue of the CTA-pair GEMM costs, piece by piece (debug build -DISTVT_GEMM_TRACE;
# ISTVT_TRACE_NOEPI: 1 release only, 2 TMEM reads only, 5 reads + math, 3 reads + math + staging, 4 staging + TMA store
# without TMEM reads, 0 everything), with and without concurrent TMA fills (ISTVT_TRACE_NOTMA=1)
set -u
mkdir -p gpurun_out
ISTVT_BUILD_DEFS=-DISTVT_GEMM_TRACE python 2023-tifs-istvt_b200/build.py --force > gpurun_out/r6m_build.log 2>&1
{
for tma in 0 1; do
for epi in 0 1 2 5 3 4; do
  echo "== ISTVT_TRACE_NOTMA=$tma ISTVT_TRACE_NOEPI=$epi"
  ISTVT_TRACE_NOTMA=$tma ISTVT_TRACE_NOEPI=$epi python tools/gemm_bench.py --only to_qkv,ff1,to_v
done
done
} > gpurun_out/r6m_gemm_epilogue_parts.log 2>&1
cat gpurun_out/r6m_gemm_epilogue_parts.log
