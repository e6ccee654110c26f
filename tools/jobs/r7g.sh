#!/bin/bash
# r7g: the full GPU suite on the final tree (after the train-mode T = 32 logit tolerance was set from its measured spread)
set -u
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu > gpurun_out/r7g_pytest_gpu.log 2>&1
tail -5 gpurun_out/r7g_pytest_gpu.log
for i in 1 2 3; do python tools/gpu_check.py --only train_t32_oracle --out gpurun_out/r7g_t32_$i.json --timeout 300 2>&1 | grep -i "profile\|ok \|FAIL" | cut -c1-260; done
