#!/bin/bash
# r7d (8 GPUs of one box): the multi-GPU configs on the round's final kernels — nn.DataParallel check on two devices,
# C2 inference at N = 2 / 4 / 8, C3 training at N = 8 (and 4), C5 long clips at N = 8, C4 relevance at N = 8
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r7d_gpus.txt 2>&1
CUDA_VISIBLE_DEVICES=0,1 timeout 400 python tools/gpu_check.py --only data_parallel --out gpurun_out/r7d_data_parallel.json --timeout 300 > gpurun_out/r7d_data_parallel.log 2>&1
tail -4 gpurun_out/r7d_data_parallel.log
run() {  # n, name, args...
  n=$1; name=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
}
run 8 r7d_infer_n8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
run 8 r7d_train_b64_n8 --mode train --steps 5 --warmup 3
run 8 r7d_longclip_t32_b8_n8 --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
run 8 r7d_relevance_b32_n8 --mode relevance --steps 5 --warmup 3 --no-cpu-baseline
run 4 r7d_infer_n4 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
run 2 r7d_infer_n2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ('r7d_infer_n8', 'r7d_infer_n4', 'r7d_infer_n2', 'r7d_train_b64_n8', 'r7d_longclip_t32_b8_n8', 'r7d_relevance_b32_n8'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'], 1), d['unit'], round(d['ms_per_step'], 2), (d.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
