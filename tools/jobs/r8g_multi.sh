#!/bin/bash
# r8g (8 GPUs of one box): the multi-GPU lines on the final tree — nn.DataParallel check on two devices, C2 inference,
# C3 training and C5 long clips at N = 8
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0,1 timeout 200 python tools/gpu_check.py --only data_parallel --out gpurun_out/r8g_data_parallel.json --timeout 150 > gpurun_out/r8g_data_parallel.log 2>&1
tail -3 gpurun_out/r8g_data_parallel.log | cut -c1-200
run() {  # n, name, args...
  n=$1; name=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
}
run 8 r8g_infer_n8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
run 8 r8g_train_b64_n8 --mode train --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline
run 8 r8g_longclip_t32_b8_n8 --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ('r8g_infer_n8', 'r8g_train_b64_n8', 'r8g_longclip_t32_b8_n8'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'], 1), d['unit'], round(d['ms_per_step'], 2), (d.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
