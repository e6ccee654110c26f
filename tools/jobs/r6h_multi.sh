#!/bin/bash
# r6h (8 GPUs of one box): the BASELINE configs that were never run at full width —
#   C3 training step, data-parallel, 64 clips per GPU, N = 4 and N = 8 (NCCL all-reduce of the flat gradient buffer)
#   C5 long clips (32 frames), 8 clips per GPU, clip-sharded over 8 GPUs
# plus the nn.DataParallel check on two devices (tests/model_checks.py::run_data_parallel_check)
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r6h_gpus.txt 2>&1
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python tools/gpu_check.py --only data_parallel --out gpurun_out/r6h_data_parallel.json --timeout 500 > gpurun_out/r6h_data_parallel.log 2>&1
tail -3 gpurun_out/r6h_data_parallel.log
run() {  # n, name, args...
  n=$1; name=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; tail -c 600 gpurun_out/$name.json | head -c 400; echo
}
run 8 r6h_train_b64_n8 --mode train --steps 5 --warmup 3
run 4 r6h_train_b64_n4 --mode train --steps 5 --warmup 3
run 8 r6h_longclip_t32_b8_n8 --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
run 8 r6h_infer_n8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ('r6h_train_b64_n8', 'r6h_train_b64_n4', 'r6h_longclip_t32_b8_n8', 'r6h_infer_n8'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], d.get('e2e', {}).get('value'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
