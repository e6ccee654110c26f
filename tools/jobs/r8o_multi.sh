#!/bin/bash
# r8o (2 GPUs): the torchrun path of bench.py after the sampled-events change — C2 and C3 at N = 2
set -u
mkdir -p gpurun_out
run() {
  n=$1; name=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
}
run 2 r8o_infer_n2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
run 2 r8o_train_b64_n2 --mode train --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ('r8o_infer_n2', 'r8o_train_b64_n2'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'], 1), d['unit'], round(d['ms_per_step'], 2), (d.get('e2e') or {}).get('value'), d['roofline'].get('events'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
