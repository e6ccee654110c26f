#!/bin/bash
# r6f: full GPU test suite + smoke on the current tree, C2 / C3(N=1) / C4 benches, ncu of the temporal tail kernel
set -u
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/r6f_pytest_gpu.log 2>&1
tail -5 gpurun_out/r6f_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r6f_smoke.log 2>&1
tail -4 gpurun_out/r6f_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r6f_bench.json 2> gpurun_out/r6f_bench.err
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r6f_train_b64.json 2> gpurun_out/r6f_train.err
python bench.py --mode relevance --batch 32 --steps 5 --warmup 3 > gpurun_out/r6f_relevance_b32.json 2> gpurun_out/r6f_relevance.err
python - <<'PY'
import json
for f in ('r6f_bench', 'r6f_train_b64', 'r6f_relevance_b32'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['unit'], d['ms_per_step'], d.get('e2e', {}).get('value'), {k: round(v['ms_per_step'],2) for k, v in sorted(d.get('kernels', {}).items(), key=lambda kv: -kv[1]['ms_per_step'])[:9]})
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_temporal_mma_wide -s 12 -c 1 -o gpurun_out/r6f_attn_temporal_tail \
    python bench.py --frames 32 --batch 8 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6f_ncu.log 2>&1
tail -3 gpurun_out/r6f_ncu.log
