#!/bin/bash
# r7f: verification of the round's final tree — full GPU suite, smoke(), default bench line (with the reference CPU leg and
# the eager-GPU baseline), the reference arm, C3 / C4 / C5 on one GPU
set -u
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/r7f_pytest_gpu.log 2>&1
tail -4 gpurun_out/r7f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r7f_smoke.log 2>&1
tail -4 gpurun_out/r7f_smoke.log
python bench.py > gpurun_out/r7f_bench.json 2> gpurun_out/r7f_bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r7f_reference_arm.json 2>> gpurun_out/r7f_bench.err
python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7f_train_b64.json 2>> gpurun_out/r7f_bench.err
python bench.py --mode relevance --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r7f_relevance_b32.json 2>> gpurun_out/r7f_bench.err
python bench.py --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r7f_longclip_t32_b8.json 2>> gpurun_out/r7f_bench.err
python - <<'PY'
import json
for f in ('r7f_bench','r7f_reference_arm','r7f_train_b64','r7f_relevance_b32','r7f_longclip_t32_b8'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round((d.get('e2e') or {}).get('value',0),1), (d.get('clocks') or {}).get('sm_mhz'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('kind'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r7f_bench.err
