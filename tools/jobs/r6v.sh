#!/bin/bash
# r6v: verification of the tree — full GPU suite, smoke(), C2 / C3 / C4 / C5 benches on one GPU
set -u
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/r6v_pytest_gpu.log 2>&1
tail -4 gpurun_out/r6v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r6v_smoke.log 2>&1
tail -4 gpurun_out/r6v_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r6v_bench.json 2> gpurun_out/r6v_bench.err
python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6v_train_b64.json 2>> gpurun_out/r6v_bench.err
python bench.py --mode relevance --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r6v_relevance_b32.json 2>> gpurun_out/r6v_bench.err
python bench.py --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r6v_longclip_t32_b8.json 2>> gpurun_out/r6v_bench.err
python - <<'PY'
import json
for f in ('r6v_bench','r6v_train_b64','r6v_relevance_b32','r6v_longclip_t32_b8'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, d['clocks']['sm_mhz'], (d.get('roofline') or {}).get('frac'),
              {k: round(v['ms_per_step'],2) for k, v in list(d['kernels'].items())[:7]})
    except Exception as e:
        print(f, 'ERR', e)
PY
