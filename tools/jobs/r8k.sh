#!/bin/bash
# r8k: the full GPU suite, smoke(), and every bench line on the final tree of round 2
set -u
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu > gpurun_out/r8k_pytest_gpu.log 2>&1
tail -4 gpurun_out/r8k_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r8k_bench.json 2> gpurun_out/r8k_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r8k_bench_reference.json 2> gpurun_out/r8k_bench_reference.err
timeout 300 python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8k_train_b64.json 2> gpurun_out/r8k_train.err
timeout 300 python bench.py --mode relevance --batch 32 --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8k_relevance_b32.json 2> gpurun_out/r8k_relevance.err
timeout 300 python bench.py --frames 32 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r8k_longclip_t32_b8.json 2> gpurun_out/r8k_longclip.err
python - <<'PY'
import json
for f in ('bench', 'bench_reference', 'train_b64', 'relevance_b32', 'longclip_t32_b8'):
    try:
        d = json.loads(open(f'gpurun_out/r8k_{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, round(d.get('value', 0), 2), d.get('unit'), 'ms', round(d.get('ms_per_step', 0), 2), 'e2e', round((d.get('e2e') or {}).get('value', 0), 1),
          'roofline', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get('roofline') or {}).items() if k in ('achieved', 'peak', 'frac', 'bound')},
          'cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'), 'clk', (d.get('clocks') or {}).get('sm_mhz'), 'launches', d.get('gpu_launches'))
PY
