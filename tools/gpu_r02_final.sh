#!/bin/bash
# round 2, final state: the GPU suite, the smoke entry, and the driver's bench commands on one GPU
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/final_ref.err
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_final_reference_arm.json', 'gpurun_out/r02_final_bench_n1.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ('impl', 'value', 'unit', 'ms_per_step', 'gpu_launches')}, 'roofline', d.get('roofline'), 'parity', d.get('parity'), 'e2e', (d.get('e2e') or {}).get('value'),
              'target', {k: (v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), (v.get('parity') or {}).get('ok')) for k, v in (d.get('target') or {}).items()} if isinstance(d.get('target'), dict) else d.get('target'), 'clocks', d.get('clocks'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -n 3 gpurun_out/final_bench.err gpurun_out/final_ref.err
