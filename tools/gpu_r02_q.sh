#!/bin/bash
# round 2: K2 with persistent warps against K2 (bit-compared), one GPU
out=gpurun_out/r02_sweep_h_persistent.jsonl
: > $out
timeout 100 python tools/kbench.py c2 --check --steps 20 --variants k2 persist persist:point=0 >> $out 2>gpurun_out/persist.err
timeout 100 python tools/kbench.py c5 --check --steps 10 --variants k2 persist >> $out 2>>gpurun_out/persist.err
timeout 120 python tools/kbench.py s24f32 --check --steps 5 --variants k2 persist >> $out 2>>gpurun_out/persist.err
timeout 120 python tools/kbench.py c3 --check --steps 5 --variants k2 persist >> $out 2>>gpurun_out/persist.err
timeout 120 python tools/kbench.py c4 --check --steps 3 --variants k2 persist >> $out 2>>gpurun_out/persist.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep_h_persistent.jsonl'):
    d=json.loads(l); print(d.get('w'), d.get('k'), d.get('variant'), d.get('k2_ms'), d.get('gather_tbs'), d.get('same_as_first'), d.get('error'))
PY
tail -n 3 gpurun_out/persist.err
