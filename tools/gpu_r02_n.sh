#!/bin/bash
# round 2: K2 with the row ids read two rows ahead, against the times of the same day (profiles/r02_k2_decomposition.jsonl: C2 0.731 / 0.764 ms)
out=gpurun_out/r02_k2_rowid_lookahead.jsonl
: > $out
timeout 200 python tools/kbench.py c2 --check --steps 20 --variants k2:point=1 k2:point=0 tma >> $out 2>gpurun_out/look.err
timeout 200 python tools/kbench.py c2 --k 128 --check --steps 20 --variants k2:point=1 k2:point=0 tma >> $out 2>>gpurun_out/look.err
timeout 200 python tools/kbench.py c5 --check --steps 10 --variants k2 tma >> $out 2>>gpurun_out/look.err
timeout 300 python tools/kbench.py c3 --steps 5 --variants k2 >> $out 2>>gpurun_out/look.err
timeout 300 python tools/kbench.py s24f32 --steps 5 --variants k2 win:mb=64 >> $out 2>>gpurun_out/look.err
timeout 300 python tools/kbench.py c4 --steps 5 --variants k2 >> $out 2>>gpurun_out/look.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_k2_rowid_lookahead.jsonl'):
    d=json.loads(l); print(d.get('w'), d.get('k'), d.get('variant'), d.get('k2_ms'), d.get('gather_tbs'), d.get('same_as_first'))
PY
tail -n 3 gpurun_out/look.err
