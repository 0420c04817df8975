#!/bin/bash
# round 2: K2W against K2 at the row widths of the multi-GPU grids (R-MAT s24 against 64 fp32 / 32 fp64 columns = 256-byte rows), one GPU
out=gpurun_out/r02_sweep_g_l2window_256byte_rows.jsonl
: > $out
timeout 300 python tools/kbench.py s24f32 --k 64 --check --steps 5 --variants k2 win:mb=64 win:mb=32 k2:point=1 >> $out 2>gpurun_out/win256.err
timeout 300 python tools/kbench.py c4 --k 32 --check --steps 5 --variants k2 win:mb=64 >> $out 2>>gpurun_out/win256.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep_g_l2window_256byte_rows.jsonl'):
    d=json.loads(l); print(d.get('w'), d.get('k'), d.get('variant'), d.get('k2_ms'), d.get('gather_tbs'), d.get('same_as_first'), d.get('error'))
PY
tail -n 3 gpurun_out/win256.err
