#!/bin/bash
# round 2: K2's operating points on DRAM-resident 128-byte rows (the 8-GPU shape of C3: k / pc = 32 fp32 columns), one GPU
out=gpurun_out/r02_sweep_f_128byte_rows.jsonl
: > $out
timeout 300 python tools/kbench.py c3 --k 32 --check --steps 5 --variants k2 k2:point=0 k2:point=1 pf tma tma:s=2 >> $out 2>gpurun_out/r128.err || echo '{"w":"c3k32","error":"timeout or crash"}' >> $out
timeout 300 python tools/kbench.py c3 --k 64 --check --steps 5 --variants k2 k2:point=0 k2:point=1 >> $out 2>>gpurun_out/r128.err || echo '{"w":"c3k64","error":"timeout or crash"}' >> $out
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep_f_128byte_rows.jsonl'):
    d=json.loads(l); print({k:d.get(k) for k in ('w','k','variant','k2_ms','gather_tbs','same_as_first','error')})
PY
