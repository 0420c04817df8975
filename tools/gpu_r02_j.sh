#!/bin/bash
# round 2, K2W (hub rows packed into a panel under a persisting L2 access-policy window) against K2
out=gpurun_out/r02_sweep_e_l2window.jsonl
: > $out
python - <<'PY' >> $out 2>gpurun_out/win_props.err
import json
from cuda.bindings import runtime as rt
err, p = rt.cudaGetDeviceProperties(0)
print(json.dumps(dict(probe="device", l2_bytes=p.l2CacheSize, persisting_l2_max=p.persistingL2CacheMaxSize, access_policy_max_window=p.accessPolicyMaxWindowSize)))
PY
timeout 300 python tools/kbench.py s24f32 --check --steps 5 --variants k2 win:mb=32 win:mb=64 win:mb=96 win:mb=16 >> $out 2>gpurun_out/win_s24.err || echo '{"w":"s24f32","error":"timeout or crash"}' >> $out
timeout 300 python tools/kbench.py c4 --check --steps 5 --variants k2 win:mb=64 win:mb=32 >> $out 2>gpurun_out/win_c4.err || echo '{"w":"c4","error":"timeout or crash"}' >> $out
timeout 200 python tools/kbench.py c2 --check --steps 10 --variants k2 win:mb=32 >> $out 2>gpurun_out/win_c2.err || echo '{"w":"c2","error":"timeout or crash"}' >> $out
cut -c1-260 $out
tail -n 3 gpurun_out/win_*.err
