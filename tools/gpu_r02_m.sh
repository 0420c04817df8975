#!/bin/bash
# round 2: what separates K2 from the arithmetic-free gather probe on L2-resident panels?  Timing-only builds of the library that
# take one piece of the walk away at a time (-DCB_EXP_NOVAL: no value stream; NOFP: XOR instead of multiply-add; NOFLUSH: rows
# are not stored; NOROWEND: no row ends at all; ALL: everything at once = the probe's loop inside K2's chunk walk).  Results of
# these builds are wrong by construction; only the times mean something.
out=gpurun_out/r02_k2_decomposition.jsonl
: > $out
for v in default noval nofp noflush norowend all; do
  lib=combblas-spmm-test_b200/lib_var/libk2_$v.so
  [ $v = default ] && lib=combblas-spmm-test_b200/lib/libcombblas_b200.so
  for wl in "c2" "c2 --k 128"; do
    CB_LIB=$PWD/$lib timeout 200 python tools/kbench.py $wl --steps 20 --variants k2:point=1 k2:point=0 2>>gpurun_out/decomp.err | sed "s/^{/{\"build\": \"$v\", /" >> $out
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_k2_decomposition.jsonl'):
    d=json.loads(l); print(d.get('build'), d.get('w'), d.get('k'), d.get('variant'), d.get('k2_ms'), d.get('gather_tbs'))
PY
tail -n 3 gpurun_out/decomp.err
