#!/usr/bin/env python
"""Quick kernel timing across workloads (device-resident operands, profile events on K2).
usage: python tools/kbench.py c2 c5 ... [--steps 10]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cbb200_loader
from bench import WORKLOADS, INITIATOR, NPDT, alg_bytes
cb = cbb200_loader.load_package()

ap = argparse.ArgumentParser()
ap.add_argument("workloads", nargs="+")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--k", type=int, default=None)
a = ap.parse_args()
ctx = cb.Context(0)
for name in a.workloads:
    w = dict(WORKLOADS[name])
    if a.k: w["k"] = a.k
    sr = {"plus_times": cb.PLUS_TIMES, "min_plus": cb.MIN_PLUS, "or_and": cb.OR_AND, "select_max": cb.MAX_SEL2ND}[w["sr"]]
    xdt = NPDT[w["xdt"]]; s_t = np.dtype(xdt).itemsize
    adt = cb.PATTERN if w["adt"] is None else cb.capi.CODE_OF[np.dtype(NPDT[w["adt"]])]
    s_val = 0 if w["adt"] is None else np.dtype(NPDT[w["adt"]]).itemsize
    N = 1 << w["scale"]; k = w["k"]
    t = ctx.gen_rmat_tile(w["scale"], w["ef"], 0, INITIATOR[w["gen"]], w["sym"], val_dtype=adt, val_seed=1)
    X = ctx.dense(N, k, xdt); X.generate(42, 0, 0, k, w["kind"]); Y = ctx.dense(N, k, xdt)
    for _ in range(3): ctx.spmm_local(t, X, Y, sr)
    ctx.sync(); ctx.profile(True); ctx.timer_start()
    for _ in range(a.steps): ctx.spmm_local(t, X, Y, sr)
    ms = ctx.timer_stop() / a.steps
    pm, pn = ctx.profile_read(); ctx.profile(False)
    k2 = pm["spmm"] / max(pn["spmm"], 1)
    b = alg_bytes(t.nnz, t.m, t.nzc, k, s_val, s_t)
    g = t.nnz * (4 + s_val) + t.nnz * k * s_t + t.m * k * s_t
    print(json.dumps(dict(w=name, k=k, nnz=t.nnz, ms_step=round(ms, 4), k2_ms=round(k2, 4), fill_ms=round(pm["fill"] / max(pn["fill"], 1), 4),
                          fix_ms=round(pm["fixup"] / max(pn["fixup"], 1), 4), tflops=round(2 * t.nnz * k / ms / 1e9, 2),
                          alg_gbs=round(b / k2 / 1e6, 1), frac=round(b / k2 / 1e6 / 6542.1, 4), gather_tbs=round(g / k2 / 1e9, 2),
                          chunks=t.nchunks, split=t.nsplit, hub=t.hub_info())), flush=True)
    for h in (t, X, Y): h.free()
ctx.close()
