#!/usr/bin/env python
"""Kernel timing across workloads and K2 variants (device-resident operands, profile events on K2).  The tile and the
panels of a workload are built once and every variant runs on them in the same process.

usage: python tools/kbench.py c2 c5 ... [--steps 10] [--variants k2 k2:slab=128 hub:c=4,slab=128 ring hub:c=2,ring=8 ...]
variant = kind[:key=value,...];  kind k2 (round-1 kernel: slab = column-slab bytes, point = 0 deep / 1 wide),
pipe (K2P: d = ring depth 4 / 8, slab), hub (K2H: c = cluster size, slab), ring (K2R), hub with ring=8 (both)."""
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
ap.add_argument("--variants", nargs="*", default=["k2"])
ap.add_argument("--check", action="store_true", help="compare every variant's Y with the first variant's, bit for bit")
a = ap.parse_args()


def parse(v):
    kind, _, rest = v.partition(":")
    kv = dict(x.split("=") for x in rest.split(",") if x)
    return kind, {k: int(x) for k, x in kv.items()}


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
    ref_sum = None
    for v in a.variants:
        kind, o = parse(v)
        ctx.k2_config(0, -1); ctx.hub_config(0); ctx.ring_config(0); ctx.k2_pipe(0); ctx.k2_l2(0)
        try:
            if kind == "k2":
                ctx.k2_config(o.get("slab", 0), o.get("point", -1))
            elif kind == "pf":
                ctx.k2_config(o.get("slab", 0), o.get("point", -1)); ctx.k2_pipe(1)
            elif kind == "pipe":
                ctx.k2_config(o.get("slab", 0), -1); ctx.k2_pipe(o.get("d", 4)); ctx.k2_l2(o.get("l2", 0))
            elif kind == "tma":                      # bulk-copy ring (K2T): s = stages (2 / 4), w = warps per CTA
                os.environ["CB_TMA_STAGES"] = str(o.get("s", 0)); os.environ["CB_TMA_WARPS"] = str(o.get("w", 0))
                ctx.k2_pipe(16)
            elif kind == "persist":                  # K2 with persistent warps (chunk groups from a counter)
                ctx.k2_config(0, o.get("point", -1)); ctx.k2_pipe(64)
            elif kind == "win":                      # K2W: hub panel of mb megabytes under a persisting L2 window
                ctx.k2_l2(o.get("mb", 64)); ctx.k2_pipe(32)
            elif kind == "hub":
                ctx.hub_config(1, o.get("c", 4), o.get("slab", 0)); ctx.ring_config(o.get("ring", 0))
            elif kind == "ring":
                ctx.ring_config(8)
            else:
                raise SystemExit(f"unknown variant {v}")
            for _ in range(3): ctx.spmm_local(t, X, Y, sr)
            ctx.sync(); ctx.profile(True); ctx.timer_start()
            for _ in range(a.steps): ctx.spmm_local(t, X, Y, sr)
            ms = ctx.timer_stop() / a.steps
            pm, pn = ctx.profile_read(); ctx.profile(False)
        except cb.capi.CBError as e:
            print(json.dumps(dict(w=name, variant=v, error=str(e)[:200])), flush=True)
            continue
        k2 = pm["spmm"] / max(pn["spmm"], 1)
        b = alg_bytes(t.nnz, t.m, t.nzc, k, s_val, s_t)
        g = t.nnz * (4 + s_val) + t.nnz * k * s_t + t.m * k * s_t
        same = None
        if a.check:
            import zlib
            _, _, ld, _, ptr = Y.info()                      # every 257th row of Y: a strided view onto the same memory
            view = ctx.wrap(ptr, (N + 256) // 257, k, ld * 257, Y.code)
            h = zlib.crc32(view.download().tobytes()); view.free()
            if ref_sum is None: ref_sum = h
            same = h == ref_sum
        print(json.dumps(dict(w=name, variant=v, k=k, nnz=t.nnz, ms_step=round(ms, 4), k2_ms=round(k2, 4), fill_ms=round(pm["fill"] / max(pn["fill"], 1), 4),
                              fix_ms=round(pm["fixup"] / max(pn["fixup"], 1), 4), tflops=round(2 * t.nnz * k / ms / 1e9, 2),
                              alg_gbs=round(b / k2 / 1e6, 1), frac=round(b / k2 / 1e6 / 6542.1, 4), gather_tbs=round(g / k2 / 1e9, 2),
                              chunks=t.nchunks, split=t.nsplit, hub=t.hub_info(), same_as_first=same)), flush=True)
    ctx.k2_config(0, -1); ctx.hub_config(0); ctx.ring_config(0); ctx.k2_pipe(-1); ctx.k2_l2(-1)
    for h in (t, X, Y): h.free()
ctx.close()
