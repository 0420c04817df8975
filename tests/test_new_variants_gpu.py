"""The persistent variants of the local multiply (csrc/cb_spmm_hub_kernel.cuh) on the GPU, through the C ABI: K2H (hub rows
resident in cluster shared memory) and K2R (gathers pipelined through a shared-memory ring with cp.async).

STATUS: validated on hardware in round 2 (44 / 44 cases, bit-identical to K2) and part of the regular GPU suite since.  Both
variants stay opt-in in the product (cb_spmm_hub_config / cb_spmm_ring_config): on every workload measured they are slower than
K2 (profiles/r02_sweep_a_k2_k2h_k2r.jsonl).  Also here: the narrow-panel layouts (CB_K2_NARROW=1) and the column filter for
sparse right-hand sides (cb_tile_filter_columns / CB_SPGEMM_FILTER=1).

The cases run in this process on one shared context (they ran in one process each, with a hard timeout, until the kernels had
been on hardware).  The bar is stronger than parity: K2H walks chunks exactly like K2, so its result must equal K2's BIT FOR BIT for every
semiring, floating point included, and equal the oracle within the usual tolerances."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

import numpy as np

import cbb200_loader
from oracle import oracle as O

cb = cbb200_loader.load_package()

CASES = {"pt_f32": (O.PLUS_TIMES, np.float32, np.float32, "value"), "pt_f64": (O.PLUS_TIMES, np.float64, np.float64, "value"),
         "minplus_i32": (O.MIN_PLUS, np.int32, np.int32, "x_minplus"), "pt_pat_i64": (O.PLUS_TIMES, None, np.int64, "value"),
         "selmax_i32": (O.MAX_SEL2ND, None, np.int32, "value"), "or_and": (O.OR_AND, None, np.uint8, "value"),
         "minplus_i64": (O.MIN_PLUS, np.int64, np.int64, "x_minplus")}


@pytest.fixture(scope="module")
def ctx():
    # tests/conftest.py sets CB_SPMM_HUB_MIN_COVER_PCT=0 before the library reads it: small test matrices never fall back for low coverage
    c = cb.Context(0)
    yield c
    c.hub_config(0)
    c.ring_config(0)
    c.close()


def run(ctx, case, cluster, slab, scale=12, k=64, ring=0):
    sr, adt, xdt, kind = CASES[case]
    n, I, J = O.rmat_matrix(scale, 16, seed=0)
    V = None if adt is None else O.matrix_values(I, J, n, 1, adt)
    X = O.dense_operand(n, k, 42, xdt, kind)
    t = ctx.tile_from_coo(n, n, I, J, V)
    Xd, Y0, Y1, Y2 = ctx.dense_from(X), ctx.dense(n, k, xdt), ctx.dense(n, k, xdt), ctx.dense(n, k, xdt)
    try:
        ctx.hub_config(0)
        ctx.ring_config(0)
        for acc in (False, True):                               # plain K2: overwrite, then accumulate on top
            ctx.spmm_local(t, Xd, Y0, sr, accumulate=acc)
        ctx.hub_config(1 if cluster > 0 else 0, max(cluster, 0), slab)
        ctx.ring_config(ring)
        for acc in (False, True):                               # K2H / K2R: the same two calls
            ctx.spmm_local(t, Xd, Y1, sr, accumulate=acc)
        ctx.spmm_local(t, Xd, Y2, sr)
        info = t.hub_info()
        a, a1, b = Y0.download(), Y1.download(), Y2.download()
    finally:
        ctx.hub_config(0)
        ctx.ring_config(0)
        for h in (t, Xd, Y0, Y1, Y2):
            h.free()
    ref = O.spmm(sr, n, n, I, J, V, X)
    assert cluster <= 0 or (info["built"] and info["resident"] > 0), f"the hub kernel did not run: {info}"
    assert a.tobytes() == a1.tobytes(), "the persistent variant differs from K2"
    if np.issubdtype(ref.dtype, np.floating):
        tol = 1e-5 if ref.dtype == np.float32 else 1e-12
        assert (np.abs(b - ref) <= tol * np.maximum(np.abs(ref), 1e-300)).all()
    else:
        assert np.array_equal(b, ref)


@pytest.mark.parametrize("case,k,stages,warps", [("pt_f32", 128, 2, 6), ("pt_f32", 128, 2, 3), ("pt_f32", 64, 4, 6), ("pt_f32", 64, 2, 12),
                                                 ("pt_f32", 32, 4, 12), ("minplus_i32", 32, 2, 5), ("minplus_i32", 64, 4, 6), ("pt_f32", 48, 2, 6)])
def test_bulk_copy_ring_matches_k2_bitwise(ctx, monkeypatch, case, k, stages, warps):
    """K2T (csrc/cb_spmm_tma_kernel.cuh): row gathers as cp.async.bulk copies into a per-warp shared-memory ring with mbarrier
    completion.  Same walk as K2, so the same bits, overwrite and accumulate, split rows included; a width the variant does not
    cover (k = 48: not a power-of-two row) must quietly run K2.  Opt-in (cb_spmm_k2_pipe(ctx, 16)): measured 1.5-2.4x slower than
    K2 (profiles/r02_sweep_d_tma.jsonl)."""
    monkeypatch.setenv("CB_TMA_STAGES", str(stages))
    monkeypatch.setenv("CB_TMA_WARPS", str(warps))
    sr, adt, xdt, kind = CASES[case]
    n, I, J = O.rmat_matrix(13, 16, seed=0)
    V = O.matrix_values(I, J, n, 1, adt)
    X = O.dense_operand(n, k, 42, xdt, kind)
    t = ctx.tile_from_coo(n, n, I, J, V)
    Xd, Y0, Y1 = ctx.dense_from(X), ctx.dense(n, k, xdt), ctx.dense(n, k, xdt)
    try:
        ctx.k2_pipe(0)
        for acc in (False, True):
            ctx.spmm_local(t, Xd, Y0, sr, accumulate=acc)
        ctx.k2_pipe(16)
        for acc in (False, True):
            ctx.spmm_local(t, Xd, Y1, sr, accumulate=acc)
        a, b = Y0.download(), Y1.download()
    finally:
        ctx.k2_pipe(-1)
        for h in (t, Xd, Y0, Y1):
            h.free()
    assert t is not None and a.tobytes() == b.tobytes(), "the bulk-copy ring differs from K2"


@pytest.mark.parametrize("case,k", [("pt_f32", 128), ("pt_f32", 64), ("pt_f32", 100), ("pt_f64", 128), ("pt_f64", 32), ("minplus_i32", 32),
                                    ("pt_f32", 300), ("selmax_i32", 16)])
def test_persistent_warps_match_k2_bitwise(ctx, case, k):
    """K2 with persistent warps (cb_spmm_k2_pipe(ctx, 64)): one launch fills the chip and every warp takes chunk groups from a
    counter.  Same walk, so the same bits - overwrite and accumulate, split rows included; panels of more than one column slab
    (k = 300) and operand types the variant is not built for (SelectMax) quietly run K2.  Opt-in; the one variant of round 2 that
    beats K2: C2 0.733 -> 0.661 ms, C5 2.25 -> 2.05 ms, equal on the DRAM-bound inputs (profiles/r02_sweep_h_persistent.jsonl)."""
    sr, adt, xdt, kind = CASES[case]
    n, I, J = O.rmat_matrix(13, 16, seed=0)
    V = None if adt is None else O.matrix_values(I, J, n, 1, adt)
    X = O.dense_operand(n, k, 42, xdt, kind)
    t = ctx.tile_from_coo(n, n, I, J, V)
    Xd, Y0, Y1 = ctx.dense_from(X), ctx.dense(n, k, xdt), ctx.dense(n, k, xdt)
    try:
        ctx.k2_pipe(0)
        for acc in (False, True):
            ctx.spmm_local(t, Xd, Y0, sr, accumulate=acc)
        ctx.k2_pipe(64)
        for acc in (False, True):
            ctx.spmm_local(t, Xd, Y1, sr, accumulate=acc)
        a, b = Y0.download(), Y1.download()
    finally:
        ctx.k2_pipe(-1)
        for h in (t, Xd, Y0, Y1):
            h.free()
    assert a.tobytes() == b.tobytes(), "persistent warps differ from K2"


@pytest.mark.parametrize("dtype,k,mb", [(np.float32, 128, 1), (np.float32, 64, 1), (np.float64, 128, 2), (np.float64, 64, 1), (np.float32, 128, 64),
                                        (np.float32, 96, 1)])
def test_hub_panel_under_l2_window_matches_k2_bitwise(ctx, dtype, k, mb):
    """K2W (cb_spmm_k2_pipe(ctx, 32) + cb_spmm_k2_l2(ctx, mb)): the rows of the most used columns are packed into a panel that a
    persisting L2 access-policy window keeps on the chip; the column stream carries bit 30 + rank for them.  Same walk, same
    bits - overwrite and accumulate, a budget smaller and larger than the columns worth keeping, two panels of different width on
    the same tile (the remapped stream is rebuilt), and a width the variant does not cover (k = 96) quietly on K2.  Opt-in:
    DRAM traffic -25 % on R-MAT s24 but no time gained (profiles/r02_ncu_l2window_s24f32.csv)."""
    n, I, J = O.rmat_matrix(13, 16, seed=0)
    V = O.matrix_values(I, J, n, 1, dtype)
    X = O.dense_operand(n, k, 42, dtype, "value")
    t = ctx.tile_from_coo(n, n, I, J, V)
    Xd, Y0, Y1 = ctx.dense_from(X), ctx.dense(n, k, dtype), ctx.dense(n, k, dtype)
    try:
        ctx.k2_pipe(0)
        for acc in (False, True):
            ctx.spmm_local(t, Xd, Y0, O.PLUS_TIMES, accumulate=acc)
        ctx.k2_l2(mb)
        ctx.k2_pipe(32)
        for acc in (False, True):
            ctx.spmm_local(t, Xd, Y1, O.PLUS_TIMES, accumulate=acc)
        a, b = Y0.download(), Y1.download()
        X2 = O.dense_operand(n, k // 2, 7, dtype, "value")           # another width on the same tile: another number of hub rows
        X2d, Z0, Z1 = ctx.dense_from(X2), ctx.dense(n, k // 2, dtype), ctx.dense(n, k // 2, dtype)
        ctx.spmm_local(t, X2d, Z1, O.PLUS_TIMES)
        ctx.k2_pipe(0)
        ctx.spmm_local(t, X2d, Z0, O.PLUS_TIMES)
        c, d = Z0.download(), Z1.download()
        for h in (X2d, Z0, Z1):
            h.free()
    finally:
        ctx.k2_pipe(-1)
        ctx.k2_l2(-1)
        for h in (t, Xd, Y0, Y1):
            h.free()
    assert a.tobytes() == b.tobytes() and c.tobytes() == d.tobytes(), "the hub-panel variant differs from K2"


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_hub_matches_k2_bitwise_fp32(ctx, cluster):
    run(ctx, "pt_f32", cluster, 0)


@pytest.mark.parametrize("case", ["pt_f64", "minplus_i32", "pt_pat_i64", "selmax_i32", "or_and"])
def test_hub_every_semiring(ctx, case):
    run(ctx, case, 4, 0, k=64 if case != "or_and" else 256)          # boolean panels: 256 one-byte columns = 256-byte rows


@pytest.mark.parametrize("slab,k", [(128, 64), (256, 100), (512, 300), (128, 33)])
def test_hub_column_slabs_and_ragged_widths(ctx, slab, k):
    run(ctx, "pt_f32", 2, slab, k=k)


def test_hub_larger_matrix_many_chunks(ctx):
    run(ctx, "minplus_i32", 4, 128, scale=16, k=32)


# ---- K2R: gathers pipelined through a shared-memory ring (cp.async), alone (cluster 0 = no hub rows) and with hub rows
@pytest.mark.parametrize("cluster", [0, 1, 4])
def test_ring_matches_k2_bitwise_fp32(ctx, cluster):
    run(ctx, "pt_f32", cluster, 0, ring=8)


@pytest.mark.parametrize("case", ["pt_f64", "minplus_i32", "pt_pat_i64", "selmax_i32", "or_and"])
def test_ring_every_semiring(ctx, case):
    run(ctx, case, 2, 0, k=64 if case != "or_and" else 256, ring=8)


@pytest.mark.parametrize("slab,k", [(128, 64), (256, 100), (512, 300)])
def test_ring_column_slabs_and_ragged_widths(ctx, slab, k):
    run(ctx, "pt_f32", 0, slab, k=k, ring=8)


@pytest.mark.parametrize("scale,k,what,cluster,ring", [(20, 64, "pt_f32", 4, 0), (20, 64, "pt_f32", 0, 8), (20, 64, "pt_f32", 4, 8),
                                                       (22, 32, "mp_i32", 8, 0), (22, 32, "mp_i32", 2, 8)])
def test_full_size_configs_bitwise_equal_to_k2(ctx, scale, k, what, cluster, ring):
    # BASELINE configs C2 (R-MAT scale 20 x 64 fp32) and C5 (scale 22 x 32 int32 MinPlus) generated on the device
    dt, sr, vd, kind = {"pt_f32": (np.float32, cb.PLUS_TIMES, cb.F32, 0), "mp_i32": (np.int32, cb.MIN_PLUS, cb.I32, 1)}[what]
    n = 1 << scale
    t = ctx.gen_rmat_tile(scale, 16, 0, val_dtype=vd, val_seed=1)
    X = ctx.dense(n, k, dt)
    X.generate(42, 0, 0, k, kind)
    Y0, Y1 = ctx.dense(n, k, dt), ctx.dense(n, k, dt)
    try:
        ctx.hub_config(0)
        ctx.ring_config(0)
        ctx.spmm_local(t, X, Y0, sr)
        ctx.hub_config(1 if cluster > 0 else 0, max(cluster, 0), 0)
        ctx.ring_config(ring)
        ctx.spmm_local(t, X, Y1, sr)
        assert Y0.download().tobytes() == Y1.download().tobytes(), "persistent variant differs from K2 at full size"
    finally:
        ctx.hub_config(0)
        ctx.ring_config(0)
        for h in (t, X, Y0, Y1):
            h.free()


# ---- narrow panels on 1- and 2-lane virtual warps (CB_K2_NARROW=1): K2's own walker with other template arguments.  The
# switch is read once per process, so these cases share one child process.
NARROW = r"""
import os, sys, numpy as np
sys.path.insert(0, %(root)r)
import cbb200_loader
from oracle import oracle as O
cb = cbb200_loader.load_package()
CASES = {"pt_f32": (O.PLUS_TIMES, np.float32, np.float32, "value"), "minplus_i64": (O.MIN_PLUS, np.int64, np.int64, "x_minplus"),
         "or_and": (O.OR_AND, None, np.uint8, "value"), "selmax_i32": (O.MAX_SEL2ND, None, np.int32, "value"),
         "pt_f64": (O.PLUS_TIMES, np.float64, np.float64, "value")}
n, I, J = O.rmat_matrix(13, 16, seed=0)
with cb.Context(0) as ctx:
    for spec in sys.argv[1:]:
        case, k = spec.split(":")
        k = int(k)
        sr, adt, xdt, kind = CASES[case]
        V = None if adt is None else O.matrix_values(I, J, n, 1, adt)
        X = O.dense_operand(n, k, 42, xdt, kind)
        t = ctx.tile_from_coo(n, n, I, J, V)
        Xd, Y = ctx.dense_from(X), ctx.dense(n, k, xdt)
        for acc in (False, True):
            ctx.spmm_local(t, Xd, Y, sr, accumulate=acc)
        got = Y.download()
        for h in (t, Xd, Y):
            h.free()
        ref = O.spmm(sr, n, n, I, J, V, X)
        ref2 = O.spmm(sr, n, n, I, J, V, X, accum_into=ref.copy())
        if np.issubdtype(ref2.dtype, np.floating):
            tol = 1e-5 if ref2.dtype == np.float32 else 1e-12
            assert (np.abs(got - ref2) <= tol * np.maximum(np.abs(ref2), 1e-300)).all(), spec
        else:
            assert np.array_equal(got, ref2), spec
        print("narrow ok", case, k, os.environ.get("CB_K2_NARROW"))
"""
NARROW_CASES = [("pt_f32", 1), ("pt_f32", 4), ("pt_f32", 8), ("pt_f32", 5), ("minplus_i64", 1), ("minplus_i64", 4), ("or_and", 16), ("or_and", 32),
                ("selmax_i32", 3), ("pt_f64", 2), ("pt_f64", 3)]


def test_narrow_layouts_match_the_oracle():
    r = subprocess.run([sys.executable, "-c", NARROW % {"root": ROOT}] + [f"{c}:{k}" for c, k in NARROW_CASES], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, CB_K2_NARROW="1"), cwd=ROOT)
    assert r.returncode == 0 and r.stdout.count("narrow ok") == len(NARROW_CASES), r.stdout[-2000:] + r.stderr[-3000:]


# ---- column filter for sparse right-hand sides (cb_tile_filter_columns; used by the dense-panel lowering CB_SPGEMM_DENSE=1 + CB_SPGEMM_FILTER=1)
def test_column_filter_keeps_exactly_the_marked_columns(ctx):
    n, I, J = O.rmat_matrix(12, 16, seed=0)
    V = O.matrix_values(I, J, n, 1, np.float64)
    rng = np.random.default_rng(1)
    keep = (rng.random(n) < 0.1).astype(np.uint8)
    t = ctx.tile_from_coo(n, n, I, J, V)
    f = t.filter_columns(keep)
    rowptr, col, vals = f.to_csr(np.float64)
    sel = keep[J] != 0
    order = np.lexsort((J[sel], I[sel]))
    assert f.nnz == int(sel.sum()) and f.m == n and f.n == n
    assert np.array_equal(col, J[sel][order]) and np.array_equal(vals, V[sel][order])
    assert np.array_equal(np.diff(rowptr), np.bincount(I[sel], minlength=n))
    X = O.dense_operand(n, 24, 42, np.float64)
    X[keep == 0] = 0.0                                       # rows a sparse B would not have
    Xd, Y0, Y1 = ctx.dense_from(X), ctx.dense(n, 24, np.float64), ctx.dense(n, 24, np.float64)
    ctx.spmm_local(t, Xd, Y0, cb.PLUS_TIMES)
    ctx.spmm_local(f, Xd, Y1, cb.PLUS_TIMES)
    a, b = Y0.download(), Y1.download()
    for h in (t, f, Xd, Y0, Y1):
        h.free()
    assert (np.abs(a - b) <= 1e-12 * np.maximum(np.abs(a), 1e-300)).all()


def test_driver_sparse_rhs_through_the_dense_panel_lowering_with_column_filter():
    from tests.test_host_cpp import DRIVER, build
    build()
    r = subprocess.run([DRIVER, "spgemm", "11", "40", "1"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, CB_SPGEMM_DENSE="1", CB_SPGEMM_FILTER="1"))
    assert r.returncode == 0 and "SpGEMM (sparse x sparse) working correctly" in r.stderr, r.stdout + r.stderr


def test_betwcent_application_through_the_dense_panel_lowering_with_column_filter(tmp_path):
    # the reference's unmodified BetwCent on this layer with the round-1 lowering (CB_SPGEMM_DENSE=1) and CB_SPGEMM_FILTER=1: same
    # scores as the reference wrote (the default path, the device sparse x sparse product, is tests/test_host_cpp.py's BetwCent case)
    exe = os.path.join(ROOT, "oracle", "_ref", "BetwCent_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/BetwCent_b200 was not built (needs the reference tree)")
    from tests.golden.make_golden_grid import BC_BATCH, BC_K4APPROX, betwcent_input
    betwcent_input(str(tmp_path))
    out = str(tmp_path / "bc.txt")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, str(tmp_path), str(BC_K4APPROX), str(BC_BATCH), out], capture_output=True, text=True, timeout=600,
                       env=dict(env, CB_SPGEMM_DENSE="1", CB_SPGEMM_FILTER="1"))
    assert r.returncode == 0 and "Computation finished" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "grid_ref.npz"))["betwcent_p1"]
    got = np.loadtxt(out, skiprows=1)[:, 2]
    assert np.abs(got - gold).max() <= 1e-9 * np.abs(gold).max()
