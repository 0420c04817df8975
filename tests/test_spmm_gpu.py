"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors made
by the unmodified reference.  Bar: bit-exact for integer / boolean semirings; <= 1e-12 relative
(fp64) and <= 1e-5 relative (fp32) for floating point (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import cbb200_loader
from oracle import oracle as O
from tests.golden.make_golden import COMBOS

cb = cbb200_loader.load_package()
pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


@pytest.fixture(scope="module")
def ctx():
    c = cb.Context(0)
    yield c
    c.close()


def check(Y, ref):
    assert Y.dtype == ref.dtype and Y.shape == ref.shape
    if np.issubdtype(ref.dtype, np.floating):
        # relative to the magnitude of the row sum of |terms| is overkill here: operands are positive
        err = np.abs(Y.astype(np.float64) - ref.astype(np.float64))
        assert (err <= TOL[ref.dtype] * np.maximum(np.abs(ref.astype(np.float64)), np.finfo(ref.dtype).tiny)).all(), err.max()
    else:
        assert np.array_equal(Y, ref)


def gpu_spmm(ctx, m, n, I, J, V, X, sr, how="coo"):
    if how == "coo":
        t = ctx.tile_from_coo(m, n, np.asarray(I, np.int64), np.asarray(J, np.int64), V)
    elif how == "dcsc":
        cp, jc, ir, numx = O.to_dcsc(m, n, I, J, V)
        t = ctx.tile_from_csc(m, n, cp, ir, numx, jc=jc)
    else:
        cp, ir, numx = O.to_csc(m, n, I, J, V)
        t = ctx.tile_from_csc(m, n, cp.astype(np.int32), ir.astype(np.int32), numx)
    Xd = ctx.dense_from(X)
    Yd = ctx.dense(m, X.shape[1], X.dtype if X.dtype != np.bool_ else np.uint8)
    ctx.spmm_local(t, Xd, Yd, sr)
    Y = Yd.download()
    for h in (t, Xd, Yd):
        h.free()
    return Y


def test_torus_known_answer(ctx):
    g = np.load(os.path.join(G, "torus.npz"))
    Gd = np.zeros((16, 16), np.int64)
    Gd[g["ti"], g["tj"]] = 1
    Y = gpu_spmm(ctx, 16, 16, g["ti"], g["tj"], np.ones(64, np.int64), Gd, cb.PLUS_TIMES)
    assert np.array_equal(Y, g["Ydense"]) and (Y != 0).sum() == 112


@pytest.mark.parametrize("case", ["minplus_i32", "pt_f64", "pt_f32", "pt_pat_i64", "selmax_i32", "or_and"])
def test_matches_reference_run_on_process_grids(ctx, case):
    # tests/golden/grid_ref.npz: the unmodified reference's own 2x2- and 3x3-process multiplies
    from tests.golden.make_golden_grid import operands
    g = np.load(os.path.join(G, "grid_ref.npz"))
    sr, m, n, I, J, V, X = operands(case)
    Y = gpu_spmm(ctx, m, n, I, J, V, X, sr)
    for p in (4, 9):
        check(Y, g[f"{case}_p{p}"])


@pytest.mark.parametrize("how", ["coo", "dcsc", "csc"])
def test_hepth_config_c1(ctx, how):
    g = np.load(os.path.join(G, "hepth.npz"))
    m, n = int(g["m"]), int(g["n"])
    X = O.dense_operand(n, 16, 42, np.float64)
    Y = gpu_spmm(ctx, m, n, g["I"], g["J"], g["V"], X, cb.PLUS_TIMES, how)
    check(Y, g["Y"])
    # rows that are not cut by the work partition follow the reference's summation order exactly
    assert (Y == g["Y"]).mean() > 0.95
    # ... and the reference's own 2x2-process run of the same configuration (ParallelReadMM on four ranks, grid_ref.npz)
    check(Y, np.load(os.path.join(G, "grid_ref.npz"))["hepth_p4"])


@pytest.mark.parametrize("name", ["seven", "nonsym", "large"])
def test_small_fixtures(ctx, name):
    g = np.load(os.path.join(G, "small.npz"))
    m, n = int(g[name + "_m"]), int(g[name + "_n"])
    X = O.dense_operand(n, 8, 42, np.float64)
    if name == "large":
        X = X - 0.5
    Y = gpu_spmm(ctx, m, n, g[name + "_I"], g[name + "_J"], g[name + "_V"], X, cb.PLUS_TIMES)
    ref = g[name + "_Y"]
    assert np.abs(Y - ref).max() <= 1e-12 * np.abs(ref).max()       # mixed signs: scale by the largest entry


@pytest.mark.parametrize("combo", COMBOS, ids=[c[0] for c in COMBOS])
def test_rmat10_every_semiring_vs_reference_golden(ctx, combo):
    name, sr, adt, xdt, kind = combo
    g = np.load(os.path.join(G, "rmat10.npz"))
    n, I, J = int(g["n"]), g["I"].astype(np.int64), g["J"].astype(np.int64)
    V = None if adt is None else O.matrix_values(I, J, n, 1, adt)
    X = O.dense_operand(n, 8, 42, xdt, kind)
    check(gpu_spmm(ctx, n, n, I, J, V, X, sr), g["Y_" + name])


def random_case(rng, m, n, nnz, adt, xdt, k, kind="value"):
    I = rng.integers(0, m, nnz)
    J = rng.integers(0, n, nnz)
    I, J, _ = O.dedup(I, J, None, n)
    V = None if adt is None else O.matrix_values(I, J, n, 7, adt)
    X = O.dense_operand(n, k, 9, xdt, kind)
    return I, J, V, X


@pytest.mark.parametrize("k", [1, 3, 4, 13, 16, 32, 33, 64, 100, 128, 200, 300])
@pytest.mark.parametrize("combo", [COMBOS[0], COMBOS[1], COMBOS[4], COMBOS[8], COMBOS[9], COMBOS[10], COMBOS[14]],
                         ids=lambda c: c[0])
def test_panel_widths(ctx, k, combo):
    name, sr, adt, xdt, kind = combo
    rng = np.random.default_rng(k)
    m, n = 777, 501
    I, J, V, X = random_case(rng, m, n, 9000, adt, xdt, k, kind)
    check(gpu_spmm(ctx, m, n, I, J, V, X, sr), O.spmm(sr, m, n, I, J, V, X))


def test_empty_and_degenerate(ctx):
    X = O.dense_operand(6, 5, 1, np.int32)
    z = np.zeros(0, np.int64)
    Y = gpu_spmm(ctx, 5, 6, z, z, np.zeros(0, np.int32), X, cb.MIN_PLUS)
    assert (Y == np.iinfo(np.int32).max).all()                      # no nonzeros: every row is SR::id()
    J = np.arange(6)
    I = np.full(6, 2)
    Y = gpu_spmm(ctx, 5, 6, I, J, None, X, cb.MAX_SEL2ND)
    assert (Y[[0, 1, 3, 4]] == -1).all() and np.array_equal(Y[2], X.max(axis=0))
    # single column / single row / 1x1
    Y = gpu_spmm(ctx, 1, 1, [0], [0], np.array([2.5]), np.array([[4.0]]), cb.PLUS_TIMES)
    assert Y[0, 0] == 10.0
    # SelectMax keeps values below its identity when a row has entries (first product is stored, mtSpGEMM.h:410-414)
    Xn = np.full((3, 2), -7, np.int64)
    Y = gpu_spmm(ctx, 2, 3, [0], [1], None, Xn, cb.MAX_SEL2ND)
    assert np.array_equal(Y, O.spmm(O.MAX_SEL2ND, 2, 3, [0], [1], None, Xn)) and (Y[0] == -7).all() and (Y[1] == -1).all()


def test_hub_rows_are_split_and_recombined(ctx):
    # rows far longer than the chunk length exercise the carry buffer + fix-up kernel
    rng = np.random.default_rng(0)
    m, n, k = 300, 40000, 64
    hub = np.concatenate([rng.choice(n, 30000, replace=False), rng.choice(n, 5000, replace=False), rng.choice(n, 700, replace=False)])
    I = np.concatenate([np.full(30000, 7), np.full(5000, 8), np.full(700, 299), rng.integers(0, m, 4000)])
    J = np.concatenate([hub, rng.integers(0, n, 4000)])
    I, J, _ = O.dedup(I, J, None, n)
    for sr, adt, xdt, kind in [(cb.PLUS_TIMES, np.float32, np.float32, "value"), (cb.MIN_PLUS, np.int32, np.int32, "x_minplus"),
                               (cb.PLUS_TIMES, None, np.int64, "value"), (cb.PLUS_TIMES, np.float64, np.float64, "value")]:
        V = None if adt is None else O.matrix_values(I, J, n, 3, adt)
        X = O.dense_operand(n, k, 5, xdt, kind)
        t = ctx.tile_from_coo(m, n, I, J, V)
        assert t.nsplit >= 3
        t.free()
        check(gpu_spmm(ctx, m, n, I, J, V, X, sr), O.spmm(sr, m, n, I, J, V, X))


def test_accumulate_is_the_stage_merge(ctx):
    # two column blocks of A applied one after the other = one SUMMA rank's stage loop (ParFriends.h:1036-1083)
    rng = np.random.default_rng(4)
    m, n, k = 400, 600, 32
    for sr, adt, xdt, kind in [(cb.MIN_PLUS, np.int32, np.int32, "x_minplus"), (cb.MAX_SEL2ND, None, np.int32, "value"),
                               (cb.PLUS_TIMES, np.float64, np.float64, "value"), (cb.OR_AND, None, np.uint8, "value")]:
        I, J, V, X = random_case(rng, m, n, 6000, adt, xdt, k, kind)
        left = J < 250
        t1 = ctx.tile_from_coo(m, 250, I[left], J[left], None if V is None else V[left])
        t2 = ctx.tile_from_coo(m, 350, I[~left], J[~left] - 250, None if V is None else V[~left])
        X1, X2 = ctx.dense_from(X[:250]), ctx.dense_from(X[250:])
        Y = ctx.dense(m, k, X.dtype)
        ctx.spmm_local(t1, X1, Y, sr, accumulate=False)
        ctx.spmm_local(t2, X2, Y, sr, accumulate=True)
        check(Y.download(), O.spmm(sr, m, n, I, J, V, X))
        for h in (t1, t2, X1, X2, Y):
            h.free()


def test_tile_structure_round_trip(ctx):
    rng = np.random.default_rng(8)
    m, n = 1000, 900
    I, J, V, _ = random_case(rng, m, n, 20000, np.float64, np.float64, 1)
    keep = (I % 7 != 0) & (J % 5 != 0)                              # empty rows and columns
    I, J, V = I[keep], J[keep], V[keep]
    for how in ("coo", "dcsc", "csc"):
        if how == "coo":
            p = rng.permutation(len(I))
            t = ctx.tile_from_coo(m, n, I[p], J[p], V[p])
        elif how == "dcsc":
            cp, jc, ir, numx = O.to_dcsc(m, n, I, J, V)
            t = ctx.tile_from_csc(m, n, cp, ir, numx, jc=jc)
        else:
            cp, ir, numx = O.to_csc(m, n, I, J, V)
            t = ctx.tile_from_csc(m, n, cp, ir, numx)
        assert (t.nnz, t.m, t.n) == (len(I), m, n)
        assert t.nzr == len(np.unique(I)) and t.nzc == len(np.unique(J))
        rowptr, col, vals = t.to_csr(np.float64)
        order = np.lexsort((J, I))
        assert np.array_equal(col, J[order]) and np.array_equal(vals, V[order])
        assert np.array_equal(np.diff(rowptr), np.bincount(I, minlength=m))
        t.free()


@pytest.mark.parametrize("k,adt,xdt,sr,kind", [(48, np.float32, np.float32, 0, "value"), (100, np.float32, np.float32, 0, "value"),
                                               (128, np.float32, np.float32, 0, "value"), (64, np.int64, np.int64, 1, "x_minplus"),
                                               (7, None, np.int32, 2, "value"), (256, None, np.uint8, 3, "value")])
def test_host_panel_entry_point(ctx, k, adt, xdt, sr, kind):
    # host panels in / out; wide panels are pipelined as column slabs over three streams
    rng = np.random.default_rng(2)
    I, J, V, X = random_case(rng, 500, 400, 5000, adt, xdt, k, kind)
    t = ctx.tile_from_coo(500, 400, I, J, V)
    for _ in range(2):
        check(ctx.spmm_host(t, X, sr), O.spmm(sr, 500, 400, I, J, V, X))
    t.free()


def test_wrapping_foreign_device_memory(ctx):
    # panels that live in someone else's allocation (here: torch tensors) are used in place through cb_dense_wrap
    import torch
    rng = np.random.default_rng(6)
    I, J, V, X = random_case(rng, 300, 260, 4000, np.float64, np.float64, 32)
    t = ctx.tile_from_coo(300, 260, I, J, V)
    Xt = torch.from_numpy(X).cuda()
    Yt = torch.empty((300, 32), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    Xw, Yw = ctx.wrap(Xt.data_ptr(), 260, 32, 32, np.float64), ctx.wrap(Yt.data_ptr(), 300, 32, 32, np.float64)
    ctx.spmm_local(t, Xw, Yw, cb.PLUS_TIMES)
    ctx.sync()
    check(Yt.cpu().numpy(), O.spmm(O.PLUS_TIMES, 300, 260, I, J, V, X))
    with pytest.raises(cb.CBError) as e:                            # rows must start on 16-byte boundaries
        ctx.spmm_local(t, ctx.wrap(Xt.data_ptr() + 8, 260, 31, 32, np.float64), Yw, cb.PLUS_TIMES)
    assert e.value.status in (3002, 3007)
    for h in (t, Xw, Yw):
        h.free()


def test_error_codes_follow_the_reference(ctx):
    t = ctx.tile_from_coo(4, 5, [0], [1], np.ones(1, np.float32))
    X, Xbad, Y = ctx.dense(5, 3, np.float32), ctx.dense(6, 3, np.float32), ctx.dense(4, 3, np.float32)
    with pytest.raises(cb.CBError) as e:
        ctx.spmm_local(t, Xbad, Y, cb.PLUS_TIMES)
    assert e.value.status == 3002                                    # DIMMISMATCH, SpDefs.h:73
    with pytest.raises(cb.CBError) as e:
        ctx.spmm_local(t, X, ctx.dense(4, 3, np.float64), cb.PLUS_TIMES)
    assert e.value.status == 4
    with pytest.raises(cb.CBError) as e:
        ctx.spmm_local(t, X, Y, cb.MAX_SEL2ND)                       # SelectMax needs a boolean matrix
    assert e.value.status == 4
    with pytest.raises(cb.CBError) as e:
        ctx.tile_from_coo(2**31 + 5, 5, [0], [1], None)
    assert e.value.status == 6
