"""Pins the CPU oracle (oracle/spmm_oracle.c) to the reference: golden vectors made by the
unmodified reference (tests/golden/make_golden.py), the known answers in the reference tree,
and - where oracle/_ref/libcbref.so exists - the live reference on fresh random inputs."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden.make_golden import COMBOS

G = os.path.join(os.path.dirname(__file__), "golden")


def test_torus_known_answer():
    # "The nnz values should be 112" - Applications/SpMMError.cpp:80; 96 twos + 16 fours
    g = np.load(os.path.join(G, "torus.npz"))
    assert len(g["CI"]) == 112 and (g["CV"] == 2).sum() == 96 and (g["CV"] == 4).sum() == 16
    ti, tj = g["ti"], g["tj"]
    Gd = np.zeros((16, 16), np.int64)
    Gd[ti, tj] = 1
    Y = O.spmm(O.PLUS_TIMES, 16, 16, ti, tj, np.ones(64, np.int64), Gd)
    ref = np.zeros((16, 16), np.int64)
    ref[g["CI"], g["CJ"]] = g["CV"]
    assert np.array_equal(Y, ref) and np.array_equal(Y, g["Ydense"])
    assert Y[0].tolist() == [4, 0, 2, 0, 0, 2, 0, 2, 2, 0, 0, 0, 0, 2, 0, 2]


def test_hepth_config_c1_bit_exact():
    g = np.load(os.path.join(G, "hepth.npz"))
    m, n = int(g["m"]), int(g["n"])
    assert (m, n, len(g["I"])) == (8361, 8361, 31502)          # Applications/hep-th-p4.txt:9
    X = O.dense_operand(n, 16, 42, np.float64)
    Y = O.spmm(O.PLUS_TIMES, m, n, g["I"], g["J"], g["V"], X)
    assert np.array_equal(Y, g["Y"])                            # same summation order => identical bits
    assert (Y != 0).sum() == 121760                             # 7610 non-isolated rows x 16
    # independent reference path (k x dense SpMV) agrees to rounding; oracle's own SpMV restatement too
    rel = np.abs(g["Yspmv2"] - Y[:, :2]) / np.maximum(np.abs(Y[:, :2]), 1e-300)
    assert rel.max() < 1e-12
    y0 = O.spmv_pt_f64(m, n, g["I"], g["J"], g["V"], X[:, 0])
    assert np.abs(y0 - Y[:, 0]).max() <= 1e-12 * np.abs(Y[:, 0]).max()


@pytest.mark.parametrize("name", ["seven", "nonsym", "large"])
def test_small_fixtures(name):
    g = np.load(os.path.join(G, "small.npz"))
    m, n = int(g[name + "_m"]), int(g[name + "_n"])
    X = O.dense_operand(n, 8, 42, np.float64)
    if name == "large":
        X = X - 0.5
    Y = O.spmm(O.PLUS_TIMES, m, n, g[name + "_I"], g[name + "_J"], g[name + "_V"], X)
    ref = g[name + "_Y"]
    scale = np.abs(ref).max()
    assert np.abs(Y - ref).max() <= 1e-13 * scale


@pytest.mark.parametrize("combo", COMBOS, ids=[c[0] for c in COMBOS])
def test_rmat10_all_semirings_bit_exact(combo):
    name, sr, adt, xdt, kind = combo
    g = np.load(os.path.join(G, "rmat10.npz"))
    n, I, J = int(g["n"]), g["I"].astype(np.int64), g["J"].astype(np.int64)
    n2, I2, J2 = O.rmat_matrix(10, 16, seed=0)
    assert n2 == n and np.array_equal(I, I2) and np.array_equal(J, J2)   # generator is frozen
    V = None if adt is None else O.matrix_values(I, J, n, 1, adt)
    X = O.dense_operand(n, 8, 42, xdt, kind)
    Y = O.spmm(sr, n, n, I, J, V, X)
    assert Y.dtype == g["Y_" + name].dtype
    assert np.array_equal(Y, g["Y_" + name])
    if sr == O.MIN_PLUS and np.issubdtype(np.dtype(xdt), np.integer):
        assert (X == np.iinfo(xdt).max).any()                  # inf_plus saturation is exercised


def _random_case(rng, m, n, nnz, adt, xdt, k, kind="value"):
    I = rng.integers(0, m, nnz)
    J = rng.integers(0, n, nnz)
    I, J, _ = O.dedup(I, J, None, n)
    V = None if adt is None else O.matrix_values(I, J, n, 7, adt)
    X = O.dense_operand(n, k, 9, xdt, kind)
    return I, J, V, X


@pytest.mark.parametrize("grid", [(1, 2), (2, 1), (2, 2), (2, 4), (3, 2), (4, 4)])
def test_summa_emulation_matches_single_rank(grid):
    pr, pc = grid
    rng = np.random.default_rng(3)
    m, n, k = 203, 157, 13                                     # ragged: not divisible by any grid dimension
    for sr, adt, xdt, kind in [(O.MIN_PLUS, np.int32, np.int32, "x_minplus"), (O.PLUS_TIMES, None, np.int64, "value"),
                               (O.MAX_SEL2ND, None, np.int32, "value"), (O.OR_AND, None, np.uint8, "value")]:
        I, J, V, X = _random_case(rng, m, n, 900, adt, xdt, k, kind)
        assert np.array_equal(O.spmm_summa(sr, pr, pc, m, n, I, J, V, X), O.spmm(sr, m, n, I, J, V, X))
    I, J, V, X = _random_case(rng, m, n, 900, np.float64, np.float64, k)
    a, b = O.spmm_summa(O.PLUS_TIMES, pr, pc, m, n, I, J, V, X), O.spmm(O.PLUS_TIMES, m, n, I, J, V, X)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()


def test_owner_rule():
    # SpParMat.cpp:5066-5096: floor division, last grid row/col takes the remainder
    assert O.owner(10, 10, 3, 3, 9, 9) == (8, 3, 3)
    assert O.owner(10, 10, 3, 3, 2, 3) == (1, 2, 0)
    assert O.owner(2, 2, 4, 4, 1, 0) == (15, 1, 0)              # per-proc size 0 -> last processor row and column
    assert O.block_range(10, 3, 2) == (6, 4) and O.block_range(10, 3, 0) == (0, 3)
    for total, nb in [(8361, 2), (8361, 4), (7, 8), (16, 4)]:
        cover = []
        for b in range(nb):
            s, l = O.block_range(total, nb, b)
            cover += list(range(s, s + l))
        assert cover == list(range(total))


def test_generators_are_deterministic_and_partition_independent():
    i1, j1 = O.rmat_edges(8, 4, seed=5)
    ia, ja = O.rmat_edges(8, 4, seed=5, first=0, count=300)
    ib, jb = O.rmat_edges(8, 4, seed=5, first=300)
    assert np.array_equal(i1, np.concatenate([ia, ib])) and np.array_equal(j1, np.concatenate([ja, jb]))
    perm = O.scramble(np.arange(1 << 9, dtype=np.uint64), 9, 5)
    assert np.array_equal(np.sort(perm), np.arange(1 << 9, dtype=np.uint64))
    n, I, J = O.rmat_matrix(9, 16, seed=0)
    assert (I != J).all()
    key = I * n + J
    assert len(np.unique(key)) == len(key)
    assert set(zip(I.tolist(), J.tolist())) == set(zip(J.tolist(), I.tolist()))      # symmetrised
    x = O.dense_operand(50, 4, 42, np.float32)
    assert x.dtype == np.float32 and (x > 0).all() and (x < 1).all()


def test_unsupported_combination_is_an_error():
    with pytest.raises(ValueError):
        O.spmm(O.MIN_PLUS, 2, 2, [0], [1], np.ones(1, np.float32), np.ones((2, 2), np.float64))


# ------------------------------------------------------------------ live checks against the compiled reference
@pytest.mark.ref
@pytest.mark.parametrize("combo", COMBOS, ids=[c[0] for c in COMBOS])
def test_live_reference_random(combo):
    name, sr, adt, xdt, kind = combo
    rng = np.random.default_rng(11)
    for (m, n, nnz, k) in [(64, 48, 400, 5), (300, 300, 3000, 16), (17, 90, 40, 1)]:
        I, J, V, X = _random_case(rng, m, n, nnz, adt, xdt, k, kind)
        Yr, _ = O.ref_spmm(sr, m, n, I, J, V, X)
        Yo = O.spmm(sr, m, n, I, J, V, X)
        if np.issubdtype(Yr.dtype, np.floating):
            # very sparse columns take the reference's heap branch (mtSpGEMM.h:311-360) whose order among
            # equal rows is heap order; everything else is the ascending-kk hash order => identical bits
            tol = 1e-5 if Yr.dtype == np.float32 else 1e-12
            assert np.abs(Yr - Yo).max() <= tol * max(np.abs(Yr).max(), 1)
        else:
            assert np.array_equal(Yr, Yo)


@pytest.mark.ref
def test_live_reference_empty_and_degenerate():
    X = O.dense_operand(6, 3, 1, np.int32)
    Yr, _ = O.ref_spmm(O.MIN_PLUS, 5, 6, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int32), X)
    Yo = O.spmm(O.MIN_PLUS, 5, 6, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int32), X)
    assert np.array_equal(Yr, Yo) and (Yo == np.iinfo(np.int32).max).all()
    # one dense row, everything else empty
    J = np.arange(6)
    I = np.full(6, 2)
    Yr, _ = O.ref_spmm(O.MAX_SEL2ND, 5, 6, I, J, None, X)
    Yo = O.spmm(O.MAX_SEL2ND, 5, 6, I, J, None, X)
    assert np.array_equal(Yr, Yo) and (Yo[0] == -1).all() and np.array_equal(Yo[2], X.max(axis=0))


@pytest.mark.ref
def test_live_reference_spmv_path_agrees():
    rng = np.random.default_rng(5)
    I, J, V, X = _random_case(rng, 200, 180, 2500, np.int32, np.int32, 4, "x_minplus")
    a, _ = O.ref_spmm(O.MIN_PLUS, 200, 180, I, J, V, X, via=0)
    b, _ = O.ref_spmm(O.MIN_PLUS, 200, 180, I, J, V, X, via=1)
    assert np.array_equal(a, b) and np.array_equal(a, O.spmm(O.MIN_PLUS, 200, 180, I, J, V, X))


def test_c_restatement_of_the_generators_matches_numpy():
    """oracle/gen_oracle.c (what bench.py's CPU legs build full-size operands with) against the numpy recipes, bit for bit:
    matrices (symmetrised R-MAT, directed ER; row-major and column-major order), matrix values, panel columns."""
    for sc, sym, init in [(11, True, (0.57, 0.19, 0.19, 0.05)), (12, False, (0.25, 0.25, 0.25, 0.25))]:
        n, I, J = O.rmat_matrix(sc, 16, 3, init, symmetric=sym)
        n2, I2, J2 = O.rmat_matrix_fast(sc, 16, 3, init, symmetric=sym)
        assert n == n2 and np.array_equal(I, I2) and np.array_equal(J, J2)
        _, I3, J3 = O.rmat_matrix_fast(sc, 16, 3, init, symmetric=sym, col_major=True)
        o = np.lexsort((I, J))
        assert np.array_equal(I[o], I3) and np.array_equal(J[o], J3)
        for dt in (np.float32, np.float64, np.int32, np.int64):
            assert np.array_equal(O.matrix_values(I, J, n, 1, dt), O.matrix_values_fast(I, J, n, 1, dt))
        for dt, kind in ((np.float32, "value"), (np.float64, "value"), (np.int32, "x_minplus"), (np.uint8, "value")):
            assert np.array_equal(O.dense_operand(n, 24, 42, dt, kind)[:, 8:16], O.dense_columns_fast(n, 24, 8, 8, 42, dt, kind))


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref (compiled reference) not present")
def test_resident_reference_matrix_equals_the_one_shot_call():
    """cbref_matrix_* (one SpParMat, many Mult_AnXBn_Synch calls: bench.py's CPU legs) gives what cbref_spmm gives."""
    n, I, J = O.rmat_matrix_fast(10, 8, 0, col_major=True)
    V = O.matrix_values_fast(I, J, n, 1, np.float64)
    X = O.dense_operand(n, 5, 42, np.float64)
    A = O.RefMatrix(O.PLUS_TIMES, n, n, I, J, V, np.float64, colmajor_sorted=True)
    Y, sec, nnzc = A.mult(X)
    A.free()
    Y2, _ = O.ref_spmm(O.PLUS_TIMES, n, n, I, J, V, X)
    assert np.array_equal(Y, Y2) and sec > 0 and nnzc == (Y2 != 0).sum()
