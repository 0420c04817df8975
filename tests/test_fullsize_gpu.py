"""Parity at BASELINE.json's full single-GPU sizes, where the CPU oracle cannot run the whole product in seconds:
size-independent properties plus exact oracle values on sampled rows.

  C2  R-MAT scale 20 (31.4 M nnz) x k=64  fp32 PlusTimes
  C5  R-MAT scale 22 (128 M nnz)  x k=32  int32 MinPlus, and the boolean OR-AND semiring on the pattern
  C3  Erdos-Renyi n=2^24 (268 M nnz) x k=128 fp32 PlusTimes - the workload bench.py's headline is quoted on
  C4  R-MAT scale 24 (521 M nnz)  x k=128 fp64 PlusTimes - the north star's target configuration
Checks: (1) sampled rows against a numpy evaluation of the semiring on the downloaded tile rows (exact for
integers, 1e-5 relative for fp32); (2) column independence - multiplying a column sub-panel gives the same bits;
(3) checksum of checksums - column sums of Y equal column-degree-weighted column sums of X; (4) linearity;
(5) cross-semiring consistency (OR-AND == [pattern PlusTimes > 0]); (6) rows without nonzeros hold SR::id().
"""
import numpy as np
import pytest

import cbb200_loader
from oracle import oracle as O

cb = cbb200_loader.load_package()
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cb.Context(0)
    yield c
    c.close()


def x_rows(rows, k, seed, dtype, kind="value"):
    """rows of the hash-generated panel without building the whole panel"""
    idx = (np.asarray(rows, np.uint64)[:, None] * np.uint64(k) + np.arange(k, dtype=np.uint64)[None, :])
    return O.hash_values(idx.ravel(), seed, dtype, kind).reshape(len(rows), k)


def sample_rows(rowptr, rng, count):
    deg = np.diff(rowptr)
    nonempty = np.flatnonzero(deg)
    picks = set(rng.choice(nonempty, count, replace=False).tolist())
    picks.update(np.argsort(deg)[-3:].tolist())                   # the three longest rows (split across many chunks)
    picks.update(np.flatnonzero(deg == 0)[:5].tolist())           # and a few empty ones
    return sorted(picks)


def test_c2_rmat20_k64_fp32_plus_times(ctx):
    scale, k = 20, 64
    n = 1 << scale
    t = ctx.gen_rmat_tile(scale, 16, 0, val_dtype=cb.F32, val_seed=1)
    X = ctx.dense(n, k, np.float32)
    X.generate(42, 0, 0, k, 0)
    Y = ctx.dense(n, k, np.float32)
    ctx.spmm_local(t, X, Y, cb.PLUS_TIMES)
    Yh = Y.download()
    rowptr, col, vals = t.to_csr(np.float32)
    assert t.nnz == 31400128 and t.nsplit > 1000
    rng = np.random.default_rng(0)
    for r in sample_rows(rowptr, rng, 300):
        c = col[rowptr[r]:rowptr[r + 1]]
        if len(c) == 0:
            assert (Yh[r] == 0).all()
            continue
        ref = (vals[rowptr[r]:rowptr[r + 1]].astype(np.float64)[:, None] * x_rows(c, k, 42, np.float32).astype(np.float64)).sum(axis=0)
        assert (np.abs(Yh[r] - ref) <= 1e-5 * np.abs(ref)).all(), r
    # generated values are what the numpy recipe says
    sel = rng.integers(0, t.nnz, 1000)
    rows_of = np.searchsorted(rowptr, sel, side="right") - 1
    assert np.array_equal(vals[sel], O.matrix_values(rows_of, col[sel], n, 1, np.float32))
    # column independence: a 16-column sub-panel reproduces the same bits
    Xs, Ys = ctx.dense_from(np.ascontiguousarray(X.download()[:, 16:32])), ctx.dense(n, 16, np.float32)
    ctx.spmm_local(t, Xs, Ys, cb.PLUS_TIMES)
    assert np.array_equal(Ys.download(), Yh[:, 16:32])
    # checksum of checksums in float64: 1^T Y == (1^T A) X
    colw = np.bincount(col, weights=vals.astype(np.float64), minlength=n)
    lhs = Yh.astype(np.float64).sum(axis=0)
    rhs = colw @ X.download().astype(np.float64)
    assert (np.abs(lhs - rhs) <= 1e-5 * np.abs(rhs)).all()
    # linearity: A(2X) == 2(AX) exactly (power-of-two scaling commutes with rounding)
    X2 = ctx.dense_from(2 * X.download())
    ctx.spmm_local(t, X2, Ys if False else Y, cb.PLUS_TIMES)
    assert np.array_equal(Y.download(), 2 * Yh)
    for h in (t, X, Y, Xs, Ys, X2):
        h.free()


def test_c5_rmat22_k32_int32_min_plus_and_boolean(ctx):
    scale, k = 22, 32
    n = 1 << scale
    t = ctx.gen_rmat_tile(scale, 16, 0, val_dtype=cb.I32, val_seed=1)
    X = ctx.dense(n, k, np.int32)
    X.generate(42, 0, 0, k, 1)                                    # ~1% entries at INT_MAX exercise inf_plus
    Y = ctx.dense(n, k, np.int32)
    ctx.spmm_local(t, X, Y, cb.MIN_PLUS)
    Yh = Y.download()
    rowptr, col, vals = t.to_csr(np.int32)
    assert t.nnz == 128305150
    INF = np.iinfo(np.int32).max
    rng = np.random.default_rng(1)
    rows = sample_rows(rowptr, rng, 300)
    for r in rows:
        c = col[rowptr[r]:rowptr[r + 1]]
        if len(c) == 0:
            assert (Yh[r] == INF).all()                            # SR::id() of MinPlus
            continue
        xr = x_rows(c, k, 42, np.int32, "x_minplus").astype(np.int64)
        a = vals[rowptr[r]:rowptr[r + 1]].astype(np.int64)[:, None]
        prod = np.where((xr == INF) | (a == INF), INF, a + xr)
        assert np.array_equal(Yh[r], prod.min(axis=0).astype(np.int32)), r
    # idempotence of min: folding the result in again changes nothing (accumulate mode on the same product)
    ctx.spmm_local(t, X, Y, cb.MIN_PLUS, accumulate=True)
    assert np.array_equal(Y.download(), Yh)
    t.free()
    # the pattern of the same graph under the boolean semiring and under integer PlusTimes
    tp = ctx.gen_rmat_tile(scale, 16, 0)
    Xb = ctx.dense(n, k, np.uint8)
    Xb.generate(7, 0, 0, k, 0)
    Yb = ctx.dense(n, k, np.uint8)
    ctx.spmm_local(tp, Xb, Yb, cb.OR_AND)
    Xi, Yi = ctx.dense_from(Xb.download().astype(np.int32)), ctx.dense(n, k, np.int32)
    ctx.spmm_local(tp, Xi, Yi, cb.PLUS_TIMES)
    Ybh, Yih = Yb.download(), Yi.download()
    assert np.array_equal(Ybh, (Yih > 0).astype(np.uint8))
    # checksum of checksums, exact in int64: column sums of Y == column-degree-weighted column sums of X
    deg_col = np.bincount(col, minlength=n).astype(np.int64)
    assert np.array_equal(Yih.astype(np.int64).sum(axis=0), deg_col @ Xi.download().astype(np.int64))
    for r in rows[:50]:
        c = col[rowptr[r]:rowptr[r + 1]]
        assert np.array_equal(Yih[r], x_rows(c, k, 7, np.uint8).astype(np.int32).sum(axis=0) if len(c) else np.zeros(k, np.int32))
    for h in (tp, X, Y, Xb, Yb, Xi, Yi):
        h.free()


def _large_plus_times(ctx, scale, initiator, symmetric, k, dt, tol, expect_nnz):
    """C3 / C4 in the style of the C2 test, without copying the whole tile or panel to the host where a sample will do:
    sampled rows (the longest, empty ones, random ones) against a float64 numpy evaluation, column independence bit for
    bit, checksum of checksums on a column block, linearity, identity in empty rows."""
    n = 1 << scale
    code = cb.capi.CODE_OF[np.dtype(dt)]
    t = ctx.gen_rmat_tile(scale, 16, 0, initiator, symmetric, val_dtype=code, val_seed=1)
    assert t.nnz == expect_nnz, t.nnz
    X = ctx.dense(n, k, dt)
    X.generate(42, 0, 0, k, 0)
    Y = ctx.dense(n, k, dt)
    ctx.spmm_local(t, X, Y, cb.PLUS_TIMES)
    lengths = t.row_lengths()
    assert lengths.sum() == t.nnz
    rng = np.random.default_rng(2)
    nonempty = np.flatnonzero(lengths)
    picks = set(rng.choice(nonempty, 200, replace=False).tolist())
    picks.update(np.argsort(lengths)[-3:].tolist())               # the three longest rows (split across many chunks)
    picks.update(np.flatnonzero(lengths == 0)[:5].tolist())       # rows without nonzeros (none in the ER matrix)
    rows = np.array(sorted(picks), np.int64)
    off, cols, vals = t.rows(rows, lengths, dt)
    assert np.array_equal(vals, O.matrix_values(np.repeat(rows, np.diff(off)), cols, n, 1, dt))     # generated values = the numpy recipe
    got = Y.download_rows(rows)
    worst = 0.0
    for i, r in enumerate(rows):
        c = cols[off[i]:off[i + 1]]
        if len(c) == 0:
            assert (got[i] == 0).all()
            continue
        assert (np.diff(c) > 0).all()                              # columns ascending, no duplicates
        ref = (vals[off[i]:off[i + 1]].astype(np.float64)[:, None] * x_rows(c, k, 42, dt).astype(np.float64)).sum(axis=0)
        rel = np.abs(got[i] - ref) / np.abs(ref)
        worst = max(worst, rel.max())
        assert (rel <= tol).all(), (r, len(c), rel.max())
    # the panel rows on the device are what the numpy recipe says
    some = rng.integers(0, n, 64)
    assert np.array_equal(X.download_rows(some), x_rows(some, k, 42, dt))
    # column independence: a 16-column sub-panel reproduces the same bits (strided views onto the same device memory)
    _, _, ld, _, xptr = X.info()
    _, _, _, _, yptr = Y.info()
    es = np.dtype(dt).itemsize
    Xs, Ys = ctx.dense(n, 16, dt), ctx.dense(n, 16, dt)
    xv = ctx.wrap(xptr + 16 * es, n, 16, ld, dt)
    Xs.upload(xv.download())
    ctx.spmm_local(t, Xs, Ys, cb.PLUS_TIMES)
    yv = ctx.wrap(yptr + 16 * es, n, 16, ld, dt)
    Yblock = yv.download()
    assert np.array_equal(Ys.download(), Yblock)
    # checksum of checksums on that column block, in float64: 1^T Y == (1^T A) X
    rowptr, col, v = t.to_csr(dt)
    assert rowptr[-1] == t.nnz and np.array_equal(np.diff(rowptr), lengths)
    colw = np.bincount(col, weights=v.astype(np.float64), minlength=n)
    del col, v
    lhs = Yblock.astype(np.float64).sum(axis=0)
    rhs = colw @ Xs.download().astype(np.float64)
    assert (np.abs(lhs - rhs) <= tol * np.abs(rhs)).all()
    # linearity: A(2X) == 2(AX) exactly (power-of-two scaling commutes with rounding)
    Xs.upload(2 * Xs.download())
    ctx.spmm_local(t, Xs, Ys, cb.PLUS_TIMES)
    assert np.array_equal(Ys.download(), 2 * Yblock)
    for h in (xv, yv, t, X, Y, Xs, Ys):
        h.free()
    return worst


def test_c3_er24_k128_fp32_plus_times(ctx):
    """BASELINE config C3 on one GPU - the workload every BENCH / SCALE number is quoted on."""
    worst = _large_plus_times(ctx, 24, (0.25, 0.25, 0.25, 0.25), False, 128, np.float32, 1e-5, 268435327)
    assert worst < 1e-5


def test_c4_rmat24_k128_fp64_plus_times(ctx):
    """BASELINE config C4 (the north star's target: R-MAT scale 24 x 128 columns, fp64) on one GPU."""
    worst = _large_plus_times(ctx, 24, (0.57, 0.19, 0.19, 0.05), True, 128, np.float64, 1e-12, 520752026)
    assert worst < 1e-12
