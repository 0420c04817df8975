"""The device generators (csrc/cb_gen.cu) reproduce the numpy recipe in oracle/oracle.py bit for bit, so the
oracle and the GPU are always fed the same operands, for any block of any grid."""
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


@pytest.mark.parametrize("sym,init", [(True, (0.57, 0.19, 0.19, 0.05)), (False, (0.25, 0.25, 0.25, 0.25))])
def test_rmat_tile_equals_numpy_recipe(ctx, sym, init):
    scale = 11
    n, I, J = O.rmat_matrix(scale, 16, 0, init, symmetric=sym)
    V = O.matrix_values(I, J, n, 1, np.float32)
    t = ctx.gen_rmat_tile(scale, 16, 0, init, sym, val_dtype=cb.F32, val_seed=1)
    assert t.nnz == len(I)
    rowptr, col, vals = t.to_csr(np.float32)
    order = np.lexsort((J, I))
    assert np.array_equal(col, J[order]) and np.array_equal(np.diff(rowptr), np.bincount(I, minlength=n))
    assert np.array_equal(vals, V[order])
    t.free()


def test_blocks_of_a_grid_tile_the_matrix(ctx):
    scale = 10
    n, I, J = O.rmat_matrix(scale, 16, 0)
    V = O.matrix_values(I, J, n, 1, np.int32)
    total = 0
    for pr, pc in [(2, 2), (2, 4), (3, 1)]:
        total = 0
        for r in range(pr):
            r0, rl = O.block_range(n, pr, r)
            for c in range(pc):
                c0, cl = O.block_range(n, pc, c)
                t = ctx.gen_rmat_tile(scale, 16, 0, row0=r0, m=rl, col0=c0, n=cl, val_dtype=cb.I32, val_seed=1)
                sel = (I >= r0) & (I < r0 + rl) & (J >= c0) & (J < c0 + cl)
                rowptr, col, vals = t.to_csr(np.int32)
                order = np.lexsort((J[sel], I[sel]))
                assert np.array_equal(col, (J[sel] - c0)[order]) and np.array_equal(vals, V[sel][order])
                total += t.nnz
                t.free()
        assert total == len(I)


@pytest.mark.parametrize("dt,kind", [(np.float32, 0), (np.float64, 0), (np.int32, 1), (np.int64, 1), (np.uint8, 0)])
def test_dense_panel_equals_numpy_recipe(ctx, dt, kind):
    rows, k = 300, 24
    d = ctx.dense(rows, 10, dt)
    d.generate(42, row0=17, col0=5, gk=k, kind=kind)
    full = O.dense_operand(400, k, 42, dt, "x_minplus" if kind else "value")
    assert np.array_equal(d.download(), full[17:317, 5:15])
    d.free()
