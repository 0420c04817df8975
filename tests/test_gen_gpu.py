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


def test_graph500_stream_is_the_references_own(ctx):
    """cb_gen_graph500_edges against edges made by the UNMODIFIED reference generator (RefGen21::generate_kronecker_range,
    -DDETERMINISTIC seed; tests/golden/graph500_ref.npz): bit for bit, also far into the stream at scale 24."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph500_ref.npz"))
    for key in [k for k in gold.files if k.startswith("edges_")]:
        _, s, first = key.split("_")
        scale, first = int(s[1:]), int(first)
        src, dst = ctx.graph500_edges(scale, first, gold[key].shape[1])
        assert np.array_equal(src, gold[key][0]) and np.array_equal(dst, gold[key][1]), key


def test_graph500_tile_is_the_matrix_genwritematrix_builds(ctx):
    """cb_gen_graph500_tile (loops removed, duplicates summed, optionally A + A^T) against the matrix the reference's own classes
    build with the calls of ReleaseTests/GenWriteMatrix.cpp:96-124 - structure and multiplicity values; whole and as 2 x 2 blocks."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph500_ref.npz"))
    for scale, ef, sym in [(10, 16, 1), (10, 16, 0), (9, 8, 1)]:
        want = gold[f"genwrite_s{scale}_ef{ef}_sym{sym}"]
        n = 1 << scale
        t = ctx.gen_graph500_tile(scale, ef, symmetric=bool(sym), remove_loops=True, val_dtype=cb.I32)
        rowptr, col, vals = t.to_csr(np.int32)
        rows = np.repeat(np.arange(n), np.diff(rowptr))
        assert np.array_equal(rows, want[0]) and np.array_equal(col, want[1]) and np.array_equal(vals.astype(np.int64), want[2])
        t.free()
        got = []
        for i in range(2):
            for j in range(2):
                tb = ctx.gen_graph500_tile(scale, ef, symmetric=bool(sym), remove_loops=True, row0=i * n // 2, m=n // 2, col0=j * n // 2, n=n // 2,
                                           val_dtype=cb.I32)
                rp, c, v = tb.to_csr(np.int32)
                r = np.repeat(np.arange(n // 2), np.diff(rp))
                got += list(zip((r + i * n // 2).tolist(), (c + j * n // 2).tolist(), v.tolist()))
                tb.free()
        assert sorted(got) == sorted(zip(want[0].tolist(), want[1].tolist(), want[2].tolist()))
