"""The device sparse x sparse product (csrc/cb_spgemm.cu) through the C ABI: Mult_AnXBn_Synch / PSpGEMM with a SPARSE
tall-skinny right-hand side - the call Applications/SpMMError.cpp:83 and Applications/BetwCent.cpp:185,204 make.

Checked against (i) the reference's own known answer (torus G*G: 112 nonzeros, 96 twos + 16 fours, SpMMError.cpp:32-33,80),
(ii) the UNMODIFIED reference run on the same operands where oracle/_ref is present (cbref_spgemm_i64: exact triples), and
(iii) a host evaluation of every semiring: same structure (an entry exists exactly where a product exists, also when the
folded value equals the identity), values exact for integers / booleans and within the north star's tolerance for
PlusTimes in floating point.  Shapes: square, tall-skinny with 1 / 10 / 100 nonzeros per column (the shape of the
reference's only published tall-skinny numbers, ReleaseTests/SCALE26RECT8192/), empty operands, a BFS-like frontier."""
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


def host_spgemm(sr, m, I, J, V, BI, BJ, BV, dtype):
    """C = A (x).(+) B on the host: dict keyed (col, row); products folded in ascending inner index, first product stored
    (mtSpGEMM.h:395-423).  Returns column-major sorted triples."""
    brows = {}
    for p in np.lexsort((BJ, BI)):
        brows.setdefault(int(BI[p]), []).append((int(BJ[p]), BV[p] if BV is not None else dtype(1)))
    acc = {}
    order = np.lexsort((J, I))                                     # A row-major, columns ascending = ascending inner index
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    for p in order:
        i, kk = int(I[p]), int(J[p])
        if kk not in brows:
            continue
        a = V[p] if V is not None else dtype(1)
        for j, b in brows[kk]:
            if sr == O.PLUS_TIMES:
                prod = dtype(a) * dtype(b)
            elif sr == O.MIN_PLUS:
                prod = dtype(info.max) if (a == info.max or b == info.max) else dtype(int(a) + int(b))
            elif sr == O.MAX_SEL2ND:
                prod = dtype(b)
            else:
                prod = dtype(bool(a) and bool(b))
            key = (j, i)
            if key not in acc:
                acc[key] = prod
            elif sr == O.PLUS_TIMES:
                acc[key] = dtype(prod + acc[key])
            elif sr == O.MIN_PLUS:
                acc[key] = min(prod, acc[key])
            elif sr == O.MAX_SEL2ND:
                acc[key] = max(prod, acc[key])
            else:
                acc[key] = dtype(bool(prod) or bool(acc[key]))
    keys = sorted(acc)
    return (np.array([k[1] for k in keys], np.int64), np.array([k[0] for k in keys], np.int64), np.array([acc[k] for k in keys], dtype))


def sparse_rhs(n, k, per_col, seed, dtype, minplus=False):
    rng = np.random.default_rng(seed)
    BI = np.concatenate([rng.choice(n, min(per_col, n), replace=False) for _ in range(k)]).astype(np.int64)
    BJ = np.repeat(np.arange(k, dtype=np.int64), min(per_col, n))
    if dtype == np.uint8:
        BV = rng.integers(0, 2, len(BI)).astype(np.uint8)          # explicit false entries too: they still create structure
    elif np.issubdtype(dtype, np.integer):
        BV = rng.integers(-5 if not minplus else 1, 100, len(BI)).astype(dtype)
        if minplus:
            BV[rng.random(len(BI)) < 0.05] = np.iinfo(dtype).max
    else:
        BV = rng.random(len(BI)).astype(dtype) + dtype(0.25)
    return BI, BJ, BV


def run(ctx, sr, n, I, J, V, k, BI, BJ, BV, dtype):
    A = ctx.tile_from_coo(n, n, I, J, V)
    B = ctx.tile_from_coo(n, k, BI, BJ, BV)
    ci, cj, cv = ctx.spgemm_local(A, B, sr, dtype)
    A.free()
    B.free()
    return ci, cj, cv


def test_torus_known_answer(ctx):
    """SpMMError.cpp: G = 16-vertex torus (4 neighbours each), G*G has 112 nonzeros: 96 twos and 16 fours."""
    n = 16
    I, J = [], []
    for v in range(n):
        r, c = divmod(v, 4)
        for rr, cc in ((r, (c + 1) % 4), (r, (c - 1) % 4), ((r + 1) % 4, c), ((r - 1) % 4, c)):
            I.append(v)
            J.append(rr * 4 + cc)
    I, J = np.array(I, np.int64), np.array(J, np.int64)
    V = np.ones(len(I), np.int64)
    ci, cj, cv = run(ctx, cb.PLUS_TIMES, n, I, J, V, n, I, J, V, np.int64)
    assert len(ci) == 112 and (cv == 2).sum() == 96 and (cv == 4).sum() == 16
    assert (np.diff(cj * n + ci) > 0).all()                        # column-major, no duplicates


CASES = [("plus_times_i64", cb.PLUS_TIMES, np.int64, np.int64), ("plus_times_f64", cb.PLUS_TIMES, np.float64, np.float64),
         ("plus_times_f32", cb.PLUS_TIMES, np.float32, np.float32), ("plus_times_bool_i32", cb.PLUS_TIMES, None, np.int32),
         ("plus_times_boolvals_f64", cb.PLUS_TIMES, np.uint8, np.float64), ("min_plus_i32", cb.MIN_PLUS, np.int32, np.int32),
         ("select_max_i64", cb.MAX_SEL2ND, None, np.int64), ("or_and", cb.OR_AND, None, np.uint8), ("or_and_boolvals", cb.OR_AND, np.uint8, np.uint8)]


@pytest.mark.parametrize("name,sr,adt,dt", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("per_col", [1, 10, 100])
def test_tall_skinny_sparse_rhs_against_the_host_evaluation(ctx, name, sr, adt, dt, per_col):
    n, I, J = O.rmat_matrix(10, 8, seed=5)
    k = 48
    if adt is None:
        V = None
    elif adt == np.uint8:
        V = (O.hash_values(I * n + J, 3, np.uint8)).astype(np.uint8)                      # stored booleans, some false
    else:
        V = O.matrix_values(I, J, n, 1, adt)
        if sr == cb.MIN_PLUS:
            V[::37] = np.iinfo(adt).max
    BI, BJ, BV = sparse_rhs(n, k, per_col, 11 + per_col, dt, minplus=(sr == cb.MIN_PLUS))
    ci, cj, cv = run(ctx, sr, n, I, J, V, k, BI, BJ, BV, dt)
    osr = {cb.PLUS_TIMES: O.PLUS_TIMES, cb.MIN_PLUS: O.MIN_PLUS, cb.MAX_SEL2ND: O.MAX_SEL2ND, cb.OR_AND: O.OR_AND}[sr]
    Vh = None if V is None else (V.astype(bool) if adt == np.uint8 else V)
    ri, rj, rv = host_spgemm(osr, n, I, J, Vh, BI, BJ, BV, dt)
    assert np.array_equal(ci, ri) and np.array_equal(cj, rj), "structure differs"
    if np.issubdtype(dt, np.floating):
        tol = 1e-5 if dt == np.float32 else 1e-12
        assert (np.abs(cv - rv) <= tol * np.abs(rv)).all()
    else:
        assert np.array_equal(cv, rv)


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref (compiled reference) not present")
def test_square_product_equals_the_unmodified_reference(ctx):
    """A*A under PlusTimesSRing<int64,int64> through the reference's Mult_AnXBn_Synch: identical triples."""
    n, I, J = O.rmat_matrix(9, 6, seed=2)
    V = O.matrix_values(I, J, n, 1, np.int64)
    RI, RJ, RV = O.ref_spgemm_i64(n, n, n, I, J, V, I, J, V)
    ci, cj, cv = run(ctx, cb.PLUS_TIMES, n, I, J, V, n, I, J, V, np.int64)
    o = np.lexsort((RI, RJ))
    assert np.array_equal(ci, RI[o]) and np.array_equal(cj, RJ[o]) and np.array_equal(cv, RV[o])


def test_empty_operands_and_frontier(ctx):
    n, I, J = O.rmat_matrix(11, 8, seed=1)
    e = np.zeros(0, np.int64)
    ci, cj, cv = run(ctx, cb.PLUS_TIMES, n, I, J, None, 8, e, e, np.zeros(0, np.int32), np.int32)
    assert len(ci) == 0
    ci, cj, cv = run(ctx, cb.PLUS_TIMES, n, e, e, None, 8, np.arange(8), np.arange(8), np.ones(8, np.int32), np.int32)
    assert len(ci) == 0
    # a BFS frontier: one source per column (Applications/BetwCent.cpp:185): the product is the neighbourhood of each source
    src = np.array([3, 77, 512, 1999], np.int64)
    ci, cj, cv = run(ctx, cb.PLUS_TIMES, n, I, J, None, 4, src, np.arange(4), np.ones(4, np.int32), np.int32)
    for c, s in enumerate(src):
        assert np.array_equal(np.sort(ci[cj == c]), np.sort(I[J == s])) and (cv[cj == c] == 1).all()


def test_larger_product_structure_and_checksum(ctx):
    """R-MAT scale 16 times a 2^16 x 256 panel with 100 nonzeros per column (~26 M partial products): nnz and value checksum
    against scipy's CSR product of the same operands in int64."""
    sp = pytest.importorskip("scipy.sparse")
    n, I, J = O.rmat_matrix(16, 16, seed=0)
    k = 256
    BI, BJ, BV = sparse_rhs(n, k, 100, 5, np.int64)
    ci, cj, cv = run(ctx, cb.PLUS_TIMES, n, I, J, None, k, BI, BJ, BV, np.int64)
    A = sp.csr_matrix((np.ones(len(I), np.int64), (I, J)), shape=(n, n))
    B = sp.csr_matrix((BV, (BI, BJ)), shape=(n, k))
    S = (sp.csr_matrix((np.ones(len(I), np.int64), (I, J)), shape=(n, n)) @ sp.csr_matrix((np.ones(len(BI), np.int64), (BI, BJ)), shape=(n, k))).tocoo()
    C = (A @ B).tocsr()
    assert len(ci) == S.nnz                                         # structure: where any product exists (values may cancel to 0)
    assert np.array_equal(np.asarray(C[ci, cj]).ravel(), cv)
