"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/libcbref.so,
built from /root/reference by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden.py

Fixtures (inputs as COO triples + the reference's outputs):
  torus.npz       16x16 torus of Applications/SpMMError.cpp:32-33; G*G through Mult_AnXBn_Synch
                  (known answer "112 nnz", SpMMError.cpp:80) as triples, and G x dense(G).
  hepth.npz       Applications/hep-th.mtx via ParallelReadMM (31502 nnz after symmetric expansion)
                  and Y = A x X(k=16, f64, seed 42) via Mult_AnXBn_Synch  [BASELINE config C1].
  small.npz       ReleaseTests/sevenvertex.mtx (via ParallelReadMM), ReleaseTests/small_nonsym.mtx and
                  largeseq/input1_0 (mixed-sign values => cancellation) x dense k=8 f64.
  rmat10.npz      R-MAT scale 10 (our generator, symmetrised) x k=8 under every semiring/dtype
                  combination of the C ABI, reference outputs for each.
The data files the inputs come from are public test matrices shipped in the reference tree; only
their numeric content (not reference source code) is stored.
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

COMBOS = [  # (name, semiring, A dtype or None=pattern, X dtype, x kind)
    ("pt_f32", O.PLUS_TIMES, np.float32, np.float32, "value"),
    ("pt_f64", O.PLUS_TIMES, np.float64, np.float64, "value"),
    ("pt_i32", O.PLUS_TIMES, np.int32, np.int32, "value"),
    ("pt_i64", O.PLUS_TIMES, np.int64, np.int64, "value"),
    ("pt_pat_i32", O.PLUS_TIMES, None, np.int32, "value"),
    ("pt_pat_i64", O.PLUS_TIMES, None, np.int64, "value"),
    ("pt_pat_f32", O.PLUS_TIMES, None, np.float32, "value"),
    ("pt_pat_f64", O.PLUS_TIMES, None, np.float64, "value"),
    ("or_and", O.OR_AND, None, np.uint8, "value"),
    ("mp_i32", O.MIN_PLUS, np.int32, np.int32, "x_minplus"),
    ("mp_i64", O.MIN_PLUS, np.int64, np.int64, "x_minplus"),
    ("mp_f32", O.MIN_PLUS, np.float32, np.float32, "value"),
    ("mp_f64", O.MIN_PLUS, np.float64, np.float64, "value"),
    ("sm_i32", O.MAX_SEL2ND, None, np.int32, "value"),
    ("sm_i64", O.MAX_SEL2ND, None, np.int64, "value"),
]


def read_triples(path, skip_header_dims=True):
    rows = np.loadtxt(path, skiprows=1 if skip_header_dims else 0, ndmin=2)
    with open(path) as f:
        m, n, _ = (int(t) for t in f.readline().split()[:3])
    return m, n, rows[:, 0].astype(np.int64) - 1, rows[:, 1].astype(np.int64) - 1, rows[:, 2].astype(np.float64)


def main():
    if not O.ref_available():
        raise SystemExit("oracle/_ref/libcbref.so missing: run `make -C oracle ref` first")
    # --- torus
    ti = np.array(list(range(16)) * 4, np.int64)
    tj = np.array([3,0,1,2,7,4,5,6,11,8,9,10,15,12,13,14,1,2,3,0,5,6,7,4,9,10,11,8,13,14,15,12,
                   12,13,14,15,0,1,2,3,4,5,6,7,8,9,10,11,4,5,6,7,8,9,10,11,12,13,14,15,0,1,2,3], np.int64)
    ones = np.ones(64, np.int64)
    CI, CJ, CV = O.ref_spgemm_i64(16, 16, 16, ti, tj, ones, ti, tj, ones)
    G = np.zeros((16, 16), np.int64)
    G[ti, tj] = 1
    Yd, _ = O.ref_spmm(O.PLUS_TIMES, 16, 16, ti, tj, ones, G)
    np.savez_compressed(os.path.join(OUT, "torus.npz"), ti=ti, tj=tj, CI=CI, CJ=CJ, CV=CV, Ydense=Yd)
    assert len(CI) == 112

    # --- hep-th (config C1)
    m, n, I, J, V = O.ref_read_mm(os.path.join(REF, "Applications/hep-th.mtx"))
    X = O.dense_operand(n, 16, 42, np.float64)
    Y, _ = O.ref_spmm(O.PLUS_TIMES, m, n, I, J, V, X)
    Yv, _ = O.ref_spmm(O.PLUS_TIMES, m, n, I, J, V, X[:, :2], via=1)
    np.savez_compressed(os.path.join(OUT, "hepth.npz"), m=m, n=n, I=I.astype(np.int32), J=J.astype(np.int32), V=V,
                        Y=Y, Yspmv2=Yv)
    assert len(I) == 31502

    # --- small general / non-symmetric / cancelling inputs
    small = {}
    m, n, I, J, V = O.ref_read_mm(os.path.join(REF, "ReleaseTests/sevenvertex.mtx"))
    small.update(seven_m=m, seven_n=n, seven_I=I, seven_J=J, seven_V=V)
    X = O.dense_operand(n, 8, 42, np.float64)
    small["seven_Y"] = O.ref_spmm(O.PLUS_TIMES, m, n, I, J, V, X)[0]
    m, n, I, J, V = read_triples(os.path.join(REF, "ReleaseTests/small_nonsym.mtx"))
    small.update(nonsym_m=m, nonsym_n=n, nonsym_I=I, nonsym_J=J, nonsym_V=V)
    X = O.dense_operand(n, 8, 42, np.float64)
    small["nonsym_Y"] = O.ref_spmm(O.PLUS_TIMES, m, n, I, J, V, X)[0]
    m, n, I, J, V = read_triples(os.path.join(REF, "largeseq/input1_0"))
    I, J, V = O.dedup(I, J, V, n, "sum")
    small.update(large_m=m, large_n=n, large_I=I.astype(np.int32), large_J=J.astype(np.int32), large_V=V)
    X = O.dense_operand(n, 8, 42, np.float64) - 0.5          # mixed signs on both sides
    small["large_Y"] = O.ref_spmm(O.PLUS_TIMES, m, n, I, J, V, X)[0]
    np.savez_compressed(os.path.join(OUT, "small.npz"), **small)

    # --- R-MAT scale 10 under every combination
    n, I, J = O.rmat_matrix(10, 16, seed=0)
    out = dict(n=n, I=I.astype(np.int32), J=J.astype(np.int32))
    for name, sr, adt, xdt, kind in COMBOS:
        V = None if adt is None else O.matrix_values(I, J, n, 1, adt)
        X = O.dense_operand(n, 8, 42, xdt, kind)
        Y, _ = O.ref_spmm(sr, n, n, I, J, V, X)
        out["Y_" + name] = Y
    np.savez_compressed(os.path.join(OUT, "rmat10.npz"), **out)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
