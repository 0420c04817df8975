"""Generates tests/golden/grid_ref.npz by RUNNING THE UNMODIFIED REFERENCE ON A PROCESS GRID (oracle/_ref/cbref_grid:
/root/reference compiled against the process-per-rank MPI stand-in oracle/mpi_multi; `make -C oracle ref`).
Run in the build container only:

    python tests/golden/make_golden_grid.py

Contents (results of the reference's own multi-process multiply, stored instead of recomputed):
  torus_4 / torus_9 / torus_16   the report line of the SpMMError.cpp program on 2x2, 3x3, 4x4 processes
  <case>_p4 / <case>_p9          Y of Mult_AnXBn_Synch on 2x2 / 3x3 processes for the operands of tests/summa_worker.py
                                 (R-MAT scale 9, edge factor 8, seed 3, ragged m = n-5, n = n-3, k = 13)
  <case>_p4_spmv                 the same product through k x SpMV<SR>(A, FullyDistVec) on 2x2 processes
  layout_<glen>_p<p>_{until,len,owner,lind}   the FullyDistVec distribution computed by the reference's FullyDist.h
  betwcent_p1, betwcent_p4      betweenness-centrality scores written by the reference's own Applications/BetwCent.cpp (unmodified)
                                 on 1 process and on 2x2 processes for the graph of betwcent_input()
  hepth_p4, hepth_p4_report      BASELINE config C1 on 2x2 processes: Applications/hep-th.mtx read by the reference's own
                                 ParallelReadMM on four ranks, x X(k=16, fp64, seed 42) through Mult_AnXBn_Synch
Only numeric outputs are stored; inputs are regenerated from the counter-based generators at test time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SCALE, K = 9, 13
CASES = {  # the cases of tests/summa_worker.py
    "minplus_i32": (O.MIN_PLUS, np.int32, np.int32, "x_minplus"),
    "pt_f64": (O.PLUS_TIMES, np.float64, np.float64, "value"),
    "pt_f32": (O.PLUS_TIMES, np.float32, np.float32, "value"),
    "pt_pat_i64": (O.PLUS_TIMES, None, np.int64, "value"),
    "selmax_i32": (O.MAX_SEL2ND, None, np.int32, "value"),
    "or_and": (O.OR_AND, None, np.uint8, "value"),
}


# FullyDistVec layouts stored from the reference: (global length, processes); short lengths hit the "everything on the last
# processor row / column" branches of FullyDist::Owner (FullyDist.h:117-146)
LAYOUTS = [(11, 4), (8361, 4), (100, 9), (5, 9), (2, 9), (1000, 16)]


# Applications/BetwCent.cpp: <dir>/input.mtx, K4APPROX (2^K4 starting vertices), BATCHSIZE
BC_K4APPROX, BC_BATCH = 6, 32


def betwcent_input(directory):
    """input.mtx for the betweenness-centrality application: R-MAT scale 8 (symmetric pattern), Matrix Market general"""
    n, I, J = O.rmat_matrix(8, 8, seed=5)
    with open(os.path.join(directory, "input.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern general\n")
        f.write(f"{n} {n} {len(I)}\n")
        for i, j in zip(I, J):
            f.write(f"{i + 1} {j + 1}\n")
    return n


def operands(case, scale=SCALE, k=K):
    """(semiring, m, n, I, J, V, X) exactly as tests/summa_worker.py builds them with --ragged 1."""
    sr, adt, xdt, kind = CASES[case]
    n, I, J = O.rmat_matrix(scale, 8, seed=3)
    m, n = n - 5, n - 3
    keep = (I < m) & (J < n)
    I, J = I[keep], J[keep]
    V = None if adt is None else O.matrix_values(I, J, n, 7, adt)
    X = O.dense_operand(n, k, 9, xdt, kind)
    return sr, m, n, I, J, V, X


def main():
    if not O.ref_grid_available():
        raise SystemExit("oracle/_ref/cbref_grid missing: run `make -C oracle ref` first")
    out = {}
    for p in (4, 9, 16):
        out[f"torus_{p}"] = np.array(O.ref_grid_torus(p))
    for case in CASES:
        sr, m, n, I, J, V, X = operands(case)
        for p in (4, 9):
            Y, _, log = O.ref_grid_spmm(sr, p, m, n, I, J, V, X, via=0)
            assert "agrees with the reference constructor" in log
            out[f"{case}_p{p}"] = Y
        if X.dtype != np.uint8:
            out[f"{case}_p4_spmv"] = O.ref_grid_spmm(sr, 4, m, n, I, J, V, X, via=1)[0]
    for glen, p in LAYOUTS:
        for key, val in O.ref_grid_layout(glen, p).items():
            out[f"layout_{glen}_p{p}_{key}"] = val
    # the reference's own application Applications/BetwCent.cpp on the graph of betwcent_input(): 1 process and 2x2 processes
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory(prefix="cb_betwcent_") as d:
        betwcent_input(d)
        for p, exe, env in ((1, "betwcent_ref", {}), (4, "betwcent_grid", {"CBMPI_NP": "4"})):
            res = os.path.join(d, f"bc_{p}.txt")
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe), d, str(BC_K4APPROX), str(BC_BATCH), res], check=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="1", **env), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out[f"betwcent_p{p}"] = np.loadtxt(res, skiprows=1)[:, 2]
    hep = np.load(os.path.join(OUT, "hepth.npz"))
    Y, report = O.ref_grid_mm("/root/reference/Applications/hep-th.mtx", 4, O.dense_operand(int(hep["n"]), 16, 42, np.float64))
    out["hepth_p4"], out["hepth_p4_report"] = Y, np.array(report)
    np.savez_compressed(os.path.join(OUT, "grid_ref.npz"), **out)
    print("wrote grid_ref.npz:", ", ".join(sorted(out)))


if __name__ == "__main__":
    main()
