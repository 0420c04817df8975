"""parity with the reference's OWN multi-process multiply.

tests/golden/grid_ref.npz holds the results of the unmodified reference run on 2x2, 3x3 (and the SpMMError program on 4x4)
process grids in the build container (oracle/_ref/cbref_grid = /root/reference compiled against the process-per-rank MPI
stand-in oracle/mpi_multi; generator: tests/golden/make_golden_grid.py).  Checked here on CPU:
  * the C restatement's emulated SUMMA stage loop and its 1-rank multiply against those stored results;
  * the host-side stage loop of the product's plan (summa_worker --mode cpu over gloo) against them;
  * where cbref_grid exists (the build container), live runs of all four reference multiplies (Synch, DoubleBuff, Overlap,
    k x SpMV) on fresh random operands, and the MPI stand-in's own collectives.
The GPU legs (tests/test_summa_gpu.py, tests/test_spmm_gpu.py) compare the product with the same file."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden.make_golden_grid import CASES, K, SCALE, operands
from tests.test_summa_cpu import torchrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "grid_ref.npz")
needs_grid = pytest.mark.skipif(not O.ref_grid_available(), reason="oracle/_ref/cbref_grid not built (needs /root/reference)")


def close(Y, ref):
    if np.issubdtype(ref.dtype, np.floating):
        tol = 1e-5 if ref.dtype == np.float32 else 1e-12
        return bool((np.abs(Y.astype(np.float64) - ref) <= tol * np.maximum(np.abs(ref.astype(np.float64)), 1e-300)).all())
    return bool(np.array_equal(Y, ref))


def test_torus_program_on_process_grids():
    g = np.load(GOLD)
    for p in (4, 9, 16):
        line = str(g[f"torus_{p}"])
        # Applications/SpMMError.cpp:80 "The nnz values should be 112, 112, 112"; 96 twos + 16 fours (torus.npz)
        assert "products nnz 112 112 112" in line and "twos 96 fours 16" in line and "G13==G12 1 G23==G12 1" in line


@pytest.mark.parametrize("case", list(CASES))
def test_restatement_matches_reference_run_on_process_grids(case):
    g = np.load(GOLD)
    sr, m, n, I, J, V, X = operands(case)
    one = O.spmm(sr, m, n, I, J, V, X)
    for p, q in ((4, 2), (9, 3)):
        gold = g[f"{case}_p{p}"]
        assert gold.shape == one.shape
        assert close(one, gold), f"1-rank restatement vs reference on {p} processes"
        em = O.spmm_summa(sr, q, q, m, n, I, J, V, X)
        assert close(em, gold), f"emulated {q}x{q} SUMMA vs reference on {p} processes"
        if p == 4:
            # two stages: MultiwayMerge adds the stage partials in stage order, as the emulated loop does -> bit-identical
            assert np.array_equal(em, gold)
    if f"{case}_p4_spmv" in g.files:
        assert close(one, g[f"{case}_p4_spmv"])


def test_hepth_config_c1_on_2x2_processes():
    # the reference's own ParallelReadMM on four ranks + Mult_AnXBn_Synch, against our Matrix Market reader + the restatement
    g, hep = np.load(GOLD), np.load(os.path.join(ROOT, "tests", "golden", "hepth.npz"))
    assert "8361 x 8361, 31502 nonzeros on a 2 x 2 grid; C: 121760 stored entries" in str(g["hepth_p4_report"])   # hep-th-p4.txt:9
    m, n = int(hep["m"]), int(hep["n"])
    X = O.dense_operand(n, 16, 42, np.float64)
    gold = g["hepth_p4"]
    assert close(O.spmm(O.PLUS_TIMES, m, n, hep["I"], hep["J"], hep["V"], X), gold)
    # same stages; inside a stage the reference takes its heap branch for columns with flops/nnz < 2 (mtSpGEMM.h:336-347), whose
    # summation order is not ascending, so a few percent of the entries differ in the last bit
    em = O.spmm_summa(O.PLUS_TIMES, 2, 2, m, n, hep["I"], hep["J"], hep["V"], X)
    assert close(em, gold) and (em == gold).mean() > 0.9
    assert close(hep["Y"], gold)                                                                              # 1 process vs 4 processes


def test_fullydistvec_layout_matches_reference():
    # the host layer's FullyDistLayout (include/CombBLAS/FullyDistVec.h) against LengthUntil / MyLocLength / Owner of the
    # reference's FullyDist.h evaluated on real process grids
    from tests.golden.make_golden_grid import LAYOUTS
    exe = os.path.join(ROOT, "combblas-spmm-test_b200", "host", "host_logic_test")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "combblas-spmm-test_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    g = np.load(GOLD)
    for glen, p in LAYOUTS:
        q = int(round(p ** 0.5))
        r = subprocess.run([exe, "fdv", str(glen), str(q), str(q)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        got = {l.split()[0]: np.array([int(v) for v in l.split()[1:]], np.int64) for l in r.stdout.splitlines()}
        for key in ("until", "len", "owner", "lind"):
            assert np.array_equal(got[key], g[f"layout_{glen}_p{p}_{key}"]), (glen, p, key)


def test_host_stage_loop_matches_reference_run_2x2():
    r = torchrun(4, ["--mode", "cpu", "--pr", "2", "--pc", "2", "--scale", str(SCALE), "--k", str(K), "--golden", GOLD,
                     "--cases", "minplus_i32,pt_f64,or_and"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("processes: same") == 3 and r.stdout.count(": ok") == 3


@needs_grid
def test_mpi_standin_selftest(tmp_path):
    exe = str(tmp_path / "selftest")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Dmain=cb_rank_main", f"-I{ROOT}/oracle/mpi_multi",
                           f"{ROOT}/oracle/mpi_multi/selftest.cpp", f"{ROOT}/oracle/mpi_multi/cbmpi.cpp", "-o", exe], timeout=300)
    for p in (1, 4, 9):
        r = subprocess.run([exe], env=dict(os.environ, CBMPI_NP=str(p), CBMPI_TIMEOUT="60"), capture_output=True, text=True, timeout=90)
        assert r.returncode == 0 and f"cbmpi selftest ok on {p} ranks" in r.stdout, r.stdout + r.stderr


@needs_grid
def test_live_reference_grid_all_multiply_variants():
    assert "products nnz 112 112 112" in O.ref_grid_torus(4)
    rng = np.random.default_rng(5)
    m, n, k, nnz = 83, 71, 7, 900
    I, J = rng.integers(0, m, nnz), rng.integers(0, n, nnz)
    keep = np.unique(I * n + J, return_index=True)[1]
    I, J = I[keep].astype(np.int64), J[keep].astype(np.int64)
    for sr, V, X in [
        (O.MIN_PLUS, rng.integers(1, 50, len(I)).astype(np.int64), rng.integers(1, 50, (n, k)).astype(np.int64)),
        (O.PLUS_TIMES, rng.standard_normal(len(I)), rng.standard_normal((n, k))),       # mixed signs: cancellation
        (O.MAX_SEL2ND, None, rng.integers(-5, 50, (n, k)).astype(np.int32)),            # values below the identity -1
        (O.OR_AND, None, (rng.random((n, k)) < 0.3).astype(np.uint8)),
    ]:
        one = O.spmm(sr, m, n, I, J, V, X)
        for p in (4, 9):
            for via in (0, 2, 3) + ((1,) if X.dtype != np.uint8 else ()):
                Y, _, log = O.ref_grid_spmm(sr, p, m, n, I, J, V, X, via=via)
                assert "agrees with the reference constructor" in log
                if np.issubdtype(one.dtype, np.floating):
                    assert np.abs(Y - one).max() <= 1e-12 * np.abs(one).max(), (sr, p, via)
                else:
                    assert np.array_equal(Y, one), (sr, p, via)
