"""The C++ host layer (include/CombBLAS/*.h) and its driver run END TO END on CPU against tests/mock_abi - a host-memory
stand-in for the C ABI that exists only in the test tree - so SpParMat / DenseParMat / FullyDistVec / SpMM / SpMV /
Mult_AnXBn_Synch / Reduce / EWiseScale logic is exercised by the CPU suite, not only on the GPU box.  The driver verifies
every result against its own host replays with the semiring functors.  One process (1 x 1 grid); the multi-process paths
run on real GPUs (tests/test_host_cpp.py) and the distribution arithmetic is pinned to the reference in
tests/test_ref_grid.py."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "combblas-spmm-test_b200")
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    d = tmp_path_factory.mktemp("mock_abi")
    lib = str(d / "libcombblas_b200.so")
    exe = str(d / "spmm_driver_mock")
    inc = [f"-I{PKG}/include", f"-I{ROOT}/include"]
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wall", *inc, "-o", lib,
                           f"{ROOT}/tests/mock_abi/mock_combblas_b200.cpp"], timeout=300)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-Wall", "-Wno-unused-variable", "-Wno-int-in-bool-context", *inc, "-o", exe, f"{PKG}/host/spmm_driver.cpp",
                           f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}", "-lpthread"], timeout=600)
    return exe


def run(exe, *args):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, *[str(a) for a in args]], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-3000:]
    return r


def run_grid(exe, nproc, rdv, *args):
    """one OS process per rank with the launcher environment the host layer reads (RANK / WORLD_SIZE / LOCAL_RANK)"""
    procs = []
    for r in range(nproc):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(nproc), LOCAL_RANK=str(r), CB_RENDEZVOUS_DIR=str(rdv))
        procs.append(subprocess.Popen([exe, *[str(a) for a in args]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0 and "runtime error" not in se and "AddressSanitizer" not in se, so[-2000:] + se[-3000:]
    return outs[0]


@pytest.mark.parametrize("nproc,grid", [(4, ()), (2, (1, 2)), (2, (2, 1)), (6, (2, 3))])
def test_spmv_fullydistvec_and_dense_epilogues_on_process_grids(driver, tmp_path, nproc, grid):
    so, se = run_grid(driver, nproc, tmp_path, "spmv", 8, *grid)
    assert "SpMV and dense epilogues working correctly" in se and "rows reached" in so


def test_rendezvous_survives_stale_files_and_cleans_up(driver, tmp_path):
    """The file rendezvous behind the host collectives (CombBLAS/cb_mpi.h): a directory that still holds the files of a run
    that died - same names, plausible sizes, another session - must not be read as this run's data (every file carries the
    session nonce agreed on at start-up), and a run that ends normally leaves none of its files behind."""
    rdv = tmp_path / "rdv"
    rdv.mkdir()
    stale = np.random.default_rng(0).integers(0, 255, 136, dtype=np.uint8).tobytes()       # nonce + 128-byte NCCL id of "op 0"
    for q in range(4):
        (rdv / f"c1_op0.r{q}").write_bytes(stale)
        (rdv / f"c1_op1.r{q}").write_bytes(stale[:9])
        (rdv / f"hello.{q}").write_bytes(stale[:8])
        (rdv / f"ack.{q}").write_bytes(stale[:24])
    (rdv / "session").write_bytes(np.full(3 + 256, 7, np.uint64).tobytes())                 # a finished session of that dead run
    for _ in range(2):                                                                      # and twice in a row in the same directory
        so, se = run_grid(driver, 4, rdv, "torus")
        assert so.count("112 nonzeros") == 3 and "SpGEMM (sparse x sparse) working correctly" in se
        left = sorted(f.name for f in rdv.iterdir() if not f.name.startswith("mock"))
        # what may remain are the dead run's files nobody owns any more (ranks only remove their own session files)
        assert all(n.startswith("c1_op") for n in left) and len(left) <= 8, left


def test_spmmerror_program_on_2x2_processes(driver, tmp_path):
    so, se = run_grid(driver, 4, tmp_path, "torus")
    assert so.count("112 nonzeros") == 3 and "SpGEMM (sparse x sparse) working correctly" in se


def test_matrix_market_round_trip_on_2x2_processes(driver, tmp_path):
    from tests.test_host_cpp import write_mtx
    g = np.load(os.path.join(G, "small.npz"))
    m, n = int(g["nonsym_m"]), int(g["nonsym_n"])
    mtx = str(tmp_path / "a.mtx")
    write_mtx(mtx, m, n, g["nonsym_I"], g["nonsym_J"], g["nonsym_V"])
    so, se = run_grid(driver, 4, tmp_path / "rdv", "mtx", mtx, 8, str(tmp_path / "y.bin"), str(tmp_path / "copy.mtx"))
    assert "Matrix Market round trip working correctly" in se
    rows = [l.split() for l in open(tmp_path / "copy.mtx").read().splitlines()[2:]]
    got = sorted((int(r[0]) - 1, int(r[1]) - 1, float(r[2])) for r in rows)
    want = sorted(zip(g["nonsym_I"].tolist(), g["nonsym_J"].tolist(), g["nonsym_V"].tolist()))
    assert len(got) == len(want) and all(a[:2] == b[:2] and abs(a[2] - b[2]) <= 1e-15 * abs(b[2]) for a, b in zip(got, want))


def messy_matrix_market(path, hard):
    """a general real file the way tools in the wild write them: tabs, CRLF, blank and comment lines between the entries, signs,
    exponents, leading zeros and dots - and, with `hard`, one value with more digits than the device parser reproduces exactly"""
    rng = np.random.default_rng(5)
    m, n = 37, 29
    cells = rng.choice(m * n, 300, replace=False)
    toks = []
    for q, c in enumerate(cells):
        v = rng.standard_normal() * 10.0 ** int(rng.integers(-8, 9))
        toks.append([f"{v:.17g}", f"{v:.6e}", f"{v:+.3f}", f"{int(v * 100) / 100:g}", f"{abs(v):.0f}.", f"0{abs(v):.2f}", f"{v:.10E}"][q % 7])
    if hard:
        toks[11] = "0.1234567890123456789012345"
    with open(path, "w", newline="") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% a comment\n")
        f.write(f"{m} {n} {len(cells)}\n")
        for q, (c, t) in enumerate(zip(cells, toks)):
            sep, eol = ("\t", "\r\n") if q % 3 == 0 else ("  ", "\n") if q % 3 == 1 else (" ", " \n")
            f.write(f"{c // n + 1}{sep}{c % n + 1}{sep}{t}{eol}")
            if q % 50 == 7:
                f.write("\n% a comment in the data section\n")
    return m, n, sorted((int(c // n), int(c % n), float(t)) for c, t in zip(cells, toks))


@pytest.mark.parametrize("hard", [False, True])
def test_matrix_market_text_share_parser(driver, tmp_path, hard):
    # every process hands its byte range of the file to cb_tile_from_mm_text; a number outside the parser's exact range sends
    # every process to the host parser instead - the same matrix either way, value for value what float() reads
    mtx = str(tmp_path / "messy.mtx")
    m, n, want = messy_matrix_market(mtx, hard)
    so, se = run_grid(driver, 4, tmp_path / "rdv", "mtx", mtx, 8, str(tmp_path / "y.bin"), str(tmp_path / "copy.mtx"))
    assert "Matrix Market round trip working correctly" in se
    assert ("parsing it on the host" in so + se) == hard
    rows = [l.split() for l in open(tmp_path / "copy.mtx").read().splitlines()[2:]]
    got = sorted((int(r[0]) - 1, int(r[1]) - 1, float(r[2])) for r in rows)
    assert got == want


def test_spmv_fullydistvec_and_dense_epilogues(driver):
    r = run(driver, "spmv", 8)
    assert "SpMV and dense epilogues working correctly" in r.stderr and "rows reached" in r.stdout


@pytest.mark.parametrize("what", ["pt_f32", "mp_i32", "sel_i64", "bool"])
def test_spmm_every_semiring(driver, what):
    assert "SpMM working correctly" in run(driver, "rmat", 8, 12, what).stderr


def test_spmmerror_program(driver):
    r = run(driver, "torus")
    assert r.stdout.count("112 nonzeros") == 3 and "SpGEMM (sparse x sparse) working correctly" in r.stderr


def test_sparse_right_hand_side(driver):
    assert "SpGEMM (sparse x sparse) working correctly" in run(driver, "spgemm", 8, 20, 3).stderr


def test_sparse_right_hand_side_with_column_filter(driver, tmp_path, monkeypatch):
    # CB_SPGEMM_FILTER=1: A's columns that meet only empty rows of B are dropped before the multiply; same C, entry for entry
    monkeypatch.setenv("CB_SPGEMM_FILTER", "1")
    assert "SpGEMM (sparse x sparse) working correctly" in run(driver, "spgemm", 8, 20, 1).stderr      # 20 of 256 rows of B hold anything
    assert "SpGEMM (sparse x sparse) working correctly" in run(driver, "torus").stderr                   # every row active: filter declines
    one = run(driver, "spgemm", 8, 20, 1).stdout.splitlines()
    so, se = run_grid(driver, 4, tmp_path, "spgemm", 8, 20, 1)                                           # and on 2 x 2 processes:
    info = [l for l in so.splitlines() if l.startswith("As a whole")]
    assert len(info) == 3 and info == [l for l in one if l.startswith("As a whole")]                     # same A, B and C sizes


def test_matrix_market_config_c1(driver, tmp_path):
    from tests.test_host_cpp import write_mtx
    g = np.load(os.path.join(G, "hepth.npz"))
    m, n = int(g["m"]), int(g["n"])
    mtx, dump = str(tmp_path / "hepth.mtx"), str(tmp_path / "y.bin")
    write_mtx(mtx, m, n, g["I"], g["J"], g["V"])
    r = run(driver, "mtx", mtx, 16, dump, str(tmp_path / "copy.mtx"))
    assert "31502 nonzeros" in r.stdout and "SpMM working correctly" in r.stderr and "Matrix Market round trip working correctly" in r.stderr
    with open(tmp_path / "copy.mtx") as f:
        assert f.readline().startswith("%%MatrixMarket matrix coordinate real general") and f.readline().split() == ["8361", "8361", "31502"]
    Y = np.fromfile(dump, np.float64).reshape(m, 16)
    assert (np.abs(Y - g["Y"]) <= 1e-12 * np.maximum(np.abs(g["Y"]), 1e-300)).all()     # host layer + reader vs the reference's golden


REF_DRIVER = "/root/reference/ReleaseTests/MultTiming.cpp"


def write_triples(path, m, n, I, J, V):
    with open(path, "w") as f:
        f.write(f"{m} {n} {len(I)}\n")
        for i, j, v in zip(I, J, V):
            f.write(f"{i + 1} {j + 1} {float(v)!r}\n")


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="the reference tree is not mounted here")
def test_reference_multtiming_driver_compiles_unmodified_and_runs(tmp_path_factory, driver, tmp_path):
    # source-level drop-in: the reference's own ReleaseTests/MultTiming.cpp (ReadDistribute, Mult_AnXBn_DoubleBuff, Mult_AnXBn_Synch,
    # PrintInfo, getnnz, MPI_Pcontrol) compiled as it is against the host layer; nnz(C) checked against the reference library itself
    from oracle import oracle as O
    d = os.path.dirname(driver)
    exe = os.path.join(d, "MultTiming_mock")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-w", f"-I{PKG}/include/mpi_shim", f"-I{PKG}/include", f"-I{ROOT}/include",
                           "-o", exe, REF_DRIVER, f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}", "-lpthread"], timeout=600)
    rng = np.random.default_rng(3)
    m, kd, n = 60, 45, 30
    def rand(mm, nn, nz):
        I, J = rng.integers(0, mm, nz), rng.integers(0, nn, nz)
        keep = np.unique(I * nn + J, return_index=True)[1]
        return I[keep].astype(np.int64), J[keep].astype(np.int64), rng.integers(1, 9, len(keep)).astype(np.int64)
    AI, AJ, AV = rand(m, kd, 300)
    BI, BJ, BV = rand(kd, n, 200)
    a, b = str(tmp_path / "A.txt"), str(tmp_path / "B.txt")
    write_triples(a, m, kd, AI, AJ, AV)
    write_triples(b, kd, n, BI, BJ, BV)
    r = run(exe, a, b)
    if O.ref_available():
        CI, CJ, CV = O.ref_spgemm_i64(m, kd, n, AI, AJ, AV, BI, BJ, BV)
        want = len(CI)
    else:
        A = np.zeros((m, kd), np.int64); A[AI, AJ] = 1
        B = np.zeros((kd, n), np.int64); B[BI, BJ] = 1
        want = int(((A @ B) > 0).sum())
    assert f"C has a total of {want} nonzeros" in r.stderr + r.stdout
    assert r.stdout.count(f"and {want} nonzeros") == 2                      # C.PrintInfo() after DoubleBuff and after Synch
    assert "Double buffered multiplications finished" in r.stdout and "Synchronous multiplications finished" in r.stdout
    # the same unmodified program on 2 x 2 processes (ReadDistribute keeps what each process owns; two grids per process)
    so, se = run_grid(exe, 4, tmp_path / "rdv", a, b)
    assert f"C has a total of {want} nonzeros" in se + so and so.count(f"and {want} nonzeros") == 2


def read_mm_file(path):
    lines = open(path).read().splitlines()
    assert lines[0].startswith("%%MatrixMarket matrix coordinate real general")
    m, n, nnz = (int(v) for v in lines[1].split())
    ent = sorted((int(a) - 1, int(b) - 1, float(c)) for a, b, c in (l.split() for l in lines[2:]))
    assert len(ent) == nnz
    return m, n, ent


@pytest.mark.skipif(not os.path.exists("/root/reference/ReleaseTests/GenWriteMatrix.cpp"), reason="the reference tree is not mounted here")
def test_reference_genwritematrix_driver_compiles_unmodified_and_runs(driver, tmp_path):
    # the reference's generator of its benchmark inputs (CTest `GenWrMat 20 16 1 ...`, ReleaseTests/CMakeLists.txt:41): DistEdgeList,
    # GenGraph500Data, SpParMat(DEL, false), RemoveLoops, Transpose, +=, LoadImbalance, ParallelWriteMM - compiled as it is
    d = os.path.dirname(driver)
    exe = os.path.join(d, "GenWriteMatrix_mock")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-w", f"-I{PKG}/include/mpi_shim", f"-I{PKG}/include", f"-I{ROOT}/include",
                           "-o", exe, "/root/reference/ReleaseTests/GenWriteMatrix.cpp", f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}", "-lpthread"],
                          timeout=600)
    one = str(tmp_path / "one.mtx")
    r = run(exe, 9, 8, 1, one)
    assert "Symmetricized" in r.stderr and "Removed " in r.stderr
    m, n, ent = read_mm_file(one)
    assert m == n == 512 and all(i != j for i, j, _ in ent)
    pairs = {(i, j): v for i, j, v in ent}
    assert all(pairs.get((j, i)) == v for (i, j), v in pairs.items())             # A == A^T after Symmetricize
    # ... and it is THE matrix the reference's own build of this program makes (packed Graph500 stream, duplicates summed):
    # tests/golden/graph500_ref.npz holds the reference's result for `GenWriteMatrix 9 8 1`
    gold = np.load(os.path.join(G, "graph500_ref.npz"))["genwrite_s9_ef8_sym1"]
    want = sorted(zip(gold[0].tolist(), gold[1].tolist(), [float(v) for v in gold[2]]))
    assert sorted((i, j, float(v)) for i, j, v in ent) == want
    four = str(tmp_path / "four.mtx")
    run_grid(exe, 4, tmp_path / "rdv", 9, 8, 1, four)
    assert read_mm_file(four) == (m, n, ent)                                      # the same matrix from a 2 x 2 process grid


@pytest.mark.skipif(not os.path.exists("/root/reference/ReleaseTests/TransposeTest.cpp"), reason="the reference tree is not mounted here")
def test_reference_transposetest_driver_compiles_unmodified_and_passes(driver, tmp_path):
    # the reference's own self-checking ReleaseTests/TransposeTest.cpp: ReadDistribute of a matrix and of its transpose (boolean
    # SpParMat<int,bool,SpDCCols<int,bool>>), operator=, Transpose(), operator==
    d = os.path.dirname(driver)
    exe = os.path.join(d, "TransposeTest_mock")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-w", f"-I{PKG}/include/mpi_shim", f"-I{PKG}/include", f"-I{ROOT}/include",
                           "-o", exe, "/root/reference/ReleaseTests/TransposeTest.cpp", f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}", "-lpthread"],
                          timeout=600)
    g = np.load(os.path.join(G, "small.npz"))
    m, n = int(g["nonsym_m"]), int(g["nonsym_n"])
    ones = np.ones(len(g["nonsym_I"]))
    write_triples(str(tmp_path / "a.txt"), m, n, g["nonsym_I"], g["nonsym_J"], ones)
    write_triples(str(tmp_path / "at.txt"), n, m, g["nonsym_J"], g["nonsym_I"], ones)
    assert "Transpose working correctly" in run(exe, tmp_path, "a.txt", "at.txt").stderr
    so, se = run_grid(exe, 4, tmp_path / "rdv", tmp_path, "a.txt", "at.txt")
    assert "Transpose working correctly" in se
    write_triples(str(tmp_path / "bad.txt"), n, m, g["nonsym_J"][:-1], g["nonsym_I"][:-1], ones[:-1])       # one entry missing
    assert "ERROR in transpose" in run(exe, tmp_path, "a.txt", "bad.txt").stderr


def write_vector(path, vals):
    with open(path, "w") as f:
        f.write(f"{len(vals)} 1 {len(vals)}\n")
        for i, v in enumerate(vals):
            f.write(f"{i + 1} 1 {float(v)!r}\n")


@pytest.mark.skipif(not os.path.exists("/root/reference/ReleaseTests/ReduceTest.cpp"), reason="the reference tree is not mounted here")
def test_reference_reducetest_driver_compiles_unmodified_and_passes(driver, tmp_path):
    # the reference's self-checking ReleaseTests/ReduceTest.cpp: SpParMat::Reduce along rows and columns against vectors read with
    # FullyDistVec::ReadDistribute, compared with the error-tolerant operator==
    d = os.path.dirname(driver)
    exe = os.path.join(d, "ReduceTest_mock")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-w", f"-I{PKG}/include/mpi_shim", f"-I{PKG}/include", f"-I{ROOT}/include",
                           "-o", exe, "/root/reference/ReleaseTests/ReduceTest.cpp", f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}", "-lpthread"],
                          timeout=600)
    g = np.load(os.path.join(G, "small.npz"))
    m, n = int(g["nonsym_m"]), int(g["nonsym_n"])
    I, J, V = g["nonsym_I"], g["nonsym_J"], g["nonsym_V"]
    a, cs, rs = str(tmp_path / "a.txt"), str(tmp_path / "colsums.txt"), str(tmp_path / "rowsums.txt")
    write_triples(a, m, n, I, J, V)
    write_vector(cs, np.bincount(J, weights=V, minlength=n))
    write_vector(rs, np.bincount(I, weights=V, minlength=m))
    assert "Reduction via summation working correctly" in run(exe, a, cs, rs).stderr
    so, se = run_grid(exe, 4, tmp_path / "rdv", a, cs, rs)
    assert "Reduction via summation working correctly" in se
    wrong = np.bincount(I, weights=V, minlength=m)
    wrong[m // 2] += 1.0
    write_vector(rs, wrong)
    assert "ERROR in Reduce via summation" in run(exe, a, cs, rs).stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/Applications/BetwCent.cpp"), reason="the reference tree is not mounted here")
def test_reference_betwcent_application_unmodified_matches_the_reference(driver, tmp_path):
    # Applications/BetwCent.cpp - the application that calls the tall-skinny PSpGEMM with DenseParMat
    # updates around it - compiled as it is against the host layer.  Its scores are compared with the scores the reference itself
    # wrote for the same graph on 1 process and on 2x2 processes (tests/golden/grid_ref.npz; the two differ by up to 670 because
    # every processor column picks its own roots, so the 2x2 comparison checks the distributed semantics too).
    from tests.golden.make_golden_grid import BC_BATCH, BC_K4APPROX, betwcent_input
    d = os.path.dirname(driver)
    exe = os.path.join(d, "BetwCent_mock")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-w", f"-I{PKG}/include/mpi_shim", f"-I{PKG}/include", f"-I{ROOT}/include",
                           "-o", exe, "/root/reference/Applications/BetwCent.cpp", f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}", "-lpthread"],
                          timeout=600)
    gold = np.load(os.path.join(G, "grid_ref.npz"))
    betwcent_input(str(tmp_path))
    one = str(tmp_path / "bc1.txt")
    r = run(exe, tmp_path, BC_K4APPROX, BC_BATCH, one)
    assert "Computation finished" in r.stdout
    got = np.loadtxt(one, skiprows=1)[:, 2]
    assert np.abs(got - gold["betwcent_p1"]).max() <= 1e-9 * np.abs(gold["betwcent_p1"]).max()
    # the same run with the column filter for sparse right-hand sides (CB_SPGEMM_FILTER=1): BFS frontiers are sparse, the filter engages
    os.environ["CB_SPGEMM_FILTER"] = "1"
    try:
        filt = str(tmp_path / "bc1f.txt")
        run(exe, tmp_path, BC_K4APPROX, BC_BATCH, filt)
    finally:
        del os.environ["CB_SPGEMM_FILTER"]
    assert np.abs(np.loadtxt(filt, skiprows=1)[:, 2] - gold["betwcent_p1"]).max() <= 1e-9 * np.abs(gold["betwcent_p1"]).max()
    four = str(tmp_path / "bc4.txt")
    run_grid(exe, 4, tmp_path / "rdv", tmp_path, BC_K4APPROX, BC_BATCH, four)
    got4 = np.loadtxt(four, skiprows=1)[:, 2]
    assert np.abs(got4 - gold["betwcent_p4"]).max() <= 1e-9 * np.abs(gold["betwcent_p4"]).max()
    assert np.abs(gold["betwcent_p1"] - gold["betwcent_p4"]).max() > 1.0          # the two references really differ


def galerkin_inputs(d, n=40, m=12, seed=11):
    """A = L + diag(D) (triples files A, L; vector file D) and a restriction matrix T (n x m) for ReleaseTests/GalerkinNew.cpp"""
    rng = np.random.default_rng(seed)
    LI, LJ = rng.integers(0, n, 160), rng.integers(0, n, 160)
    keep = (LI != LJ)
    key = np.unique(LI[keep] * n + LJ[keep])
    LI, LJ = key // n, key % n
    LV = rng.integers(1, 9, len(LI)).astype(float)
    D = rng.integers(1, 9, n).astype(float)
    TI, TJ = rng.integers(0, n, 60), rng.integers(0, m, 60)
    key = np.unique(TI * m + TJ)
    TI, TJ = key // m, key % m
    TV = rng.integers(1, 5, len(TI)).astype(float)
    write_triples(os.path.join(d, "A.txt"), n, n, np.concatenate([LI, np.arange(n)]), np.concatenate([LJ, np.arange(n)]), np.concatenate([LV, D]))
    write_triples(os.path.join(d, "L.txt"), n, n, LI, LJ, LV)
    write_vector(os.path.join(d, "D.txt"), D)
    write_triples(os.path.join(d, "T.txt"), n, m, TI, TJ, TV)
    return [os.path.join(d, f) for f in ("A.txt", "L.txt", "D.txt", "T.txt")]


@pytest.mark.skipif(not os.path.exists("/root/reference/ReleaseTests/GalerkinNew.cpp"), reason="the reference tree is not mounted here")
def test_reference_galerkinnew_driver_compiles_unmodified_and_passes(driver, tmp_path):
    # the reference's self-checking SpGEMM test ReleaseTests/GalerkinNew.cpp: S(AT) against SLT + (S scaled by D)T - five
    # PSpGEMMs, Transpose, DimApply, +=, error-tolerant == ("Splitting approach is correct")
    d = os.path.dirname(driver)
    exe = os.path.join(d, "GalerkinNew_mock")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-w", f"-I{PKG}/include/mpi_shim", f"-I{PKG}/include", f"-I{ROOT}/include",
                           "-o", exe, "/root/reference/ReleaseTests/GalerkinNew.cpp", f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}", "-lpthread"],
                          timeout=600)
    files = galerkin_inputs(str(tmp_path))
    assert "Splitting approach is correct" in run(exe, *files).stderr
    so, se = run_grid(exe, 4, tmp_path / "rdv", *files)
    assert "Splitting approach is correct" in se


def test_prebuilt_reference_binaries_for_the_gpu_box_are_current(driver, tmp_path):
    # oracle/_ref/*_b200 are the binaries tests/test_host_cpp.py runs on the GPU box (the reference tree is not there to rebuild
    # them); run the very same files here with the mock library in front of their RUNPATH so a stale build cannot travel
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "BetwCent_b200")):
        pytest.skip("oracle/_ref/*_b200 not built (needs the reference tree)")
    from tests.golden.make_golden_grid import BC_BATCH, BC_K4APPROX, betwcent_input
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "host"), "reference_drivers"])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["LD_LIBRARY_PATH"] = os.path.dirname(driver) + ":" + env.get("LD_LIBRARY_PATH", "")
    betwcent_input(str(tmp_path))
    out = str(tmp_path / "bc.txt")
    r = subprocess.run([os.path.join(ref, "BetwCent_b200"), str(tmp_path), str(BC_K4APPROX), str(BC_BATCH), out], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "Computation finished" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    gold = np.load(os.path.join(G, "grid_ref.npz"))["betwcent_p1"]
    assert np.abs(np.loadtxt(out, skiprows=1)[:, 2] - gold).max() <= 1e-9 * np.abs(gold).max()
    files = galerkin_inputs(str(tmp_path))
    r = subprocess.run([os.path.join(ref, "GalerkinNew_b200"), *files], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "Splitting approach is correct" in r.stderr, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.parametrize("nproc,grid", [(4, ()), (6, (2, 3))])
def test_host_layer_with_a_real_mpi_has_row_and_column_communicators(tmp_path, nproc, grid):
    """-DCB_HAVE_MPI: CommGrid must split the world into processor-row / processor-column communicators the way the
    reference does (src/CommGrid.cpp:66-67); user code reduces and broadcasts over GetRowWorld() / GetColWorld().
    The MPI is the process-per-rank stand-in of oracle/mpi_multi, the C ABI the test-only mock."""
    d = tmp_path
    inc = [f"-I{PKG}/include", f"-I{ROOT}/include"]
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-fPIC", "-shared", *inc, "-o", str(d / "libcombblas_b200.so"),
                           f"{ROOT}/tests/mock_abi/mock_combblas_b200.cpp"], timeout=300)
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-w", "-DCB_HAVE_MPI", "-Dmain=cb_rank_main", f"-I{ROOT}/oracle/mpi_multi", *inc,
                           "-c", f"{ROOT}/tests/hostmpi/commgrid_mpi_test.cpp", "-o", str(d / "t.o")], timeout=300)
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-w", f"-I{ROOT}/oracle/mpi_multi", "-c", f"{ROOT}/oracle/mpi_multi/cbmpi.cpp",
                           "-o", str(d / "cbmpi.o")], timeout=300)
    subprocess.check_call(["/usr/bin/g++", "-o", str(d / "t"), str(d / "t.o"), str(d / "cbmpi.o"), f"-L{d}", "-lcombblas_b200", f"-Wl,-rpath,{d}",
                           "-lpthread"], timeout=300)
    env = dict({k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}, CBMPI_NP=str(nproc))
    r = subprocess.run([str(d / "t"), *[str(g) for g in grid]], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "CB_HAVE_MPI grid working correctly" in r.stdout, r.stdout + r.stderr
