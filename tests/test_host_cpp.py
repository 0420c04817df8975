"""The C++ host layer (combblas-spmm-test_b200/include/CombBLAS) through its MultTest-style driver: compiles with
plain g++, refuses to run without a GPU, and on a B200 reproduces the reference golden for hep-th (config C1)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "combblas-spmm-test_b200", "host")
DRIVER = os.path.join(HOST, "spmm_driver")
G = os.path.join(ROOT, "tests", "golden")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "combblas-spmm-test_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", HOST])


def test_driver_compiles_against_the_c_abi():
    build()
    assert os.path.exists(DRIVER)


def test_driver_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    build()
    r = subprocess.run([DRIVER, "rmat", "8", "4", "pt_f32"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr


def write_mtx(path, m, n, I, J, V):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{m} {n} {len(I)}\n")
        for i, j, v in zip(I, J, V):
            f.write(f"{i + 1} {j + 1} {float(v)!r}\n")


@pytest.mark.gpu
def test_driver_hepth_matches_reference_golden(tmp_path):
    build()
    g = np.load(os.path.join(G, "hepth.npz"))
    m, n = int(g["m"]), int(g["n"])
    mtx, dump = str(tmp_path / "hepth.mtx"), str(tmp_path / "y.bin")
    write_mtx(mtx, m, n, g["I"], g["J"], g["V"])
    r = subprocess.run([DRIVER, "mtx", mtx, "16", dump, str(tmp_path / "copy.mtx")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "31502 nonzeros" in r.stdout and "SpMM working correctly" in r.stderr and "Matrix Market round trip working correctly" in r.stderr
    Y = np.fromfile(dump, np.float64).reshape(m, 16)
    ref = g["Y"]
    assert (np.abs(Y - ref) <= 1e-12 * np.maximum(np.abs(ref), 1e-300)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["pt_f32", "mp_i32", "sel_i64", "bool"])
def test_driver_generated_matrix_every_semiring(what):
    build()
    r = subprocess.run([DRIVER, "rmat", "12", "24", what], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SpMM working correctly" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_driver_spmmerror_program_sparse_times_sparse():
    # Applications/SpMMError.cpp: three products of 16x16 torus matrices, "The nnz values should be 112, 112, 112" (:80)
    build()
    r = subprocess.run([DRIVER, "torus"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("112 nonzeros") == 3 and r.stdout.count("64 nonzeros") == 3
    assert "SpGEMM (sparse x sparse) working correctly" in r.stderr


@pytest.mark.gpu
def test_driver_sparse_rhs_min_plus_structure_and_values():
    build()
    r = subprocess.run([DRIVER, "spgemm", "11", "40", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SpGEMM (sparse x sparse) working correctly" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_driver_spmv_fullydistvec_and_dense_epilogues():
    # SpMV<SR>(A, FullyDistVec) under MinPlus / PlusTimes / SelectMax, DenseParMat::Reduce,
    # DenseParMat += SpParMat, SpParMat::EWiseScale; every result checked inside the driver
    build()
    r = subprocess.run([DRIVER, "spmv", "12"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SpMV and dense epilogues working correctly" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_driver_spmv_on_a_process_grid():
    import torch
    from tests.test_summa_cpu import free_port
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    build()
    n = 4 if torch.cuda.device_count() >= 4 else 2
    grid = ["2", "2"] if n == 4 else ["1", "2"]
    def cmd(port):
        return [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), DRIVER, "spmv", "12"] + grid
    from tests.test_summa_cpu import run_retrying_ports
    r = run_retrying_ports(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SpMV and dense epilogues working correctly" in r.stderr, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_reference_multtiming_driver_unmodified_on_the_gpu(tmp_path):
    # oracle/_ref/MultTiming_b200 = the reference's own ReleaseTests/MultTiming.cpp compiled UNMODIFIED against this host layer
    # and libcombblas_b200.so in the build container (host/Makefile target reference_drivers); it travels prebuilt
    exe = os.path.join(ROOT, "oracle", "_ref", "MultTiming_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/MultTiming_b200 was not built (needs the reference tree)")
    from tests.test_host_mock_cpu import write_triples
    rng = np.random.default_rng(3)
    m, kd, n = 600, 450, 64

    def rand(mm, nn, nz):
        I, J = rng.integers(0, mm, nz), rng.integers(0, nn, nz)
        keep = np.unique(I * nn + J, return_index=True)[1]
        return I[keep].astype(np.int64), J[keep].astype(np.int64), rng.integers(1, 9, len(keep)).astype(np.int64)
    AI, AJ, AV = rand(m, kd, 6000)
    BI, BJ, BV = rand(kd, n, 900)
    a, b = str(tmp_path / "A.txt"), str(tmp_path / "B.txt")
    write_triples(a, m, kd, AI, AJ, AV)
    write_triples(b, kd, n, BI, BJ, BV)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, a, b], capture_output=True, text=True, timeout=300, env=env)
    A = np.zeros((m, kd), np.int64); A[AI, AJ] = 1
    B = np.zeros((kd, n), np.int64); B[BI, BJ] = 1
    want = int(((A @ B) > 0).sum())
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"C has a total of {want} nonzeros" in r.stderr + r.stdout and r.stdout.count(f"and {want} nonzeros") == 2
    assert "Synchronous multiplications finished" in r.stdout


@pytest.mark.gpu
def test_reference_transposetest_driver_unmodified_on_the_gpu(tmp_path):
    # the reference's self-checking ReleaseTests/TransposeTest.cpp, compiled unmodified against this layer (oracle/_ref/TransposeTest_b200)
    exe = os.path.join(ROOT, "oracle", "_ref", "TransposeTest_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/TransposeTest_b200 was not built (needs the reference tree)")
    from tests.test_host_mock_cpu import write_triples
    g = np.load(os.path.join(G, "small.npz"))
    m, n = int(g["nonsym_m"]), int(g["nonsym_n"])
    ones = np.ones(len(g["nonsym_I"]))
    write_triples(str(tmp_path / "a.txt"), m, n, g["nonsym_I"], g["nonsym_J"], ones)
    write_triples(str(tmp_path / "at.txt"), n, m, g["nonsym_J"], g["nonsym_I"], ones)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, str(tmp_path), "a.txt", "at.txt"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "Transpose working correctly" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_betwcent_application_unmodified_on_the_gpu(tmp_path):
    # oracle/_ref/BetwCent_b200 = the reference's Applications/BetwCent.cpp compiled UNMODIFIED against this host layer and
    # libcombblas_b200.so: batched BFS + back-propagation, every step a tall-skinny PSpGEMM on the GPU.  Scores against the ones
    # the reference itself wrote for the same graph (tests/golden/grid_ref.npz, betwcent_p1).
    exe = os.path.join(ROOT, "oracle", "_ref", "BetwCent_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/BetwCent_b200 was not built (needs the reference tree)")
    from tests.golden.make_golden_grid import BC_BATCH, BC_K4APPROX, betwcent_input
    betwcent_input(str(tmp_path))
    out = str(tmp_path / "bc.txt")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, str(tmp_path), str(BC_K4APPROX), str(BC_BATCH), out], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "Computation finished" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    gold = np.load(os.path.join(G, "grid_ref.npz"))["betwcent_p1"]
    got = np.loadtxt(out, skiprows=1)[:, 2]
    assert np.abs(got - gold).max() <= 1e-9 * np.abs(gold).max()


@pytest.mark.gpu
def test_reference_betwcent_application_unmodified_on_2x2_gpus(tmp_path):
    # the same unmodified application on a 2x2 GPU grid against the scores the reference wrote on 2x2 PROCESSES (betwcent_p4)
    import torch
    from tests.test_summa_cpu import free_port
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    exe = os.path.join(ROOT, "oracle", "_ref", "BetwCent_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/BetwCent_b200 was not built (needs the reference tree)")
    from tests.golden.make_golden_grid import BC_BATCH, BC_K4APPROX, betwcent_input
    betwcent_input(str(tmp_path))
    out = str(tmp_path / "bc.txt")
    def cmd(port):
        return [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
               "--master-port", str(port), exe, str(tmp_path), str(BC_K4APPROX), str(BC_BATCH), out]
    from tests.test_summa_cpu import run_retrying_ports
    r = run_retrying_ports(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "Computation finished" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    gold = np.load(os.path.join(G, "grid_ref.npz"))["betwcent_p4"]
    got = np.loadtxt(out, skiprows=1)[:, 2]
    assert np.abs(got - gold).max() <= 1e-9 * np.abs(gold).max()


@pytest.mark.gpu
def test_reference_galerkinnew_driver_unmodified_on_the_gpu(tmp_path):
    # the reference's self-checking SpGEMM test ReleaseTests/GalerkinNew.cpp, compiled unmodified against this layer
    exe = os.path.join(ROOT, "oracle", "_ref", "GalerkinNew_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/GalerkinNew_b200 was not built (needs the reference tree)")
    from tests.test_host_mock_cpu import galerkin_inputs
    files = galerkin_inputs(str(tmp_path))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, *files], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "Splitting approach is correct" in r.stderr, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_reference_reducetest_driver_unmodified_on_the_gpu(tmp_path):
    # the reference's self-checking ReleaseTests/ReduceTest.cpp, compiled unmodified against this layer (oracle/_ref/ReduceTest_b200)
    exe = os.path.join(ROOT, "oracle", "_ref", "ReduceTest_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ReduceTest_b200 was not built (needs the reference tree)")
    from tests.test_host_mock_cpu import write_triples, write_vector
    g = np.load(os.path.join(G, "small.npz"))
    m, n = int(g["nonsym_m"]), int(g["nonsym_n"])
    I, J, V = g["nonsym_I"], g["nonsym_J"], g["nonsym_V"]
    a, cs, rs = str(tmp_path / "a.txt"), str(tmp_path / "colsums.txt"), str(tmp_path / "rowsums.txt")
    write_triples(a, m, n, I, J, V)
    write_vector(cs, np.bincount(J, weights=V, minlength=n))
    write_vector(rs, np.bincount(I, weights=V, minlength=m))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, a, cs, rs], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "Reduction via summation working correctly" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_genwritematrix_driver_unmodified_on_the_gpu(tmp_path):
    # oracle/_ref/GenWriteMatrix_b200 = the reference's own ReleaseTests/GenWriteMatrix.cpp (its benchmark-input generator) compiled
    # UNMODIFIED against this host layer: DistEdgeList -> SpParMat on the device generator, RemoveLoops, Transpose, +=, ParallelWriteMM
    exe = os.path.join(ROOT, "oracle", "_ref", "GenWriteMatrix_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/GenWriteMatrix_b200 was not built (needs the reference tree)")
    from tests.test_host_mock_cpu import read_mm_file
    out = str(tmp_path / "scale10.mtx")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, "10", "16", "1", out], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "Symmetricized" in r.stderr, r.stdout + r.stderr
    m, n, ent = read_mm_file(out)
    assert m == n == 1024 and len(ent) > 10000 and all(i != j for i, j, _ in ent)
    pairs = {(i, j): v for i, j, v in ent}
    assert all(pairs.get((j, i)) == v for (i, j), v in pairs.items())
    # the matrix the reference's own build of this program writes for the same arguments (tests/golden/make_golden_graph500.py)
    gold = np.load(os.path.join(G, "graph500_ref.npz"))["genwrite_s10_ef16_sym1"]
    assert sorted((i, j, float(v)) for i, j, v in ent) == sorted(zip(gold[0].tolist(), gold[1].tolist(), [float(v) for v in gold[2]]))


@pytest.mark.gpu
def test_driver_spmmerror_program_on_a_2x2_grid():
    import torch
    from tests.test_summa_cpu import free_port
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    build()
    def cmd(port):
        return [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node=4", "--master-addr",
               "127.0.0.1", "--master-port", str(port), DRIVER, "torus"]
    from tests.test_summa_cpu import run_retrying_ports
    r = run_retrying_ports(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("112 nonzeros") == 3 and "working correctly" in r.stderr, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_driver_multi_process_grid():
    import torch
    from tests.test_summa_cpu import free_port
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    build()
    n = 4 if torch.cuda.device_count() >= 4 else 2
    grid = ["2", "2"] if n == 4 else ["1", "2"]
    def cmd(port):
        return [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), DRIVER, "rmat", "14", "32", "mp_i32"] + grid
    from tests.test_summa_cpu import run_retrying_ports
    r = run_retrying_ports(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "As a whole: 16384 rows" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
