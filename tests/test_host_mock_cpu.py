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


def test_matrix_market_config_c1(driver, tmp_path):
    from tests.test_host_cpp import write_mtx
    g = np.load(os.path.join(G, "hepth.npz"))
    m, n = int(g["m"]), int(g["n"])
    mtx, dump = str(tmp_path / "hepth.mtx"), str(tmp_path / "y.bin")
    write_mtx(mtx, m, n, g["I"], g["J"], g["V"])
    r = run(driver, "mtx", mtx, 16, dump)
    assert "31502 nonzeros" in r.stdout and "SpMM working correctly" in r.stderr
    Y = np.fromfile(dump, np.float64).reshape(m, 16)
    assert (np.abs(Y - g["Y"]) <= 1e-12 * np.maximum(np.abs(g["Y"]), 1e-300)).all()     # host layer + reader vs the reference's golden
