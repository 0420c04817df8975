"""CPU-only checks of the C++ host layer's containers and arithmetic (no GPU, no CommGrid): tile formats, wire-format
round trip, tolerant equality, Matrix Market expansion, semiring functors, ownership rule - against the numpy
restatements in oracle/oracle.py and the reference's documented behaviour."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "combblas-spmm-test_b200", "host")
EXE = os.path.join(HOST, "host_logic_test")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "combblas-spmm-test_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", HOST])


def run(*args):
    r = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    out = {}
    for line in r.stdout.splitlines():
        key, *vals = line.split()
        out[key] = vals
    return out


def ints(v):
    return np.array([int(x) for x in v], np.int64)


def test_tile_formats_match_the_numpy_restatement(tmp_path):
    rng = np.random.default_rng(0)
    m, n = 57, 91
    I, J = rng.integers(0, m, 600), rng.integers(0, n, 600)
    J[J % 7 == 0] = 3                                             # empty columns -> nzc < n, and duplicates to merge
    V = rng.standard_normal(600).round(6)
    path = tmp_path / "t.txt"
    with open(path, "w") as f:
        for i, j, v in zip(I, J, V):
            f.write(f"{i} {j} {float(v)!r}\n")
    out = run("tile", m, n, path)
    Iu, Ju, Vu = O.dedup(I, J, V, n, "sum")
    cp, jc, ir, numx = O.to_dcsc(m, n, Iu, Ju, Vu)
    assert int(out["nnz"][0]) == len(Iu)
    assert ints(out["ess"]).tolist() == [len(Iu), m, n, len(jc)]            # {nnz, m, n, nzc}, SpDCCols.cpp:787-795
    assert np.array_equal(ints(out["cp"]), cp) and np.array_equal(ints(out["jc"]), jc) and np.array_equal(ints(out["ir"]), ir)
    assert np.allclose(np.array(out["numx"], float), numx, rtol=1e-15, atol=1e-15)
    assert ints(out["arrs"]).tolist() == [3, 1, len(jc) + 1, len(jc), len(ir), len(ir)]        # Arr = {cp, jc, ir | numx}
    assert out["roundtrip"] == ["1"] and out["tolerant"] == ["1"] and out["different"] == ["0"]
    ccp, cir, cnum = O.to_csc(m, n, Iu, Ju, Vu)
    assert ints(out["csc_ess"]).tolist() == [len(Iu), m, n]                  # CSC has 3 essentials, SpCCols.cpp:46
    assert np.array_equal(ints(out["csc_jc"]), ccp) and np.array_equal(ints(out["csc_ir"]), cir)
    assert np.allclose(np.array(out["csc_num"], float), cnum, rtol=1e-15, atol=1e-15)
    assert ints(out["tr_ess"]).tolist() == [len(Iu), n, m, len(np.unique(Iu))]
    assert ints(out["tuples_back"]).tolist() == [len(Iu), int(ir[0]), int(jc[0])]


def test_matrix_market_expansion(tmp_path):
    # symmetric real file: transpose of every off-diagonal entry is added; duplicates merged with maximum (the drivers' BinOp)
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% comment\n5 5 5\n1 1 2.5\n3 1 1.25\n5 2 -4\n5 2 7\n4 4 1\n")
    out = run("mm", p)
    m, n, I, J, V = O.read_mm(str(p))
    assert out["dims"] == ["5", "5", "8"]                                     # 5 stored + 3 mirrored
    order = np.lexsort((I, J))
    assert np.array_equal(ints(out["rows"]), I[order]) and np.array_equal(ints(out["cols"]), J[order])
    assert np.allclose(np.array(out["vals"], float), V[order])
    q = tmp_path / "p.mtx"
    q.write_text("%%MatrixMarket matrix coordinate pattern general\n3 4 3\n1 4\n2 2\n3 1\n")
    out = run("mm", q)
    assert out["dims"] == ["3", "4", "3"] and out["vals"] == ["1", "1", "1"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "small.npz"))            # sevenvertex via the real ParallelReadMM
    r = tmp_path / "seven.mtx"
    with open(r, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"7 7 {len(g['seven_I'])}\n")
        for i, j, v in zip(g["seven_I"], g["seven_J"], g["seven_V"]):
            f.write(f"{i + 1} {j + 1} {float(v)!r}\n")
    out = run("mm", r)
    order = np.lexsort((g["seven_I"], g["seven_J"]))
    assert np.array_equal(ints(out["rows"]), g["seven_I"][order]) and np.allclose(np.array(out["vals"], float), g["seven_V"][order])


def test_semiring_functors_and_codes():
    out = run("semirings")
    INF = 2**31 - 1
    assert ints(out["mp"]).tolist() == [INF, 3, 10, INF, INF, INF]             # id, min, inf_plus with saturation (Semirings.h:40-47)
    assert out["mp_axpy"] == ["5"]
    assert [float(x) for x in out["pt"]] == [0.0, 3.75, 3.0]
    assert ints(out["ptb"]).tolist() == [0, 9, 0]                              # static_cast<T>(bool) * x
    assert ints(out["bb"]).tolist() == [0, 1, 0, 0, 1]                         # bool + bool = OR, bool * bool = AND
    assert ints(out["sm"]).tolist() == [-1, -1, 42, -5]                        # id -1, max, multiply returns its 2nd argument
    assert ints(out["ops"]).tolist() == [0, 1, 2, 3, 0, 0]                     # ABI opcodes; MinPlus<bool,bool> unsupported
    assert ints(out["codes"]).tolist() == [3001, 3002, 3003, 3004, 3005, 3007]  # SpDefs.h:72-78


@pytest.mark.parametrize("args", [(3, 3, 10, 10, 9, 9), (3, 3, 10, 10, 2, 3), (4, 4, 2, 2, 1, 0), (2, 4, 8361, 16, 8360, 15), (2, 2, 8361, 8361, 4180, 4179)])
def test_owner_rule_matches_oracle(args):
    pr, pc, m, n, r, c = args
    out = run("owner", *args)
    assert tuple(ints(out["owner"]).tolist()) == O.owner(m, n, pr, pc, r, c)
    assert tuple(ints(out["lastblock"]).tolist()) == O.block_range(m, pr, pr - 1)


def test_unsupported_semiring_is_a_compile_time_error(tmp_path):
    # no CPU fallback by design: a semiring outside the ABI must not compile (static_assert in SpMM<SR>)
    src = tmp_path / "bad.cpp"
    src.write_text('#include "CombBLAS/CombBLAS.h"\n'
                   "using namespace combblas;\n"
                   "template <class T1, class T2> struct MaxTimesSRing { typedef T2 T_promote; static T2 id() { return 0; } };\n"
                   "int main() {\n"
                   "  std::shared_ptr<CommGrid> g;\n"
                   "  SpParMat<int64_t, double, SpDCCols<int64_t, double>> A(g);\n"
                   "  DenseParMat<int64_t, double> X(0.0, g, 1, 1);\n"
                   "  auto Y = SpMM<MaxTimesSRing<double, double>>(A, X);\n"
                   "  return 0;\n}\n")
    inc = [f"-I{os.path.join(ROOT, 'combblas-spmm-test_b200', 'include')}", f"-I{os.path.join(ROOT, 'include')}"]
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only"] + inc + [str(src)], capture_output=True, text=True)
    assert r.returncode != 0 and "not implemented by the B200 SpMM" in r.stderr
    ok = tmp_path / "ok.cpp"
    ok.write_text(src.read_text().replace("MaxTimesSRing<double, double>>(A, X)", "MinPlusSRing<double, double>>(A, X)"))
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-Wno-int-in-bool-context"] + inc + [str(ok)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_hub_selection_rule():
    # cb_hub_select_host (csrc/cb_hub.cu): most frequent columns first, ties by ascending column, columns with < 2 nonzeros never
    import cbb200_loader
    cb = cbb200_loader.load_package()
    counts = np.array([5, 0, 9, 1, 9, 2, 2, 7], np.int32)
    cols, cum = cb.capi.hub_select(counts, 16)
    assert cols.tolist() == [2, 4, 7, 0, 5, 6] and cum.tolist() == [9, 18, 25, 30, 32, 34]
    cols, cum = cb.capi.hub_select(counts, 3)
    assert cols.tolist() == [2, 4, 7] and cum.tolist() == [9, 18, 25]
    assert cb.capi.hub_select(np.zeros(10, np.int32), 4)[0].size == 0
    rng = np.random.default_rng(0)
    counts = rng.zipf(1.3, 5000).clip(0, 100000).astype(np.int32)
    cols, cum = cb.capi.hub_select(counts, 300)
    order = np.lexsort((np.arange(len(counts)), -counts.astype(np.int64)))
    order = order[counts[order] >= 2][:300]
    assert np.array_equal(cols, order) and np.array_equal(cum, np.cumsum(counts[order].astype(np.int64)))
