"""bench.py's host-side arithmetic (no GPU): grid shapes, block ranges, the traffic model, the workload table."""
import json
import os
import subprocess
import sys

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grid_shapes_and_blocks():
    assert [bench.grid_shape(n) for n in (1, 2, 4, 8, 6, 9)] == [(1, 1), (1, 2), (2, 2), (2, 4), (2, 3), (3, 3)]
    assert bench.block_range(10, 3, 0) == (0, 3) and bench.block_range(10, 3, 2) == (6, 4)


def test_traffic_model_matches_baseline_md():
    # C3 = 19.4 GB, C2 ~ 0.69 GB, C5 (pattern A, i32 panel) ~ 1.38 GB
    c3 = bench.alg_bytes(268435327, 1 << 24, 1 << 24, 128, 4, 4)
    assert abs(c3 - 19.4e9) < 0.1e9
    c2 = bench.alg_bytes(31400128, 1 << 20, 646872, 64, 4, 4)
    assert abs(c2 - 0.69e9) < 0.01e9
    c5 = bench.alg_bytes(128305150, 1 << 22, 2396659, 32, 0, 4)
    assert abs(c5 - 1.38e9) < 0.03e9


def test_workloads_cover_the_baseline_configs():
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfg) == 5
    w = bench.WORKLOADS
    assert (w["c2"]["scale"], w["c2"]["k"], w["c2"]["xdt"]) == (20, 64, "f32")
    assert (w["c3"]["scale"], w["c3"]["k"], w["c3"]["gen"], w["c3"]["sym"]) == (24, 128, "er", False)
    assert (w["c4"]["scale"], w["c4"]["k"], w["c4"]["xdt"]) == (24, 128, "f64")
    assert (w["c5"]["scale"], w["c5"]["k"], w["c5"]["sr"]) == (22, 32, "min_plus") and w["c5b"]["sr"] == "or_and"


def test_reference_arm_runs_without_a_gpu():
    """--impl reference on a small workload: the same matrix and panel as our arm (leading rows x leading columns), steps and
    warm-up as asked, no GPU, none of the product's libraries."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1",
                        "--ref-budget", "20"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["steps"] == 2 and line["warmup"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "GFLOP/s"
    assert line["config"]["workload"].startswith("tiny: ") and "of the same panel" in line["config"]["sample"]
    # the sample is a power-of-two fraction of the rows of the SAME matrix (how large depends on how fast this machine is)
    import re
    rows = int(re.search(r"rows \[0, (\d+)\) of the 16384 x 16384 matrix", line["config"]["sample"]).group(1))
    assert rows >= 256 and 16384 % rows == 0


def test_host_evaluation_of_sampled_rows_matches_the_oracle():
    """bench.py's in-run parity check evaluates sampled rows with numpy; that evaluation itself is checked here against the
    oracle (whole product) for every semiring the workloads use."""
    import numpy as np
    from oracle import oracle as O
    n, I, J = O.rmat_matrix(9, 8, seed=3)
    order = np.lexsort((J, I))
    I, J = I[order], J[order]
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(I, minlength=n), out=rowptr[1:])
    rows = np.array([0, 5, 17, int(np.argmax(np.diff(rowptr))), n - 1])
    off = np.concatenate([[0], np.cumsum(np.diff(rowptr)[rows])])
    sel = np.concatenate([np.arange(rowptr[r], rowptr[r + 1]) for r in rows])
    for name, sr, adt, xdt, kind in (("c2", O.PLUS_TIMES, np.float32, np.float32, "value"), ("c4", O.PLUS_TIMES, np.float64, np.float64, "value"),
                                     ("c5", O.MIN_PLUS, np.int32, np.int32, "x_minplus"), ("c5b", O.OR_AND, None, np.uint8, "value")):
        w = bench.WORKLOADS[name]
        V = None if adt is None else O.matrix_values(I, J, n, 1, adt)
        X = O.dense_operand(n, 8, 42, xdt, kind)
        ref = O.spmm(sr, n, n, I, J, V, X)
        cols = J[sel]
        ucols = np.unique(cols)
        got = bench.host_rows(w, off, cols, None if V is None else V[sel], X[ucols], ucols)
        for i, r in enumerate(rows):
            if got[i] is None:
                assert rowptr[r] == rowptr[r + 1] and (ref[r] == bench.semiring_identity(w)).all()
            elif name in ("c2", "c4"):
                assert np.allclose(got[i], ref[r].astype(np.float64), rtol=bench.TOL[w["xdt"]], atol=0)
            else:
                assert np.array_equal(got[i], ref[r])


def test_every_tool_script_compiles_and_shell_scripts_parse():
    # the scripts under tools/ are how the files under profiles/ were made; they only run on a GPU box, so at least keep them loadable
    import glob
    import py_compile
    import subprocess
    for f in sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))):
        py_compile.compile(f, doraise=True)
    for f in sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh"))):
        assert subprocess.run(["bash", "-n", f]).returncode == 0, f
