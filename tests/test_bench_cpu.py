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
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-scale", "10", "--ref-cols", "4"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "GFLOP/s"
