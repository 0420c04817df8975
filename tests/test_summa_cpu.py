"""Host-side logic of the multi-GPU path, on CPU: the stage plan exported by the C library and the 2D block
distribution, exercised in real multi-process runs over gloo (world sizes 2 and 4)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import cbb200_loader
from oracle import oracle as O

cb = cbb200_loader.load_package()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_retrying_ports(make_cmd, **kw):
    """subprocess.run(make_cmd(port)) with a fresh free port, again if the launcher found the port taken after all"""
    for attempt in range(4):
        r = subprocess.run(make_cmd(free_port()), **kw)
        if r.returncode == 0 or "EADDRINUSE" not in (r.stderr or ""):
            return r
    return r


def torchrun(n, args, timeout=600, extra_env=None):
    env = dict(os.environ, OMP_NUM_THREADS="2", **(extra_env or {}))
    for attempt in range(4):        # a port that was free a moment ago can be taken by the time the launcher binds it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", str(free_port()), os.path.join(ROOT, "tests", "summa_worker.py")] + args
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            return r
    return r


@pytest.mark.parametrize("pr,pc,gn", [(1, 1, 10), (2, 2, 10), (2, 2, 11), (2, 4, 16), (2, 4, 1000003), (3, 2, 10), (4, 1, 7), (1, 4, 7),
                                      (4, 4, 2), (2, 4, 1 << 24), (8, 1, 100), (3, 5, 77)])
def test_stage_plan_covers_inner_dimension_once(pr, pc, gn):
    seg, a_owner, x_owner = cb.capi.summa_plan(pr, pc, gn)
    assert seg[0] == 0 and seg[-1] == gn and all(b > a for a, b in zip(seg, seg[1:]))
    assert len(a_owner) <= pr + pc - 1
    for s in range(len(a_owner)):
        lo, hi = seg[s], seg[s + 1]
        c0, cl = cb.capi.block_range(gn, pc, a_owner[s])
        x0, xl = cb.capi.block_range(gn, pr, x_owner[s])
        assert c0 <= lo and hi <= c0 + cl          # the stage lies inside ONE column block of A ...
        assert x0 <= lo and hi <= x0 + xl          # ... and inside ONE row block of X
        assert O.owner(gn, gn, pr, pc, lo, lo)[0] == x_owner[s] * pc + a_owner[s]     # reference Owner() agrees
    if pr == pc and gn % pr == 0:
        assert len(a_owner) == pc and a_owner == list(range(pc)) and x_owner == list(range(pr))   # stages = grcols


def test_block_range_matches_reference_owner_rule():
    for total, nb in [(8361, 2), (8361, 4), (7, 8), (1 << 20, 3)]:
        for b in range(nb):
            assert cb.capi.block_range(total, nb, b) == O.block_range(total, nb, b)


@pytest.mark.parametrize("pr,pc", [(1, 2), (2, 1)])
def test_summa_host_logic_world2_gloo(pr, pc):
    r = torchrun(2, ["--mode", "cpu", "--pr", str(pr), "--pc", str(pc), "--cases", "minplus_i32,pt_f64,selmax_i32"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(": ok") == 3


def test_summa_host_logic_world4_gloo():
    r = torchrun(4, ["--mode", "cpu", "--pr", "2", "--pc", "2", "--cases", "minplus_i32,pt_pat_i64,or_and", "--k", "5"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(": ok") == 3
