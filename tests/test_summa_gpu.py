"""The real multi-GPU SUMMA (cb_spmm_summa over NCCL) against the single-rank oracle.  Needs >= 2 GPUs on the box;
run with `gpurun --gpus N`.  One process per GPU, launched with torch.distributed.run."""
import pytest
import torch

from tests.test_summa_cpu import torchrun

pytestmark = pytest.mark.gpu
ALL = "minplus_i32,pt_f64,pt_f32,pt_pat_i64,selmax_i32,or_and"


def need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("pr,pc", [(1, 2), (2, 1)])
def test_summa_2gpu(pr, pc):
    need(2)
    r = torchrun(2, ["--mode", "gpu", "--pr", str(pr), "--pc", str(pc), "--cases", ALL, "--scale", "11", "--k", "24"])
    assert r.returncode == 0 and r.stdout.count(": ok") == 6, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("variant", ["nccl_transport", "no_cache", "nccl_no_cache", "no_merge"])
@pytest.mark.parametrize("pr,pc", [(1, 2), (2, 1)])
def test_summa_2gpu_alternate_paths(pr, pc, variant):
    # the NCCL broadcast transport, the reference-style re-broadcast of A on every call, and the unfused stage loop
    need(2)
    env = {}
    if "nccl" in variant:
        env["CB_SUMMA_TRANSPORT"] = "nccl"
    if variant == "no_merge":
        env["CB_SUMMA_MERGE"] = "0"
    args = ["--mode", "gpu", "--pr", str(pr), "--pc", str(pc), "--cases", "minplus_i32,pt_f64,or_and", "--scale", "11", "--k", "24"]
    if "no_cache" in variant:
        args.append("--no-cache-a")
    r = torchrun(2, args, extra_env=env)
    assert r.returncode == 0 and r.stdout.count(": ok") == 3, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("pr,pc", [(2, 2), (1, 4), (4, 1)])
def test_summa_4gpu(pr, pc):
    need(4)
    r = torchrun(4, ["--mode", "gpu", "--pr", str(pr), "--pc", str(pc), "--cases", ALL, "--scale", "12", "--k", "33"])
    assert r.returncode == 0 and r.stdout.count(": ok") == 6, r.stdout[-3000:] + r.stderr[-3000:]


def test_summa_2x2_matches_reference_run_on_2x2_processes():
    # the product on a 2x2 GPU grid against the stored results of the unmodified reference on a 2x2 PROCESS grid
    # (tests/golden/grid_ref.npz, made by oracle/_ref/cbref_grid): same distribution, same stages
    need(4)
    import os
    from tests.golden.make_golden_grid import K, SCALE
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid_ref.npz")
    r = torchrun(4, ["--mode", "gpu", "--pr", "2", "--pc", "2", "--cases", ALL, "--scale", str(SCALE), "--k", str(K), "--golden", gold])
    assert r.returncode == 0 and r.stdout.count("processes: same") == 6 and r.stdout.count(": ok") == 6, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("pr,pc", [(2, 4), (4, 2)])
def test_summa_8gpu(pr, pc):
    need(8)
    r = torchrun(8, ["--mode", "gpu", "--pr", str(pr), "--pc", str(pc), "--cases", ALL, "--scale", "12", "--k", "64"])
    assert r.returncode == 0 and r.stdout.count(": ok") == 6, r.stdout[-3000:] + r.stderr[-3000:]
