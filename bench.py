#!/usr/bin/env python
"""bench.py - SpMM GFLOP/s + achieved HBM GB/s vs roofline (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5|...] [--impl ours|reference]

A "step" is one multiply Y = A (x).(+) X over the whole synthetic workload (operands resident in HBM);
`e2e` is the same multiply through the host-panel C-ABI call (X copied up from pinned host memory and Y
copied back inside the timed region).  `--impl reference` times the unmodified reference CPU
implementation (oracle/_ref, Mult_AnXBn_Synch with the dense panel stored as SpDCCols) on the host cores
for a bounded column sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> (generator, scale, edgefactor, symmetric, k, X dtype, A dtype (None = pattern), semiring name, x kind)
WORKLOADS = {
    "c2": dict(desc="R-MAT scale 20 ef 16 (symmetrised) x dense k=64, fp32 PlusTimes", gen="rmat", scale=20, ef=16, sym=True,
               k=64, xdt="f32", adt="f32", sr="plus_times", kind=0),
    "c3": dict(desc="Erdos-Renyi n=2^24 avg degree 16 (directed) x dense k=128, fp32 PlusTimes", gen="er", scale=24, ef=16,
               sym=False, k=128, xdt="f32", adt="f32", sr="plus_times", kind=0),
    "c4": dict(desc="R-MAT scale 24 ef 16 (symmetrised) x dense k=128, fp64 PlusTimes", gen="rmat", scale=24, ef=16, sym=True,
               k=128, xdt="f64", adt="f64", sr="plus_times", kind=0),
    "c5": dict(desc="R-MAT scale 22 ef 16 (symmetrised) x dense k=32, int32 MinPlus", gen="rmat", scale=22, ef=16, sym=True,
               k=32, xdt="i32", adt="i32", sr="min_plus", kind=1),
    "c5b": dict(desc="R-MAT scale 22 ef 16 (symmetrised, pattern) x dense k=32, boolean OR-AND", gen="rmat", scale=22, ef=16,
                sym=True, k=32, xdt="u8", adt=None, sr="or_and", kind=0),
    "s24f32": dict(desc="R-MAT scale 24 ef 16 (symmetrised) x dense k=128, fp32 PlusTimes", gen="rmat", scale=24, ef=16,
                   sym=True, k=128, xdt="f32", adt="f32", sr="plus_times", kind=0),
    "tiny": dict(desc="R-MAT scale 14 ef 16 (symmetrised) x dense k=64, fp32 PlusTimes", gen="rmat", scale=14, ef=16, sym=True,
                 k=64, xdt="f32", adt="f32", sr="plus_times", kind=0),
}
INITIATOR = {"rmat": (0.57, 0.19, 0.19, 0.05), "er": (0.25, 0.25, 0.25, 0.25)}
SEED_GRAPH, SEED_A, SEED_X = 0, 1, 42
NPDT = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64, "u8": np.uint8}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.rows, self.proc = device_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def grid_shape(n):
    """pr x pc for n ranks, pr <= pc, as square as possible: 1x1, 1x2, 2x2, 2x4."""
    pr = int(np.floor(np.sqrt(n)))
    while n % pr:
        pr -= 1
    return pr, n // pr


def block_range(total, nb, b):
    """Owner rule of the reference (SpParMat.cpp:5066-5096): floor division, last block takes the remainder."""
    per = total // nb
    start = b * per
    return start, (total - start if b == nb - 1 else per)


def alg_bytes(nnz, m, nzc, k, s_val, s_t):
    """Compulsory traffic of one multiply (BASELINE.md section 6 / SURVEY.md section 8d)."""
    return nnz * (4 + s_val) + (m + 1) * 4 + nzc * k * s_t + m * k * s_t


def run_reference(args, w):
    """--impl reference: the unmodified reference on the host cores, bounded column sample, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    sr = {"plus_times": O.PLUS_TIMES, "min_plus": O.MIN_PLUS, "or_and": O.OR_AND, "select_max": O.MAX_SEL2ND}[w["sr"]]
    # bounded sample: the same generator two scales down and a 4-column panel keeps the whole K+W run within
    # minutes at the reference's ~0.4 GFLOP/s; the metric (GFLOP/s) is size-normalised
    scale = min(w["scale"], args.ref_scale)
    kp = min(w["k"], args.ref_cols)
    n, I, J = O.rmat_matrix(scale, w["ef"], SEED_GRAPH, INITIATOR[w["gen"]], symmetric=w["sym"])
    V = None if w["adt"] is None else O.matrix_values(I, J, n, SEED_A, NPDT[w["adt"]])
    X = O.dense_operand(n, kp, SEED_X, NPDT[w["xdt"]], "x_minplus" if w["kind"] else "value")
    engine = "reference" if O.ref_available() else "port"
    cores = os.cpu_count()
    times = []
    how = "plain-C port, 1 thread"
    for it in range(args.warmup + args.steps):
        if engine == "reference":
            sec, how = O.ref_best_time(sr, n, n, I, J, V, X, cores)
        else:
            t0 = time.perf_counter()
            O.spmm(sr, n, n, I, J, V, X)
            sec = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(sec)
    t = float(np.mean(times))
    gflops = 2.0 * len(I) * kp / t / 1e9
    sample = f"{w['gen']} scale {scale} ({len(I)} nnz) x {kp} of {w['k']} columns, {w['xdt']} {w['sr']}, Mult_AnXBn_Synch, {how}"
    print(json.dumps({
        "impl": "reference", "metric": "spmm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": w["xdt"], "data": "synthetic", "config": {"workload": args.workload + ": " + w["desc"], "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": engine, "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-scale", type=int, default=18, help="largest generator scale the CPU reference sample uses")
    ap.add_argument("--ref-cols", type=int, default=16, help="panel columns of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--grid", default=None, help="process grid as PRxPC (default: as square as possible, pr <= pc)")
    ap.add_argument("--no-cache-a", action="store_true", help="re-broadcast the A parts every multiply like the reference does")
    ap.add_argument("--ring", type=int, default=0, help="opt into the ring-pipelined variant of the local multiply (K2R): depth 8; off by default")
    ap.add_argument("--hub", default=None, help="opt into the hub variant of the local multiply (K2H): CLUSTER[:SLAB_BYTES], e.g. 4 or 4:128; "
                                               "off by default (validated on the emulator only so far)")
    args = ap.parse_args()
    if args.workload is None:
        # One workload for every N so the 1/2/4/8 series is a strong-scaling series: BASELINE configs[2], the configuration
        # the metric is quoted on "at 1/2/4/8 B200" (ER n=2^24 x k=128 fp32).  It fits one GPU (2.1 GB A + 2 x 8.6 GB panels).
        # configs[1] (R-MAT s20 x 64) and the others run with --workload c2|c4|c5|...; their numbers are in profiles/.
        args.workload = "c3"
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import cbb200_loader
    cb = cbb200_loader.load_package()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    import torch
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        holder = [cb.capi.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(holder, src=0)
        uid = holder[0]
    pr, pc = grid_shape(world) if not args.grid else tuple(int(v) for v in args.grid.lower().split("x"))
    assert pr * pc == world, f"grid {pr}x{pc} does not match {world} ranks"
    ctx = cb.Context(local, rank, world, pr, pc, uid)
    if args.hub:
        hub_cluster, _, hub_slab = args.hub.partition(":")
        ctx.hub_config(1, int(hub_cluster), int(hub_slab or 0))
    if args.ring:
        ctx.ring_config(args.ring)
    if args.no_cache_a:
        ctx.summa_cache_a(False)

    sr = {"plus_times": cb.PLUS_TIMES, "min_plus": cb.MIN_PLUS, "or_and": cb.OR_AND, "select_max": cb.MAX_SEL2ND}[w["sr"]]
    xdt = NPDT[w["xdt"]]
    s_t = np.dtype(xdt).itemsize
    adt_code = cb.PATTERN if w["adt"] is None else cb.capi.CODE_OF[np.dtype(NPDT[w["adt"]])]
    s_val = 0 if w["adt"] is None else np.dtype(NPDT[w["adt"]]).itemsize
    N = 1 << w["scale"]
    k = w["k"]
    # this rank's blocks (reference distribution: SpParMat.cpp:5066-5096)
    r0, rl = block_range(N, pr, ctx.myprocrow)
    c0, cl = block_range(N, pc, ctx.myproccol)
    k0, kl = block_range(k, pc, ctx.myproccol)
    t_setup = time.time()
    tile = ctx.gen_rmat_tile(w["scale"], w["ef"], SEED_GRAPH, INITIATOR[w["gen"]], w["sym"], r0, rl, c0, cl, adt_code, SEED_A)
    X = ctx.dense(rl if world > 1 else N, kl, xdt)        # X tile: rows of block-row myprocrow, columns of block myproccol
    x_r0, x_rl = block_range(N, pr, ctx.myprocrow)
    X.generate(SEED_X, x_r0, k0, k, w["kind"])
    Y = ctx.dense(rl, kl, xdt)
    ctx.sync()
    t_setup = time.time() - t_setup

    def step():
        if world == 1:
            ctx.spmm_local(tile, X, Y, sr)
        else:
            ctx.spmm_summa(tile, X, Y, sr, N, N, k)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launches
    ctx.profile(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms_total = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    prof_ms, prof_n = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    # global counts + max over ranks
    stats = torch.tensor([ms_total, float(tile.nnz), float(tile.nzc), float(launches), prof_ms["spmm"], float(prof_n["spmm"])],
                         dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, nnz_total, launches_total = mx[0].item(), sm[1].item(), sm[3].item()
    else:
        nnz_total, launches_total = float(tile.nnz), float(launches)
    ms_step = ms_total / args.steps
    flops = 2.0 * nnz_total * k
    gflops = flops / (ms_step * 1e-3) / 1e9

    # roofline of the dominant kernel (K2, cb_spmm_kernel) on this rank: algorithmic bytes of the local multiply
    peak, peak_src = measured_peak_gbs()
    k2_ms = prof_ms["spmm"] / max(prof_n["spmm"], 1)
    stages = len(cb.capi.summa_plan(pr, pc, N)[1])
    summa_ms = ctx.summa_times() if world > 1 else None
    # this rank multiplies its whole block-row of A (nnz/pr nonzeros, all column blocks) by its k-block of X
    nzc_total = float(tile.nzc) if dist is None else sm[2].item()       # nonempty columns summed over all tiles
    nnz_rank = nnz_total / pr
    nzc_rank = nzc_total / pr                                            # nonempty columns of a block-row (grid-row average)
    b_alg_local = alg_bytes(nnz_rank, tile.m, nzc_rank, kl, s_val, s_t)
    k2_per_step = prof_n["spmm"] / args.steps if args.steps else 1       # launches per multiply (1, or one per X owner / stage)
    k2_ms_step = prof_ms["spmm"] / args.steps if args.steps else 0.0     # K2 time per multiply on this rank
    achieved = b_alg_local / (k2_ms_step * 1e-3) / 1e9 if k2_ms_step > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "cb_spmm_kernel", "kernel_ms": k2_ms, "kernel_launches_per_step": k2_per_step,
                "kernel_ms_per_step": k2_ms_step, "kernel_share_of_step": k2_ms_step / ms_step if ms_step else None,
                "peak_source": peak_src, "alg_bytes_per_step_this_rank": b_alg_local,
                "gather_bytes_per_step_this_rank": nnz_rank * (4 + s_val) + nnz_rank * kl * s_t + tile.m * kl * s_t,
                "whole_job_alg_gbs": alg_bytes(nnz_total, N, nzc_total / pr if world > 1 else tile.nzc, k, s_val, s_t) / (ms_step * 1e-3) / 1e9}
    traffic_file = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if world == 1 and os.path.exists(traffic_file):       # the capture is of the single-GPU launch
        try:
            roofline["traffic"] = json.load(open(traffic_file))["dram_bytes_per_launch"]
        except Exception:
            pass

    # e2e: the multiply through the public host-panel path, every step: X panel copied up from pinned host memory,
    # multiply, Y panel copied back.  1 GPU: cb_spmm_host.  N GPUs: each rank uploads its X tile, cb_spmm_summa, downloads
    # its Y tile (what SpMM<SR>(A, X) of the C++ layer does); timed with a barrier on both sides, max over ranks.
    tdt = getattr(torch, {"f32": "float32", "f64": "float64", "i32": "int32", "i64": "int64", "u8": "uint8"}[w["xdt"]])
    xr, xc = (N, k) if world == 1 else (rl, kl)
    Xh = torch.empty((xr, xc), dtype=tdt, pin_memory=True)
    Yh = torch.empty((rl, xc), dtype=tdt, pin_memory=True)
    X.download(Xh.numpy())

    def e2e_step():
        if world == 1:
            ctx.spmm_host(tile, Xh.numpy(), sr, Yh.numpy())
        else:
            X.upload(Xh.numpy())
            ctx.spmm_summa(tile, X, Y, sr, N, N, k)
            Y.download(Yh.numpy())

    e2e_step()                                                   # warm-up (allocates the workspace panels)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    te = (time.perf_counter() - t0) / args.e2e_steps
    if dist is not None:
        tt = torch.tensor([te], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = tt.item()
    e2e = {"value": flops / te / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(N * k * s_t), "d2h_bytes_per_step": int(N * k * s_t),
           "ms_per_step": te * 1e3, "resident": "A (uploaded once, as SpParMat construction does); X and Y cross PCIe every step"}
    checksum = float(np.asarray(Yh.numpy()[:1024], dtype=np.float64).sum())

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as O
            osr = {"plus_times": O.PLUS_TIMES, "min_plus": O.MIN_PLUS, "or_and": O.OR_AND, "select_max": O.MAX_SEL2ND}[w["sr"]]
            scale = min(w["scale"], args.ref_scale)
            kp = min(k, args.ref_cols)
            n_s, I, J = O.rmat_matrix(scale, w["ef"], SEED_GRAPH, INITIATOR[w["gen"]], symmetric=w["sym"])
            V = None if w["adt"] is None else O.matrix_values(I, J, n_s, SEED_A, NPDT[w["adt"]])
            Xs = O.dense_operand(n_s, kp, SEED_X, xdt, "x_minplus" if w["kind"] else "value")
            cores = os.cpu_count()
            if O.ref_available():
                sec, how = O.ref_best_time(osr, n_s, n_s, I, J, V, Xs, cores, reps=2)
                kind = "reference"
            else:
                t0 = time.perf_counter()
                O.spmm(osr, n_s, n_s, I, J, V, Xs)
                sec, kind, how = time.perf_counter() - t0, "port", "1 thread"
            cpu_baseline = {"value": 2.0 * len(I) * kp / sec / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
                            "sample": f"{w['gen']} scale {scale} ({len(I)} nnz) x {kp} of {k} columns, {w['xdt']} {w['sr']}, "
                                      f"Mult_AnXBn_Synch, {how}, {sec:.2f} s"}
        except Exception as ex:          # the checker is optional for the number; say why it is missing
            cpu_baseline = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}

    if rank == 0:
        out = {
            "metric": "spmm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": w["xdt"],
            "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "n": N, "nnz": int(nnz_total), "k": k, "semiring": w["sr"],
                       "grid": f"{pr}x{pc}", "stages": stages, "l2_policy": "inputs larger than L2 (A+X+Y per GPU >> 126 MB), no flush",
                       "generator": "counter-based Kronecker (csrc/cb_gen.cu), seed 0", "setup_s": round(t_setup, 3),
                       "chunks": tile.nchunks, "split_rows": tile.nsplit,
                       "a_parts_cached": (world > 1 and not args.no_cache_a), "local_kernel": " + ".join(([f"K2H hub variant {args.hub}"] if args.hub else []) + ([f"K2R ring depth {args.ring}"] if args.ring else [])) or "K2",
                       "summa_last_call_ms": None if summa_ms is None else {"stage_loop": summa_ms[0], "comm_stream_busy": summa_ms[1]}},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches_total), "clocks": clocks,
            "checksum_first_rows": checksum,
        }
        print(json.dumps(out))
    ctx.sync()
    for h in (tile, X, Y):
        h.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
