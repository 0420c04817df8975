#!/usr/bin/env python
"""The reference's only published tall-skinny runs (ReleaseTests/SCALE26RECT8192/cores4096_density*: `MultTime rmat26.txt
fringe_scale26_rect8192_sparse<d>` = Mult_AnXBn_Synch<PlusTimesSRing<double,double>> of an R-MAT scale-26 matrix with a SPARSE
2^26 x 8192 operand holding d nonzeros per column, 1.76 - 3.97 s per multiply on 4096 MPI tasks for d = 1 .. 10 000) on this
library's device sparse x sparse product (cb_spgemm_summa: expand - stable sort - reduce-by-key, A parts and B tiles travelling
over NCCL every call like the reference's broadcasts).

usage:  python -m torch.distributed.run --nproc-per-node N tools/spgemm_rect.py --scale 26 --densities 1 10 100 1000 10000
        python tools/spgemm_rect.py --scale 24                       (one GPU)
One JSON line per density: seconds per multiply (wall clock around the collective call, barrier + device synchronisation on both
sides, max over ranks, after one warm-up call as MultTiming.cpp does), partial products, nnz(C), and the reference's published
time for that density at scale 26.  Inputs: this repository's R-MAT generator (symmetrised, fp64 values) and a hash-placed fringe
(column j holds the rows hash(j, t) mod n, t < d; duplicates merged) - the reference's input files are not part of its tree."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cbb200_loader  # noqa: E402
from bench import INITIATOR, block_range, grid_shape  # noqa: E402

PUBLISHED = {1: 1.760686, 10: 2.137648, 100: 2.504020, 1000: 3.842937, 10000: 3.965003, 100000: 17.138062}


def splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--cols", type=int, default=8192)
    ap.add_argument("--densities", type=int, nargs="*", default=[1, 10, 100, 1000])
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--grid", default="")
    ap.add_argument("--no-cache-a", action="store_true", help="re-send A's parts on every call, as the reference's broadcasts do")
    ap.add_argument("--directed", action="store_true", help="do not symmetrise A (the generator takes < 2^31 candidate edges: scale 26 needs this)")
    a = ap.parse_args()
    cb = cbb200_loader.load_package()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    dist, uid = None, None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        holder = [cb.capi.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(holder, src=0)
        uid = holder[0]
    pr, pc = grid_shape(world) if not a.grid else tuple(int(v) for v in a.grid.lower().split("x"))
    ctx = cb.Context(local, rank, world, pr, pc, uid)
    if a.no_cache_a:
        ctx.summa_cache_a(False)
    n, k = 1 << a.scale, a.cols
    r0, rl = block_range(n, pr, ctx.myprocrow)
    c0, cl = block_range(n, pc, ctx.myproccol)
    x0, xl = block_range(n, pr, ctx.myprocrow)             # B is distributed like any matrix: rows over pr, columns over pc
    k0, kl = block_range(k, pc, ctx.myproccol)
    t0 = time.time()
    A = ctx.gen_rmat_tile(a.scale, 16, 0, INITIATOR["rmat"], not a.directed, r0, rl, c0, cl, cb.F64, 1)
    ctx.sync()
    nnzA = A.nnz
    if dist is not None:
        tot = torch.tensor([nnzA], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        nnzA = int(tot.item())
    if rank == 0:
        print(json.dumps(dict(setup="A", scale=a.scale, symmetrised=not a.directed, nnz=nnzA, grid=f"{pr}x{pc}", seconds=round(time.time() - t0, 2))), flush=True)
    L = cb.capi.lib()
    for d in a.densities:
        # this rank's block of the fringe
        j = np.repeat(np.arange(k0, k0 + kl, dtype=np.uint64), d)
        t = np.tile(np.arange(d, dtype=np.uint64), kl)
        h = splitmix(j * np.uint64(1000003) + t + np.uint64(12345))
        rows = (h % np.uint64(n)).astype(np.int64)
        keep = (rows >= x0) & (rows < x0 + xl)
        bi, bj = rows[keep] - x0, j[keep].astype(np.int64) - k0
        key = np.unique(bj * xl + bi)                       # a column that drew a row twice keeps one entry
        bj, bi = key // xl, key % xl
        bv = (splitmix(key.astype(np.uint64) + np.uint64(99)) >> np.uint64(11)).astype(np.float64) / float(1 << 53) + 0.5
        B = ctx.tile_from_coo(xl, kl, bi, bj, bv)
        secs, info = [], None
        for it in range(a.iters + 1):                       # the first call warms up (MultTiming.cpp:83)
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            c = ctypes.c_void_p()
            cb.capi._check(L.cb_spgemm_summa(ctx.h, A.h, B.h, cb.PLUS_TIMES, cb.F64, n, n, k, ctypes.byref(c)), ctx.h)
            ctx.sync()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t1
            nnz, m_, k_, dt_ = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
            L.cb_coo_info(c, ctypes.byref(nnz), ctypes.byref(m_), ctypes.byref(k_), ctypes.byref(dt_))
            L.cb_coo_free(c)
            v = torch.tensor([dt, float(nnz.value), float(B.nnz)], dtype=torch.float64, device="cuda")
            if dist is not None:
                mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                sm = v.clone(); dist.all_reduce(sm)
                dt, nnzC, nnzB = float(mx[0]), int(sm[1]), int(sm[2])
            else:
                nnzC, nnzB = int(v[1]), int(v[2])
            if it > 0:
                secs.append(dt)
            info = (nnzC, nnzB)
        B.free()
        if rank == 0:
            print(json.dumps(dict(workload=f"R-MAT scale {a.scale} x sparse 2^{a.scale} x {k}, {d} nonzeros per column, fp64 PlusTimes", n_gpus=world,
                                  grid=f"{pr}x{pc}", a_parts_cached=not a.no_cache_a and pc > 1, seconds_per_multiply=round(float(np.mean(secs)), 5), runs=[round(s, 5) for s in secs],
                                  nnz_A=nnzA, nnz_B=info[1], nnz_C=info[0],
                                  reference_published_seconds_scale26_4096_tasks=PUBLISHED.get(d) if a.scale == 26 else None)), flush=True)
    A.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
