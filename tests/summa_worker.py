"""One rank of a SUMMA parity run, launched by torch.distributed.run (see tests/test_summa_*.py).

  --mode cpu : host logic only.  The C library's stage plan (cb_summa_plan) and the reference's block distribution
               drive a stage loop whose transport is gloo broadcasts on row / column process groups and whose local
               multiply is the CPU ORACLE (this is a test of the distributed plumbing, not of the product kernel).
  --mode gpu : the product.  Every rank builds its A tile / X tile on its GPU and calls cb_spmm_summa (NCCL).
Rank 0 gathers the Y tiles and compares with the single-rank oracle: bit-exact for integer semirings,
relative tolerance for floating point.  Exit code 0 = parity.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cbb200_loader  # noqa: E402
from oracle import oracle as O  # noqa: E402

cb = cbb200_loader.load_package()

CASES = {
    "minplus_i32": (O.MIN_PLUS, np.int32, np.int32, "x_minplus"),
    "pt_f64": (O.PLUS_TIMES, np.float64, np.float64, "value"),
    "pt_f32": (O.PLUS_TIMES, np.float32, np.float32, "value"),
    "pt_pat_i64": (O.PLUS_TIMES, None, np.int64, "value"),
    "selmax_i32": (O.MAX_SEL2ND, None, np.int32, "value"),
    "or_and": (O.OR_AND, None, np.uint8, "value"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cpu")
    ap.add_argument("--pr", type=int, required=True)
    ap.add_argument("--pc", type=int, required=True)
    ap.add_argument("--scale", type=int, default=9)
    ap.add_argument("--k", type=int, default=13)
    ap.add_argument("--cases", default="minplus_i32,pt_f64")
    ap.add_argument("--ragged", type=int, default=1, help="drop trailing rows/cols so nothing divides evenly")
    ap.add_argument("--golden", default="", help="npz of the reference's own multi-process results (tests/golden/grid_ref.npz): "
                    "also compare the gathered Y with the entry <case>_p<world>")
    ap.add_argument("--no-cache-a", action="store_true", help="re-broadcast the A parts on every multiply (reference behaviour)")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert a.pr * a.pc == world
    if a.mode == "gpu":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    pr, pc = a.pr, a.pc
    myrow, mycol = rank // pc, rank % pc                      # CommGrid.h:106-110
    row_groups = [dist.new_group([r * pc + c for c in range(pc)]) for r in range(pr)]
    col_groups = [dist.new_group([r * pc + c for r in range(pr)]) for c in range(pc)]

    # global operands, identical on every rank (counter-based generators)
    n, I, J = O.rmat_matrix(a.scale, 8, seed=3)
    m = n
    if a.ragged:
        m, n = n - 5, n - 3
        keep = (I < m) & (J < n)
        I, J = I[keep], J[keep]
    k = a.k
    ctx = None
    if a.mode == "gpu":
        holder = [cb.capi.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(holder, src=0)
        ctx = cb.Context(local, rank, world, pr, pc, holder[0])
        if a.no_cache_a:
            ctx.summa_cache_a(False)
    failures = 0
    if a.mode == "gpu":
        # distributed ingestion (cb_tile_from_distributed_coo): every rank hands in a slice of the triples - owned by anybody, with
        # duplicates across ranks - and must end up with exactly its block, duplicates merged by the rule asked for
        Vg = O.matrix_values(I, J, n, 5, np.float64)
        dupI, dupJ = I[::7], J[::7]                                           # every 7th entry a second time, with another value
        allI, allJ = np.concatenate([I, dupI]), np.concatenate([J, dupJ])
        allV = np.concatenate([Vg, Vg[::7] + 1.0])
        r0_, rl_ = cb.capi.block_range(m, pr, myrow)
        c0_, cl_ = cb.capi.block_range(n, pc, mycol)
        for dup_op, name in ((2, "max"), (1, "sum")):
            t = ctx.tile_from_distributed_coo(m, n, allI[rank::world], allJ[rank::world], allV[rank::world], dup_op)
            rp, cc, vv = t.to_csr(np.float64)
            t.free()
            want = {}
            for i_, j_, v_ in zip(allI.tolist(), allJ.tolist(), allV.tolist()):
                if r0_ <= i_ < r0_ + rl_ and c0_ <= j_ < c0_ + cl_:
                    k_ = (i_ - r0_, j_ - c0_)
                    want[k_] = v_ if k_ not in want else (max(want[k_], v_) if dup_op == 2 else want[k_] + v_)
            got = dict(zip(zip(np.repeat(np.arange(rl_), np.diff(rp)).tolist(), cc.tolist()), vv.tolist()))
            ok_ing = got == want
            flags = [None] * world if rank == 0 else None
            dist.gather_object(ok_ing, flags, dst=0)
            if rank == 0:
                print(f"[summa gpu {pr}x{pc}] distributed ingestion ({name} of duplicates): {'every block right' if all(flags) else 'WRONG'}", flush=True)
                failures += 0 if all(flags) else 1
        tp = ctx.tile_from_distributed_coo(m, n, allI[rank::world], allJ[rank::world], None, 0)
        ok_pat = tp.nnz == len(set(k_ for k_ in zip(allI.tolist(), allJ.tolist()) if r0_ <= k_[0] < r0_ + rl_ and c0_ <= k_[1] < c0_ + cl_))
        tp.free()
        flags = [None] * world if rank == 0 else None
        dist.gather_object(ok_pat, flags, dst=0)
        if rank == 0 and not all(flags):
            print(f"[summa gpu {pr}x{pc}] distributed ingestion (pattern): WRONG", flush=True)
            failures += 1
        # the same from TEXT (cb_tile_from_mm_text): every rank hands in the lines that start inside its byte range of one Matrix
        # Market data section; parsed, routed and merged on the GPUs
        lines = "".join(f"{i_ + 1} {j_ + 1}\t{v_!r}\n" for i_, j_, v_ in zip(allI.tolist(), allJ.tolist(), allV.tolist())).encode()
        lo_, hi_ = len(lines) * rank // world, len(lines) * (rank + 1) // world
        lo_ = lo_ if lo_ == 0 or lines[lo_ - 1:lo_] == b"\n" else lines.index(b"\n", lo_) + 1
        hi_ = hi_ if hi_ == len(lines) or lines[hi_ - 1:hi_] == b"\n" else lines.index(b"\n", hi_) + 1
        t = ctx.tile_from_mm_text(m, n, lines[lo_:hi_], val_dtype=cb.F64, dup_op=2)
        rp, cc, vv = t.to_csr(np.float64)
        t.free()
        want = {}
        for i_, j_, v_ in zip(allI.tolist(), allJ.tolist(), allV.tolist()):
            if r0_ <= i_ < r0_ + rl_ and c0_ <= j_ < c0_ + cl_:
                k_ = (i_ - r0_, j_ - c0_)
                want[k_] = v_ if k_ not in want else max(want[k_], v_)
        got = dict(zip(zip(np.repeat(np.arange(rl_), np.diff(rp)).tolist(), cc.tolist()), vv.tolist()))
        flags = [None] * world if rank == 0 else None
        dist.gather_object(got == want, flags, dst=0)
        if rank == 0:
            print(f"[summa gpu {pr}x{pc}] Matrix Market text shares parsed on the device: {'every block right' if all(flags) else 'WRONG'}", flush=True)
            failures += 0 if all(flags) else 1
    for case in a.cases.split(","):
        sr, adt, xdt, kind = CASES[case]
        V = None if adt is None else O.matrix_values(I, J, n, 7, adt)
        X = O.dense_operand(n, k, 9, xdt, kind)
        r0, rl = cb.capi.block_range(m, pr, myrow)
        c0, cl = cb.capi.block_range(n, pc, mycol)
        x0, xl = cb.capi.block_range(n, pr, myrow)
        k0, kl = cb.capi.block_range(k, pc, mycol)
        sel = (I >= r0) & (I < r0 + rl) & (J >= c0) & (J < c0 + cl)
        Il, Jl, Vl = I[sel] - r0, J[sel] - c0, (None if V is None else V[sel])
        Xl = np.ascontiguousarray(X[x0:x0 + xl, k0:k0 + kl])
        if a.mode == "gpu":
            tile = ctx.tile_from_coo(rl, cl, Il, Jl, Vl)
            Xd, Yd = ctx.dense_from(Xl) if kl else ctx.dense(xl, 0, xdt), ctx.dense(rl, kl, xdt)
            for _ in range(4):                                  # later calls run on the cached, merged block-row (both buffer parities)
                ctx.spmm_summa(tile, Xd, Yd, sr, m, n, k)
            Yl = Yd.download() if kl else np.zeros((rl, 0), xdt)
            # the host-panel entry (cb_spmm_summa_host: column slabs through H2D / stage loop / D2H) must give the same bits
            Yh = np.empty((rl, kl), Xl.dtype)
            ctx.spmm_summa_host(tile, Xl, Yh, sr, m, n, k)
            host_same = bool(np.array_equal(Yh, Yl))
            # sparse right-hand side (cb_spgemm_summa): keep about one entry of X in seven; this rank's block of C as triples
            keepB = (O.hash_values(np.arange(n * k, dtype=np.uint64), 77, np.int32).reshape(n, k) % 7) == 0
            Bl_i, Bl_j = np.nonzero(keepB[x0:x0 + xl, k0:k0 + kl])
            Bl_v = Xl[Bl_i, Bl_j] if kl else np.zeros(0, Xl.dtype)
            Bt = ctx.tile_from_coo(xl, kl, Bl_i.astype(np.int64), Bl_j.astype(np.int64), Bl_v)
            sp_i, sp_j, sp_v = ctx.spgemm_summa(tile, Bt, sr, Xl.dtype, m, n, k)
            sp2 = ctx.spgemm_summa(tile, Bt, sr, Xl.dtype, m, n, k)        # second product with the same A: its remote parts are resident now
            spgemm_again_same = all(np.array_equal(u, v) for u, v in zip((sp_i, sp_j, sp_v), sp2))
            Bt.free()
            for h in (tile, Xd, Yd):
                h.free()
        else:
            seg, a_owner, x_owner = cb.capi.summa_plan(pr, pc, n)
            Yl = None
            for s in range(len(a_owner)):
                lo, hi = seg[s], seg[s + 1]
                # A part: columns [lo,hi) of the tile of grid column a_owner[s], broadcast along my processor row
                if a_owner[s] == mycol:
                    ps = (Jl >= lo - c0) & (Jl < hi - c0)
                    obj = [(Il[ps], Jl[ps] - (lo - c0), None if Vl is None else Vl[ps])]
                else:
                    obj = [None]
                dist.broadcast_object_list(obj, src=myrow * pc + a_owner[s], group=row_groups[myrow])
                Ai, Aj, Av = obj[0]
                # X panel: rows [lo,hi) of the tile of grid row x_owner[s], broadcast along my processor column
                obj = [Xl[lo - x0:hi - x0] if x_owner[s] == myrow else None]
                dist.broadcast_object_list(obj, src=x_owner[s] * pc + mycol, group=col_groups[mycol])
                Xs = obj[0]
                if kl == 0:
                    continue
                Yl = O.spmm(sr, rl, hi - lo, Ai, Aj, Av, Xs, accum_into=Yl)
            if Yl is None:
                Yl = np.zeros((rl, kl), xdt)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((r0, k0, Yl), gathered, dst=0)
        if a.mode == "gpu":
            extra = [None] * world if rank == 0 else None
            dist.gather_object((host_same and spgemm_again_same, sp_i + r0, sp_j + k0, sp_v), extra, dst=0)
            if rank == 0:
                from tests.test_spgemm_gpu import host_spgemm
                keepB = (O.hash_values(np.arange(n * k, dtype=np.uint64), 77, np.int32).reshape(n, k) % 7) == 0
                BI, BJ = np.nonzero(keepB)
                BV = X[BI, BJ]
                Vh = None if V is None else V
                ri, rj, rv = host_spgemm(sr, m, I, J, Vh, BI.astype(np.int64), BJ.astype(np.int64), BV, X.dtype.type)
                gi = np.concatenate([e[1] for e in extra]); gj = np.concatenate([e[2] for e in extra]); gv = np.concatenate([e[3] for e in extra])
                o = np.lexsort((gi, gj))
                gi, gj, gv = gi[o], gj[o], gv[o]
                sp_ok = bool(np.array_equal(gi, ri) and np.array_equal(gj, rj))
                if sp_ok:
                    if np.issubdtype(rv.dtype, np.floating):
                        tolv = 1e-5 if rv.dtype == np.float32 else 1e-12
                        sp_ok = bool((np.abs(gv - rv) <= tolv * np.abs(rv)).all())
                    else:
                        sp_ok = bool(np.array_equal(gv, rv))
                hp_ok = all(e[0] for e in extra)
                print(f"[summa {a.mode} {pr}x{pc}] {case} host-panel path and repeated sparse product: {'same bits' if hp_ok else 'DIFFERENT'}; sparse rhs ({len(ri)} entries): {'matches' if sp_ok else 'WRONG'}", flush=True)
                failures += 0 if (hp_ok and sp_ok) else 1
        if rank == 0:
            Y = np.zeros((m, k), X.dtype if X.dtype != np.bool_ else np.uint8)
            for (rr, kk, y) in gathered:
                Y[rr:rr + y.shape[0], kk:kk + y.shape[1]] = y
            ref = O.spmm(sr, m, n, I, J, V, X)
            if np.issubdtype(ref.dtype, np.floating):
                tol = 1e-5 if ref.dtype == np.float32 else 1e-12
                ok = bool((np.abs(Y - ref) <= tol * np.maximum(np.abs(ref), 1e-300)).all())
            else:
                ok = bool(np.array_equal(Y, ref))
            if a.golden:
                # the unmodified reference run on pr x pc processes (tests/golden/make_golden_grid.py)
                gold = np.load(a.golden)[f"{case}_p{world}"]
                if np.issubdtype(ref.dtype, np.floating):
                    gok = bool((np.abs(Y - gold) <= tol * np.maximum(np.abs(gold.astype(np.float64)), 1e-300)).all())
                else:
                    gok = bool(np.array_equal(Y, gold))
                print(f"[summa {a.mode} {pr}x{pc}] {case} vs reference on {world} processes: {'same' if gok else 'DIFFERENT'}", flush=True)
                ok = ok and gok
            print(f"[summa {a.mode} {pr}x{pc}] {case}: {'ok' if ok else 'MISMATCH'}", flush=True)
            failures += 0 if ok else 1
    flag = torch.tensor([failures], device=f"cuda:{local}" if a.mode == "gpu" else "cpu")
    dist.broadcast(flag, src=0)
    if ctx is not None:
        ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
