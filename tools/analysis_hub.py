#!/usr/bin/env python
"""Offline (CPU) analysis behind DESIGN.md section 4: how much of an R-MAT SpMM's gather traffic could be served on-SM
by 2D tiles of the degree-sorted matrix staged in shared memory.  Uses the same generator as the bench."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
row_bytes = int(sys.argv[2]) if len(sys.argv) > 2 else 256
t0 = time.time()
n, I, J = O.rmat_matrix(scale, 16, 0)
nnz = len(I)
print(f"R-MAT scale {scale}: n={n} nnz={nnz} ({time.time()-t0:.0f} s to generate)")
deg = np.bincount(I, minlength=n)
order = np.argsort(-deg, kind="stable")
rank = np.empty(n, np.int64); rank[order] = np.arange(n)
ri, rj = rank[I], rank[J]
print("\n| H (top columns/rows by degree) | nnz with col in top-H | nnz in top-H x top-H | density of that block |")
print("|---|---|---|---|")
for H in (1024, 4096, 8192, 16384, 65536, 262144):
    c = (rj < H).mean(); b = ((ri < H) & (rj < H)).mean()
    print(f"| {H} | {100*c:.1f} % | {100*b:.1f} % | {b*nnz/H/H:.4f} |")
# 2D tiles of the degree-sorted matrix: X column block of C rows + Y row block of R rows live in shared memory
smem = 200 * 1024
print(f"\nTiles with X block C rows + Y block R rows of {row_bytes} B in ~200 KB of shared memory; a tile pays (C + 2R) row transfers,")
print("the plain gather pays one per nonzero: profitable iff nnz_tile > C + 2R.\n")
print("| R x C | tiles considered (top 64K x 64K corner) | nnz in profitable tiles | row transfers saved | predicted gather traffic |")
print("|---|---|---|---|---|")
corner = 65536
m = (ri < corner) & (rj < corner)
ci, cj = ri[m], rj[m]
for R, C in ((128, 512), (256, 512), (256, 256), (512, 256)):
    if (R + C) * row_bytes > smem:
        continue
    tid = (ci // R) * (corner // C) + (cj // C)
    cnt = np.bincount(tid, minlength=(corner // R) * (corner // C))
    prof = cnt > (C + 2 * R)
    nnz_prof = cnt[prof].sum()
    saved = nnz_prof - prof.sum() * (C + 2 * R)
    print(f"| {R} x {C} | {len(cnt)} | {100*nnz_prof/nnz:.1f} % of all nnz in {prof.sum()} tiles | {100*saved/nnz:.1f} % of all gathers | {100*(1-saved/nnz):.1f} % of today's |")
# per-SM window reuse: how many distinct columns does a window of W consecutive nonzeros (row-major order) touch
o = np.lexsort((J, I)); Js = J[o]
print("\n| window of consecutive nonzeros | distinct columns / nonzeros |")
print("|---|---|")
for W in (2048, 8192, 32768):
    nb = len(Js) // W
    pick = np.linspace(0, nb - 1, 40).astype(int)
    u = np.mean([len(np.unique(Js[b*W:(b+1)*W])) / W for b in pick])
    print(f"| {W} | {u:.3f} |")
