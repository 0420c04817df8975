#!/usr/bin/env python
"""Offline (CPU) model of the persistent variants of DESIGN.md section 4b on the bench's own R-MAT inputs: for every
(cluster size, slab width, ring on/off) the planner of csrc/cb_hub.cu would choose, how many hub rows are resident, which
share of the row gathers they serve, and how that share splits into own-SM and neighbour-SM reads.
    python tools/hub_model.py <scale> <row_bytes> [smem_kb]          e.g. 20 256   (C2),  22 128  (C5),  24 512
Uses the same generator, hub selection rule (count >= 2, most frequent first) and arithmetic as the product."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
row_bytes = int(sys.argv[2]) if len(sys.argv) > 2 else 256
smem_kb = int(sys.argv[3]) if len(sys.argv) > 3 else 200
BT, RING_D = 1024, 8
t0 = time.time()
n, I, J = O.rmat_matrix(scale, 16, 0)
nnz = len(J)
counts = np.bincount(J, minlength=n)
order = np.lexsort((np.arange(n), -counts))
order = order[counts[order] >= 2][:65535]
cum = np.cumsum(counts[order])
print(f"R-MAT scale {scale}: n={n} nnz={nnz}, {len(order)} hub candidates ({time.time() - t0:.0f} s)\n")
print(f"panel rows of {row_bytes} B, {smem_kb} KB of shared memory per SM, ring = {BT} threads x {RING_D} x 16 B = {BT * RING_D * 16 // 1024} KB\n")
print("| slab B | passes over A | ring | cluster | resident hub rows | gathers served on SMs | own SM | neighbour SM | still through L2 |")
print("|---|---|---|---|---|---|---|---|---|")
for slab in (128, 256, 512):
    if slab > max(row_bytes, 128) and slab != 128:
        continue
    passes = -(-row_bytes // slab)
    for ring in (0, RING_D):
        ring_bytes = BT * ring * 16
        slots = (smem_kb * 1024 - ring_bytes) // slab
        for cs in (1, 2, 4, 8):
            nhub = int(min(len(order), slots * cs))
            cover = cum[nhub - 1] / nnz if nhub else 0.0
            print(f"| {slab} | {passes} | {'on' if ring else 'off'} | {cs} | {nhub} | {100 * cover:.1f} % | {100 * cover / cs:.1f} % | "
                  f"{100 * cover * (cs - 1) / cs:.1f} % | {100 * (1 - cover):.1f} % |")
