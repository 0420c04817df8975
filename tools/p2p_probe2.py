"""Does a peer DMA copy make progress while the DRAM-bound SpMM kernel runs on both GPUs?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cbb200_loader
cb = cbb200_loader.load_package()
ER = (0.25, 0.25, 0.25, 0.25)
ctxs = [cb.Context(0), cb.Context(1)]
objs = []
for c in ctxs:
    t = c.gen_rmat_tile(22, 16, 0, ER, False, val_dtype=cb.F32)
    X = c.dense(1 << 22, 128, np.float32); X.generate(42); Y = c.dense(1 << 22, 128, np.float32)
    objs.append((t, X, Y))
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda:0"); b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
s0 = torch.cuda.Stream(device=0)
def sync():
    for c in ctxs: c.sync()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
def kernels(k):
    for _ in range(k):
        for c, (t, X, Y) in zip(ctxs, objs): c.spmm_local(t, X, Y, cb.PLUS_TIMES)
kernels(2); sync()
for name, nk in [("copy alone", 0), ("copy + 6 kernels", 6), ("kernels alone", -6)]:
    sync(); t0 = time.perf_counter()
    if nk: kernels(abs(nk))
    if nk >= 0:
        with torch.cuda.stream(s0): b.copy_(a, non_blocking=True)
        s0.synchronize()
    tc = time.perf_counter() - t0
    sync(); tt = time.perf_counter() - t0
    print(f"{name}: copy finished at {tc*1e3:.2f} ms ({n/tc/1e9:.0f} GB/s), everything at {tt*1e3:.2f} ms", flush=True)
