"""Peer-copy bandwidth between two GPUs of one process, alone and while both GPUs run an HBM-bound kernel."""
import time, torch
assert torch.cuda.device_count() >= 2
n = 1 << 30  # 1 GiB
a = torch.empty(n, dtype=torch.uint8, device="cuda:0"); b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
c = torch.empty(n, dtype=torch.uint8, device="cuda:1"); d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
big0 = torch.empty(1 << 28, dtype=torch.float32, device="cuda:0"); big1 = torch.empty(1 << 28, dtype=torch.float32, device="cuda:1")
idx0 = torch.randint(0, 1 << 21, (1 << 24,), device="cuda:0"); idx1 = torch.randint(0, 1 << 21, (1 << 24,), device="cuda:1")
tab0 = torch.randn(1 << 21, 64, device="cuda:0"); tab1 = torch.randn(1 << 21, 64, device="cuda:1")
s0 = torch.cuda.Stream(device=0); s1 = torch.cuda.Stream(device=1)
k0 = torch.cuda.Stream(device=0); k1 = torch.cuda.Stream(device=1)
def sync():
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
def copy_pair(bidir):
    with torch.cuda.stream(s0): b.copy_(a, non_blocking=True)
    if bidir:
        with torch.cuda.stream(s1): d.copy_(c, non_blocking=True)
def load():
    with torch.cuda.device(0), torch.cuda.stream(k0): x = tab0[idx0].sum()
    with torch.cuda.device(1), torch.cuda.stream(k1): y = tab1[idx1].sum()
for name, bidir, busy in [("one-way idle", False, False), ("two-way idle", True, False), ("one-way busy", False, True), ("two-way busy", True, True)]:
    for it in range(3):
        sync(); t = time.perf_counter()
        if busy:
            for _ in range(8): load()
        copy_pair(bidir)
        s0.synchronize(); s1.synchronize(); tc = time.perf_counter() - t
        sync(); tt = time.perf_counter() - t
    print(f"{name}: copy done after {tc*1e3:.2f} ms -> {n/tc/1e9:.0f} GB/s per direction (all work {tt*1e3:.2f} ms)")
