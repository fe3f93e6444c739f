"""H2D bandwidth at the e2e size (82 GB pinned): alone, chunked, and under a concurrent SYRK."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
dev = torch.device('cuda', 0)
N, D = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, 1024
t0 = time.perf_counter()
Xh = torch.empty((N, D), dtype=torch.float64, pin_memory=True)
Xh.fill_(0.01)
print('pinned alloc + fill %.1f s' % (time.perf_counter() - t0), flush=True)
Xd = torch.empty((N, D), dtype=torch.float64, device=dev)
gb = N * D * 8 / 1e9
cs = torch.cuda.Stream(device=dev)
def timed(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print('%-44s %.1f ms  %.1f GB/s' % (label, dt * 1e3, gb / dt), flush=True)
for _ in range(2):
    timed('whole tensor', lambda: Xd.copy_(Xh, non_blocking=True))
def chunks(n):
    with torch.cuda.stream(cs):
        for c in range(n):
            r0, r1 = N * c // n, N * (c + 1) // n
            Xd[r0:r1].copy_(Xh[r0:r1], non_blocking=True)
timed('16 chunks, side stream', lambda: chunks(16))
timed('64 chunks, side stream', lambda: chunks(64))
s = torch.ones(N // 16, dtype=torch.float64, device=dev)
def with_compute():
    chunks(16)
    for c in range(16):
        ops.syrk_weighted(Xd[:N // 16], s)
timed('16 chunks + 16 concurrent SYRKs (1/16 each)', with_compute)
