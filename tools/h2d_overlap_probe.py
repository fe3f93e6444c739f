"""Does an H2D copy overlap with compute?  Copy 16 GB while different kernels run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
dev = torch.device('cuda', 0)
N, D = 2_000_000, 1024
Xh = torch.empty((N, D), dtype=torch.float64, pin_memory=True); Xh.fill_(0.01)
Xd = torch.empty((N, D), dtype=torch.float64, device=dev)
Y = torch.randn(1_000_000, D, dtype=torch.float64, device=dev)
s = torch.ones(Y.shape[0], dtype=torch.float64, device=dev)
A = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
big = torch.empty(1 << 30, dtype=torch.float32, device=dev)
gb = N * D * 8 / 1e9
cs = torch.cuda.Stream(device=dev)
ws = torch.cuda.Stream(device=dev)
def copy():
    with torch.cuda.stream(cs):
        Xd.copy_(Xh, non_blocking=True)
def timed(label, fn, work=None):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print('%-50s %.1f ms' % (label, dt * 1e3), flush=True)
    return dt
def n_syrk(k):
    for _ in range(k): ops.syrk_weighted(Y, s)
def n_mm(k):
    for _ in range(k): torch.mm(A, A)
def n_fill(k):
    for _ in range(k): big.fill_(1.0)
def n_sleep(k):
    torch.cuda._sleep(int(k))
timed('copy alone', copy)
for name, fn, k in [('syrk_weighted x8', n_syrk, 8), ('cublas dgemm 8192^3 x9', n_mm, 9), ('fill 4 GB x400', n_fill, 400),
                    ('sleep kernel', n_sleep, 5e8)]:
    t_k = timed(name + ' alone', lambda: fn(k))
    timed('copy || ' + name, lambda: (copy(), fn(k)))
    def on_side():
        copy()
        with torch.cuda.stream(ws):
            fn(k)
    timed('copy || ' + name + ' (compute on a side stream)', on_side)
