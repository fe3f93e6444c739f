import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vittles_b200 as vt
from vittles_b200 import ops, objectives
from vittles_b200._arrays import to_device
N, D = 2_000_000, 1024
dev = torch.device('cuda', 0)
Xh = torch.empty((N, D), dtype=torch.float64, pin_memory=True); Xh.fill_(0.01)
yh = torch.ones(N, dtype=torch.float64).pin_memory()
def now(): return time.perf_counter()
for rep in range(3):
    torch.cuda.synchronize()
    o = objectives.GLMObjective.__new__(objectives.GLMObjective)
    o._pending = []
    t0 = now(); Xd = o._stream_from_host(Xh, None, 16); t1 = now()
    y = to_device(yh, Xd.device).reshape(-1).contiguous(); t2 = now()
    torch.cuda.synchronize(); t3 = now()
    print('rep %d: stream_from_host host %.1f ms | y to_device %.1f ms | sync %.1f ms' % (rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
    del Xd, o
