import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vittles_b200 as vt
from vittles_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
D = 1024
dev = torch.device('cuda', 0)
X = ops.synth_design(1, 0, N, D, dev)
y = (torch.rand(N, device=dev, dtype=torch.float64) < 0.5).double()
theta = torch.zeros(D, device=dev, dtype=torch.float64)
w = torch.ones(N, device=dev, dtype=torch.float64)
Xh = torch.empty((N, D), dtype=torch.float64, pin_memory=True); Xh.copy_(X)
yh = y.cpu().pin_memory(); th = theta.cpu().pin_memory(); wh = w.cpu().pin_memory()
def sync(): torch.cuda.synchronize()
def timed(fn, reps=3):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    sync(); return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    Xd = Xh.to(dev, non_blocking=True); return Xd
def resident():
    o = vt.objectives.GLMObjective(X, y); s = vt.HyperparameterSensitivityLinearApproximation(o, theta, w); return s
def streamed():
    o = vt.objectives.GLMObjective(Xh, yh)
    s = vt.HyperparameterSensitivityLinearApproximation(o, th, wh); return s
def streamed_stats_only():
    o = vt.objectives.GLMObjective(Xh, yh); st, H = o.vt_stats_and_hessian(theta, w); return H
print('h2d only      %.1f ms  (%.1f GB/s)' % (timed(h2d), 0), flush=True)
t = timed(h2d); print('h2d only      %.1f ms  (%.1f GB/s)' % (t, 8.0 * N * D / t / 1e6))
print('resident      %.1f ms' % timed(resident))
del X
print('streamed H    %.1f ms' % timed(streamed_stats_only))
print('streamed full %.1f ms' % timed(streamed))
sync()
t0 = time.perf_counter(); o = vt.objectives.GLMObjective(Xh, yh); t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
print('constructor host time %.1f ms, until copies done %.1f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
print('slice pinned?', Xh[100:200].is_pinned())
o = vt.objectives.GLMObjective(Xh, yh)
t0 = time.perf_counter(); st, H = o.vt_stats_and_hessian(theta, w); t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
print('sweep host time %.1f ms, until done %.1f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
X = Xh.to(dev)
chunk = X[:125000]
s = torch.rand(125000, device=dev, dtype=torch.float64)
Hc = torch.empty(D, D, device=dev, dtype=torch.float64)
print('syrk on one 125k chunk %.2f ms; on 2M rows %.2f ms' % (timed(lambda: ops.syrk_weighted(chunk, s, out=Hc), 10), timed(lambda: ops.syrk_weighted(X, torch.ones(N, device=dev, dtype=torch.float64), out=Hc), 3)))
