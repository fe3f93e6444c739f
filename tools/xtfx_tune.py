#!/usr/bin/env python3
"""Timing of the HBM-bound passes over X (stats, HVP, dirderiv, gemv)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device('cuda', 0)
X = ops.synth_design(1, 0, N, D, dev)
theta = 0.3 * ops.synth_theta(1, D, dev)
y = (torch.rand(N, device=dev, dtype=torch.float64) < 0.5).double()
w = torch.ones(N, device=dev, dtype=torch.float64)
v = torch.rand(D, device=dev, dtype=torch.float64)
hbm = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else 6534.0
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
z, resid, s, grad = ops.glm_stats(X, theta, y, w)
gb = 8.0 * N * D / 1e9
for name, fn in [('stats+grad', lambda: ops.glm_stats(X, theta, y, w)),
                 ('stats only', lambda: ops.glm_stats(X, theta, y, w, want_grad=False)),
                 ('hvp', lambda: ops.glm_hvp(X, s, v)),
                 ('dirderiv q=1', lambda: ops.glm_dirderiv(X, z, v[None, :].contiguous(), w)),
                 ('dirderiv q=3', lambda: ops.glm_dirderiv(X, z, torch.stack([v, v, v]), w))]:
    t = timeit(fn)
    print('%-14s N=%d D=%d  %.3f ms  %.0f GB/s (%.1f%% of %.0f)' % (name, N, D, t, gb / t * 1e3, 100 * gb / t * 1e3 / hbm, hbm))
