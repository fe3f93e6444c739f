#!/usr/bin/env python3
"""Quick timing of the two FP64 tensor-core hot kernels on a D=1024 shard."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vittles_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device('cuda', 0)
X = ops.synth_design(1, 0, N, D, dev)
s = torch.rand(N, device=dev, dtype=torch.float64)
r = torch.rand(N, device=dev, dtype=torch.float64)
hinv = torch.eye(D, device=dev, dtype=torch.float64) + 0.01 * torch.rand(D, D, device=dev, dtype=torch.float64)
peak = ops.fp64_peak_probe(0.2)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


S = torch.empty((D, N), device=dev, dtype=torch.float64)
H = torch.empty((D, D), device=dev, dtype=torch.float64)
t_apply = timeit(lambda: ops.ij_apply(hinv, X, r, out=S))
t_syrk = timeit(lambda: ops.syrk_weighted(X, s, out=H))
St = torch.empty((N, D), device=dev, dtype=torch.float64)
t_applyT = timeit(lambda: ops.gemm(X, hinv, 'KC', 'KC', alpha=-1.0, rowscale=r, out=St))
print('apply transposed-out (N,D): %.3f ms %.2f TF (%.1f%%)' % (t_applyT, 2.0 * D * D * N / t_applyT / 1e9, 100 * 2.0 * D * D * N / t_applyT / 1e9 / peak))
Ns = 16384
Ss = torch.empty((D, Ns), device=dev, dtype=torch.float64)
t_small = timeit(lambda: ops.ij_apply(hinv, X[:Ns], r[:Ns], out=Ss), reps=20)
print('apply small N=16384 (D,N): %.3f ms %.2f TF' % (t_small, 2.0 * D * D * Ns / t_small / 1e9))
ta = 2.0 * D * D * N / t_apply / 1e9
ts = float(D) * (D + 1) * N / t_syrk / 1e9
print('peak %.2f TF | apply %.3f ms %.2f TF (%.1f%%) | syrk %.3f ms %.2f TF alg (%.1f%%), %.2f TF executed' % (
    peak, t_apply, ta, 100 * ta / peak, t_syrk, ts, 100 * ts / peak, ts * (36 * 128 * 128 * 2) / (D * (D + 1.0))))
# correctness spot check against cuBLAS (torch) on a slice
n = 4096
ref = -(hinv @ (X[:n] * r[:n, None]).T)
err = float((S[:, :n] - ref).abs().max() / ref.abs().max())
Href = (X * s[:, None]).T @ X
errh = float((H - Href).abs().max() / Href.abs().max())
print('rel err apply %.2e  syrk %.2e' % (err, errh))
