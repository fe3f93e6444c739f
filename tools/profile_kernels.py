#!/usr/bin/env python3
"""Launch each hot kernel a fixed number of times on a D=1024 shard so that ncu
can capture them by name (profiles/README.md has the command lines)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vittles_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda', 0)
X = ops.synth_design(1, 0, N, D, dev)
theta = 0.3 * ops.synth_theta(1, D, dev)
y = (torch.rand(N, device=dev, dtype=torch.float64) < 0.5).double()
w = torch.ones(N, device=dev, dtype=torch.float64)
hinv = torch.eye(D, device=dev, dtype=torch.float64) + 0.01 * torch.rand(D, D, device=dev, dtype=torch.float64)
v = torch.rand(D, device=dev, dtype=torch.float64)
dirs = torch.rand(2, D, device=dev, dtype=torch.float64)
for _ in range(REPS):
    z, resid, s, grad = ops.glm_stats(X, theta, y, w)
for _ in range(REPS):
    H = ops.syrk_weighted(X, s)
S = None
for _ in range(REPS):
    S = ops.ij_apply(hinv, X, resid, out=S)
for _ in range(REPS):
    q = ops.glm_hvp(X, s, v)
for _ in range(REPS):
    dd = ops.glm_dirderiv(X, z, dirs, w)
delta = torch.rand(N, device=dev, dtype=torch.float64)
for _ in range(REPS):
    p = ops.gemv(S, delta)
# optional TF32 path (tcgen05): two conversion chunks of the apply and of the Hessian assembly
NT = min(N, 2 * 37888)
S32 = H32 = None
for prec in ('tf32', 'tf32x3'):
    S32 = ops.ij_apply(hinv, X[:NT], resid[:NT], precision=prec)
    H32 = ops.syrk_weighted(X[:NT], s[:NT].abs(), precision=prec)
# FP64-grade INT8 engine (tcgen05.mma.kind::i8): two chunks of the apply and of the Hessian assembly
NO = min(N, 2 * 32768)
So = ops.ij_apply(hinv, X[:NO], resid[:NO], precision='f64_ozaki')
Ho = ops.syrk_weighted(X[:NO], s[:NO].abs(), precision='f64_ozaki')
torch.cuda.synchronize()
print('profile run done', float(H[0, 0]), float(q[0]), float(dd[0]), float(p[0]))
