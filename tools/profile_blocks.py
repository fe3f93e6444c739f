#!/usr/bin/env python3
"""Launch the block-arrow (config 3) and dense-Cholesky (config 4) kernels once each so that ncu can capture them by
name: block_potrf / block_trsm / block_solve / tall_* / gmm_blocks (csrc/blockchol.cu), chol_diag_kernel and the
panel / trailing-update / substitution GEMMs (csrc/chol.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import vittles_b200 as vt  # noqa: E402
from vittles_b200 import ops  # noqa: E402
from vittles_b200.block_solver import BlockArrowSolver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dev = torch.device('cuda', 0)
K, d = 20, 16
M, Dg = K - 1, K * d
g = torch.Generator(device=dev).manual_seed(3)
centers = 0.35 * torch.randn(K, d, device=dev, dtype=torch.float64, generator=g)
lab = torch.randint(0, K, (N,), device=dev, generator=g)
X = centers[lab] + torch.randn(N, d, device=dev, dtype=torch.float64, generator=g)
obj = vt.objectives.GMMVBObjective(X, K, prior_prec=0.5)
m = centers.clone()
for _ in range(5):
    c = 0.5 * ((X[:, None, :] - m[None]) ** 2).sum(-1) + np.log(K)
    r = torch.softmax(-c, dim=1)
    m = (r.T @ X) / (r.sum(0)[:, None] + 0.5)
c = 0.5 * ((X[:, None, :] - m[None]) ** 2).sum(-1) + np.log(K)
rho = (-(c - c[:, -1:]))[:, :M].contiguous()
x = torch.cat([m.reshape(-1), rho.reshape(-1)])
sa = torch.as_tensor(obj.sparsity_array(), dtype=torch.int64, device=dev)
h = obj.vt_block_hessian(x, sa, which='full')
solver = BlockArrowSolver(h, overwrite=True)
b = torch.randn(h.shape[0], 8, device=dev, dtype=torch.float64, generator=g)
x1 = solver.solve(b[:, 0].contiguous())
x8 = solver.solve(b)
D = 4096
A = torch.randn(D, D + 64, device=dev, dtype=torch.float64, generator=g)
H = ops.gemm(A, A, 'KC', 'KC', alpha=1.0 / D)
H.diagonal().add_(1.0)
fac = ops.potrf(H)
Xs = fac.solve(torch.randn(D, 2048, device=dev, dtype=torch.float64, generator=g))
torch.cuda.synchronize()
print('profile run done', float(x1[0]), float(x8[0, 0]), float(Xs[0, 0]))
