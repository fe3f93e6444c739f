"""One factorisation (and one solve) at the given size, for a launch list under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/chol_once.py 4096 2048"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vittles_b200 import ops  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(D)
A = torch.randn(D, D + 64, device=dev, dtype=torch.float64, generator=g)
H = ops.gemm(A, A, 'KC', 'KC', alpha=1.0 / D)
H.diagonal().add_(1.0)
B = torch.randn(D, K, device=dev, dtype=torch.float64, generator=g)
torch.cuda.synchronize()
fac = ops.potrf(H.clone(), overwrite=True)
X = fac.solve(B)
torch.cuda.synchronize()
print('residual', float((H @ X - B).abs().max() / B.abs().max()))
