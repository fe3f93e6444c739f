import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
dev = torch.device('cuda', 0)
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = torch.randn(D, D + 8, device=dev, dtype=torch.float64)
H = ops.gemm(A, A, 'KC', 'KC', alpha=1.0 / D); H.diagonal().add_(1.0)
fac = ops.potrf(H)
v = torch.randn(D, device=dev, dtype=torch.float64)
for _ in range(2):
    x = fac.solve(v)
torch.cuda.synchronize()
