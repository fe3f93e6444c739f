import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
dev = torch.device('cuda', 0)
M = N = 1536; K = 8192
A = torch.rand(M, K, device=dev, dtype=torch.float64); B = torch.rand(N, K, device=dev, dtype=torch.float64)
C = torch.empty(M, N, device=dev, dtype=torch.float64)
for _ in range(2):
    ops.gemm(A, B, 'KC', 'KC', out=C)
torch.cuda.synchronize()
