import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
dev = torch.device('cuda', 0)
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for D in (1024, 2048, 4096):
    A = torch.randn(D, D + 8, device=dev, dtype=torch.float64)
    H = ops.gemm(A, A, 'KC', 'KC', alpha=1.0 / D); H.diagonal().add_(1.0)
    fac = ops.potrf(H)
    eye = torch.eye(D, device=dev, dtype=torch.float64)
    v = torch.randn(D, device=dev, dtype=torch.float64)
    print('D=%d potrf %.3f ms | inverse (potrs on I) %.3f ms | potrs 1 rhs %.3f ms | clone %.3f ms | torch cholesky %.3f ms, cholesky_inverse %.3f ms' % (
        D, timed(lambda: ops.potrf(H)), timed(lambda: fac.solve(eye.clone(), overwrite=True)), timed(lambda: fac.solve(v)),
        timed(lambda: H.clone()), timed(lambda: torch.linalg.cholesky(H)), timed(lambda: torch.cholesky_inverse(fac.L.tril()))))
