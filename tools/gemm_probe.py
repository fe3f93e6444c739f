import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
dev = torch.device('cuda', 0)
peak = ops.fp64_peak_probe(0.2)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def report(name, M, N, K, t):
    tf = 2.0 * M * N * K / t / 1e9
    print('%-34s M=%d N=%d K=%d  %.3f ms  %.2f TF (%.1f%%)' % (name, M, N, K, t, tf, 100 * tf / peak))
# (a) KC x KC, few tiles, long K
M = N = 1536; K = 32768
A = torch.rand(M, K, device=dev, dtype=torch.float64); B = torch.rand(N, K, device=dev, dtype=torch.float64)
C = torch.empty(M, N, device=dev, dtype=torch.float64)
report('KCxKC long-K (144 tiles)', M, N, K, timeit(lambda: ops.gemm(A, B, 'KC', 'KC', out=C)))
At = A.T.contiguous(); Bt = B.T.contiguous()
report('KSxKS long-K (144 tiles)', M, N, K, timeit(lambda: ops.gemm(At, Bt, 'KS', 'KS', out=C)))
del A, B, At, Bt
# (b) apply-like shapes in both modes
M, N, K = 1024, 65536, 1024
H = torch.rand(M, K, device=dev, dtype=torch.float64); X = torch.rand(N, K, device=dev, dtype=torch.float64)
S = torch.empty(M, N, device=dev, dtype=torch.float64)
report('KCxKC apply-shape', M, N, K, timeit(lambda: ops.gemm(H, X, 'KC', 'KC', out=S)))
Ht = H.T.contiguous(); Xt = X.T.contiguous()
report('KSxKS apply-shape', M, N, K, timeit(lambda: ops.gemm(Ht, Xt, 'KS', 'KS', out=S)))
report('KCxKS apply-shape', M, N, K, timeit(lambda: ops.gemm(H, Xt, 'KC', 'KS', out=S)))
# (c) K sweep for KCxKC at fixed tiles
for K in (256, 512, 2048, 4096):
    H = torch.rand(M, K, device=dev, dtype=torch.float64); X = torch.rand(N, K, device=dev, dtype=torch.float64)
    report('KCxKC K-sweep', M, N, K, timeit(lambda: ops.gemm(H, X, 'KC', 'KC', out=S)))
