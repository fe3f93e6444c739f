set -x
python tools/chol_probe.py
VT_CHOL_LOOKAHEAD=0 python tools/chol_probe.py
python -m pytest tests/test_gpu_solver.py tests/test_gpu_lrcov.py tests/test_gpu_ij.py -q 2>&1 | tail -3
for i in 1 2 3; do python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from vittles_b200 import ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(1)
worst = 0.0
for D in (512, 640, 1024, 1500, 2048, 4096, 4000):
    A = torch.randn(D, D + 32, device=dev, dtype=torch.float64, generator=g)
    H = A @ A.T / D + torch.eye(D, device=dev, dtype=torch.float64)
    for rep in range(4):
        L = torch.tril(ops.potrf(H).L)
        err = float((L @ L.T - H).abs().max() / H.abs().max())
        worst = max(worst, err)
print('worst reconstruction error over repeated factorisations', worst)
PY
done
