"""Phase clocks of the last chol_diag_kernel launch (library built with VT_NVCC_EXTRA=-DVT_CHOL_TIMING):
    VT_NVCC_EXTRA=-DVT_CHOL_TIMING python -m vittles_b200.build --force && gpurun -- python tools/chol_diag_phases.py"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vittles_b200 import _cabi, ops  # noqa: E402

lib = _cabi.require_cuda()
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(128, 192, device=dev, dtype=torch.float64, generator=g)
H = A @ A.T / 128 + torch.eye(128, device=dev, dtype=torch.float64)
for _ in range(3):
    ops.potrf(H)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
rc = lib.vt_debug_chol_clk(buf)
t = list(buf)
names = {0: 'start', 1: 'staged'}
for p in range(4):
    names.update({2 + 5 * p: 'p%d A1 (warp 0: 32x32 chol)' % p, 3 + 5 * p: 'p%d sync after A1/Binv' % p,
                  4 + 5 * p: 'p%d A2 (panel substitution)' % p, 5 + 5 * p: 'p%d sync' % p, 6 + 5 * p: 'p%d A3 (DMMA update)' % p})
names.update({22: 'loop done', 23: 'last Binv row', 24: 'written back'})
prev = t[0]
rows = []
for k in sorted(names):
    if t[k] == 0:
        continue
    rows.append({'phase': names[k], 'clk_since_start': t[k] - t[0], 'clk_step': t[k] - prev})
    prev = t[k]
print(json.dumps({'rc': rc, 'phases': rows}, indent=1))
