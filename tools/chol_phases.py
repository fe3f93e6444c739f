"""Phase timing of the Cholesky diagonal kernel (needs a library built with
VT_NVCC_EXTRA=-DVT_CHOL_TIMING python -m vittles_b200.build --force)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops, _cabi
dev = torch.device('cuda', 0)
D = 128
A = torch.randn(D, D + 8, device=dev, dtype=torch.float64)
H = A @ A.T / D + torch.eye(D, device=dev, dtype=torch.float64)
for _ in range(3):
    ops.potrf(H)
torch.cuda.synchronize()
lib = _cabi.load()
buf = (ctypes.c_longlong * 32)()
lib.vt_debug_chol_clk(buf)
t = list(buf)
print('load            %6d clk' % (t[1] - t[0]))
for p in range(4):
    b = 5 * p
    print('panel %d: A1 %6d | wait %5d' % (p, t[2 + b] - (t[1] if p == 0 else t[6 + 5 * (p - 1)]), t[3 + b] - t[2 + b]), end='')
    if p < 3:
        print(' | A2 %6d (+sync %5d) | A3 %6d' % (t[4 + b] - t[3 + b], t[5 + b] - t[4 + b], t[6 + b] - t[5 + b]))
    else:
        print()
print('final binv      %6d clk' % (t[23] - t[22]))
print('stores          %6d clk' % (t[24] - t[23]))
print('total           %6d clk = %.1f us at 1.965 GHz' % (t[24] - t[0], (t[24] - t[0]) / 1965.0))
