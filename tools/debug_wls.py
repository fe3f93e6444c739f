import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vittles_b200 as vt
from oracle import fixtures, sensitivity as osens
objective, run_regression, x, y = fixtures.wls_fixture()
n = len(y)
w1 = np.ones(n); dw = np.random.RandomState(1).uniform(size=n) - 0.5
theta0 = run_regression(torch.as_tensor(w1)).numpy()
obj = vt.objectives.GLMObjective(x, y, family='gaussian')
xt, yt = torch.as_tensor(x), torch.as_tensor(y)
def f(theta, w):
    z = xt @ theta
    return torch.sum(w * (0.5 * z * z - yt * z))
g = torch.func.grad(f, argnums=0)
rng = np.random.RandomState(2)
dirs = [rng.normal(size=2) for _ in range(3)]
for m, ne in [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (0, 2)]:
    ours = obj.vt_directional_derivative(theta0, w1, dirs[:m], [dw] * ne).cpu().numpy()
    ref = osens.directional_derivative(g, theta0, w1, dirs[:m], [dw] * ne)
    print(m, ne, ours, ref)
print('H', obj.vt_hessian(theta0, w1).cpu().numpy(), (x.T * w1) @ x)
