"""Where does the end-to-end step spend its time?  Same calls as bench.py's run_e2e,
with a synchronize + wall-clock stamp after each API call (2 steps; the second is printed)."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vittles_b200 as vt
from vittles_b200 import ops
dev = torch.device('cuda', 0)
N, D = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, 1024
X_host = torch.empty((N, D), dtype=torch.float64, pin_memory=True)
Xd = ops.synth_design(1, 0, N, D, dev); X_host.copy_(Xd); del Xd
y_host = (torch.rand(N) < 0.5).double().pin_memory()
w_host = torch.ones(N, dtype=torch.float64).pin_memory()
w1 = torch.ones(N, dtype=torch.float64); w1[::7] = 0.0; w1_host = w1.pin_memory()
theta_host = torch.zeros(D, dtype=torch.float64).pin_memory()
torch.cuda.empty_cache()
def now():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t = [now()]
    o = vt.objectives.GLMObjective(X_host, y_host, family='logistic'); t.append(now())
    sens = vt.HyperparameterSensitivityLinearApproximation(o, theta_host, w_host); t.append(now())
    pred = sens.predict_opt_par_from_hyper_par(w1_host); t.append(now())
    hess = sens.get_hessian_at_opt(); t.append(now())
    del o, sens, pred, hess; t.append(now())
    print('step %d: GLMObjective %.0f | Sens ctor %.0f | predict %.0f | get_hessian %.0f | del %.0f | total %.0f ms; reserved %.1f GB, gc %s' % (
        (it,) + tuple((t[i + 1] - t[i]) * 1e3 for i in range(5)) + ((t[-1] - t[0]) * 1e3, torch.cuda.memory_reserved() / 1e9, gc.get_count())), flush=True)
