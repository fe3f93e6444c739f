"""Timeline of the streamed e2e sweep at full size: when does each chunk's copy end, when
does its stats + SYRK run?  (CUDA events; all times in ms from the first copy's start.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vittles_b200 import ops
dev = torch.device('cuda', 0)
N, D, NC = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, 1024, 16
Xh = torch.empty((N, D), dtype=torch.float64, pin_memory=True); Xh.fill_(0.01)
Xd = torch.empty((N, D), dtype=torch.float64, device=dev)
y = torch.ones(N, dtype=torch.float64, device=dev)
theta = torch.zeros(D, dtype=torch.float64, device=dev)
z = torch.empty(N, dtype=torch.float64, device=dev); resid = torch.empty_like(z); s = torch.empty_like(z)
Hc = torch.empty((D, D), dtype=torch.float64, device=dev)
ops.syrk_weighted(Xd[:N // NC], s[:N // NC], out=Hc)      # warm the workspace
ops.glm_stats(Xd[:N // NC], theta, y[:N // NC], None, 'logistic', want_grad=True, out=(z[:N // NC], resid[:N // NC], s[:N // NC]))
torch.cuda.synchronize()
for mode in ('copy only', 'copy + compute'):
    cs = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)
    E = lambda: torch.cuda.Event(enable_timing=True)
    t0 = E(); t0.record(cur); cs.wait_stream(cur)
    cend, kstart, kend = [], [], []
    with torch.cuda.stream(cs):
        for c in range(NC):
            r0, r1 = N * c // NC, N * (c + 1) // NC
            Xd[r0:r1].copy_(Xh[r0:r1], non_blocking=True)
            e = E(); e.record(cs); cend.append(e)
    if mode != 'copy only':
        for c in range(NC):
            r0, r1 = N * c // NC, N * (c + 1) // NC
            cur.wait_event(cend[c])
            a = E(); a.record(cur); kstart.append(a)
            ops.glm_stats(Xd[r0:r1], theta, y[r0:r1], None, 'logistic', want_grad=True, out=(z[r0:r1], resid[r0:r1], s[r0:r1]))
            ops.syrk_weighted(Xd[r0:r1], s[r0:r1], out=Hc)
            b = E(); b.record(cur); kend.append(b)
    torch.cuda.synchronize()
    print(mode)
    print('  copy ends :', ' '.join('%6.0f' % t0.elapsed_time(e) for e in cend))
    if kstart:
        print('  kern start:', ' '.join('%6.0f' % t0.elapsed_time(e) for e in kstart))
        print('  kern end  :', ' '.join('%6.0f' % t0.elapsed_time(e) for e in kend))
