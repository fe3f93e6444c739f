#!/usr/bin/env python3
"""Timing of the dense Cholesky factorisation and the multi-right-hand-side solve (config 4 sizes)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vittles_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
peak = ops.fp64_peak_probe(0.2)
for D, K in ((1024, 1024), (2048, 2048), (4096, 2048), (4096, 64)):
    g = torch.Generator(device=dev).manual_seed(D)
    A = torch.randn(D, D + 64, device=dev, dtype=torch.float64, generator=g)
    H = ops.gemm(A, A, 'KC', 'KC', alpha=1.0 / D)
    H.diagonal().add_(1.0)
    B = torch.randn(D, K, device=dev, dtype=torch.float64, generator=g)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, out
    Lbuf = torch.empty_like(H)

    def factor():                                   # into a preallocated buffer: no allocator traffic in the timing
        Lbuf.copy_(H)
        return ops.potrf(Lbuf, overwrite=True)
    factor()
    t_f, fac = timed(factor)
    t_s, X = timed(lambda: fac.solve(B))
    resid = float((H @ X - B).abs().max() / B.abs().max())
    print(json.dumps({'D': D, 'K': K, 'potrf_ms': t_f, 'potrf_tflops': D ** 3 / 3.0 / t_f / 1e9,
                      'potrf_frac_fp64_peak': D ** 3 / 3.0 / t_f / 1e9 / peak, 'potrs_ms': t_s,
                      'potrs_tflops': 2.0 * D * D * K / t_s / 1e9, 'potrs_frac_fp64_peak': 2.0 * D * D * K / t_s / 1e9 / peak,
                      'residual_rel': resid, 'fp64_peak_tflops': peak}), flush=True)
