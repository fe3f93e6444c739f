#!/usr/bin/env python3
"""Timing of the block-arrow kernels at the config-3 block shape (M = 19, Dg = 320)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vittles_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
G, M, Dg = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000, 19, 320
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(G, M, M, device=dev, dtype=torch.float64, generator=g)
blocks = a @ a.transpose(1, 2) / M + torch.eye(M, device=dev, dtype=torch.float64)
del a
C = torch.randn(G, M, Dg, device=dev, dtype=torch.float64, generator=g)
u = torch.randn(G * M, device=dev, dtype=torch.float64, generator=g)
xg = torch.randn(Dg, device=dev, dtype=torch.float64, generator=g)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_clone = timed(lambda: blocks.clone())
t_potrf = timed(lambda: ops.block_potrf(blocks.clone())) - t_clone
Lb = ops.block_potrf(blocks.clone())
t_trsm = timed(lambda: ops.block_trsm(Lb, C))                       # in place, repeated: timing only
t_trsmt = timed(lambda: ops.block_trsm(Lb, C, transpose=True))
Z2 = C.reshape(G * M, Dg)
t_colsum = timed(lambda: ops.tall_colsum(Z2, u))
t_gemv = timed(lambda: ops.tall_gemv(Z2, xg))
zb = 8.0 * G * M * Dg
print(json.dumps({'G': G, 'block_potrf_ms': t_potrf, 'block_potrf_gbs': 16.0 * G * M * M / t_potrf / 1e6,
                  'block_trsm_ms': t_trsm, 'block_trsm_gbs': 2 * zb / t_trsm / 1e6,
                  'block_trsmt_ms': t_trsmt, 'block_trsmt_gbs': 2 * zb / t_trsmt / 1e6,
                  'tall_colsum_ms': t_colsum, 'tall_colsum_gbs': zb / t_colsum / 1e6,
                  'tall_gemv_ms': t_gemv, 'tall_gemv_gbs': zb / t_gemv / 1e6}), flush=True)
