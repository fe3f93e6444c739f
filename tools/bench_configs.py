#!/usr/bin/env python3
"""Measure BASELINE configs 3, 4 and 5 on one B200 (the headline config 2 is
bench.py).  Prints one JSON line per config with per-kernel times and the
achieved fraction of the bounding roofline (HBM copy bandwidth from
MEASURED_PEAKS.json, FP64 tensor peak from the in-run DMMA probe).

    python tools/bench_configs.py [3] [4] [5]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import vittles_b200 as vt  # noqa: E402
from vittles_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
try:
    HBM = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    HBM = 6534.1
PEAK = ops.fp64_peak_probe(0.2)


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def config3(N=1_000_000, K=20, d=16):
    """GMM mean-field VB with per-observation local parameters: block assembly,
    batched block Cholesky, Z = L^-1 C, tensor-core Schur complement, one solve."""
    M, Dg = K - 1, K * d
    g = torch.Generator(device=dev).manual_seed(3)
    centers = 0.35 * torch.randn(K, d, device=dev, dtype=torch.float64, generator=g)
    lab = torch.randint(0, K, (N,), device=dev, generator=g)
    X = centers[lab] + torch.randn(N, d, device=dev, dtype=torch.float64, generator=g)
    obj = vt.objectives.GMMVBObjective(X, K, prior_prec=0.5)
    # set-up (untimed): a few coordinate-ascent sweeps so that the Hessian is positive definite
    m = centers.clone()
    for _ in range(10):
        c = 0.5 * ((X[:, None, :] - m[None]) ** 2).sum(-1) + np.log(K)
        r = torch.softmax(-c, dim=1)
        m = (r.T @ X) / (r.sum(0)[:, None] + 0.5)
    c = 0.5 * ((X[:, None, :] - m[None]) ** 2).sum(-1) + np.log(K)
    rho = (-(c - c[:, -1:]))[:, :M].contiguous()
    del c, r
    x = torch.cat([m.reshape(-1), rho.reshape(-1)])
    sa = torch.as_tensor(obj.sparsity_array(), dtype=torch.int64, device=dev)
    res = {'config': 'GMM-VB K=20 d=16 N=1M (M=19 local, Dg=320 global)', 'n_obs': N}
    t_asm, h = timed(lambda: obj.vt_block_hessian(x, sa, which='full'), reps=2)
    bytes_asm = 8.0 * N * (M * M + M * Dg + d + M + 2 * K)
    res['assemble_blocks'] = {'ms': t_asm, 'gb_per_s': bytes_asm / t_asm / 1e6, 'frac_hbm': bytes_asm / t_asm / 1e6 / HBM}
    blocks0 = h.blocks.clone()
    t_potrf, Lb = timed(lambda: ops.block_potrf(blocks0.clone()), reps=3)
    t_clone, _ = timed(lambda: blocks0.clone(), reps=3)
    t_potrf -= t_clone
    res['block_potrf'] = {'ms': t_potrf, 'gb_per_s': 16.0 * N * M * M / t_potrf / 1e6,
                          'frac_hbm': 16.0 * N * M * M / t_potrf / 1e6 / HBM}
    cross = h.cross
    t0 = time.perf_counter()
    ops.block_trsm(Lb, cross)           # in place (no room to clone 48.6 GB three times); timed once, warm clocks
    torch.cuda.synchronize()
    t_trsm = (time.perf_counter() - t0) * 1e3
    res['block_trsm'] = {'ms': t_trsm, 'gb_per_s': 16.0 * N * M * Dg / t_trsm / 1e6,
                         'frac_hbm': 16.0 * N * M * Dg / t_trsm / 1e6 / HBM}
    Z2 = cross.reshape(N * M, Dg)
    t_schur, ZtZ = timed(lambda: ops.syrk_weighted(Z2), reps=2)
    fl = float(N) * M * Dg * (Dg + 1)
    res['schur_syrk'] = {'ms': t_schur, 'tflops_algorithmic': fl / t_schur / 1e9, 'frac_fp64_peak': fl / t_schur / 1e9 / PEAK,
                         'gb_per_s': 8.0 * N * M * Dg / t_schur / 1e6}
    S = h.hgg - ZtZ
    t_chol, fac = timed(lambda: ops.potrf(S), reps=3)
    res['schur_potrf_320'] = {'ms': t_chol}
    # one full solve through the public solver object (factor reuse): the pieces above, assembled by hand
    from vittles_b200.block_solver import BlockArrowSolver
    solver = BlockArrowSolver.__new__(BlockArrowSolver)
    solver.h, solver.d, solver.sa, solver.gi = h, h.shape[0], sa, h.global_inds
    solver.G, solver.M, solver.Dg, solver.Lb, solver.Z, solver.schur = N, M, Dg, Lb, cross, fac
    b = torch.randn(h.shape[0], device=dev, dtype=torch.float64, generator=g)
    t_solve, xsol = timed(lambda: solver._solve_vec(b), reps=3)
    res['solve_one_rhs'] = {'ms': t_solve, 'gb_per_s': 2 * 8.0 * N * M * Dg / t_solve / 1e6,
                            'frac_hbm': 2 * 8.0 * N * M * Dg / t_solve / 1e6 / HBM}
    # residual check without forming H: r = H x - b, block-arrow mat-vec with the ORIGINAL blocks / cross
    h2 = obj.vt_block_hessian(x, sa, which='full')
    xl = xsol[sa]
    rl = torch.einsum('gij,gj->gi', h2.blocks, xl) + torch.einsum('gik,k->gi', h2.cross, xsol[h.global_inds]) - b[sa]
    rg = torch.einsum('gik,gi->k', h2.cross, xl) + h2.hgg @ xsol[h.global_inds] - b[h.global_inds]
    res['residual_rel'] = float(max(rl.abs().max(), rg.abs().max()) / b.abs().max())
    res['total_factor_ms'] = t_asm + t_potrf + t_trsm + t_schur + t_chol
    return res


def config4(D=4096, Kmom=2048):
    """Dense LR covariance: potrf(4096) + potrs with 2048 right-hand sides + J H^-1 J^T."""
    g = torch.Generator(device=dev).manual_seed(4)
    A = torch.randn(D, D + 64, device=dev, dtype=torch.float64, generator=g)
    H = ops.gemm(A, A, 'KC', 'KC', alpha=1.0 / D)
    H.diagonal().add_(1.0)
    J = torch.randn(Kmom, D, device=dev, dtype=torch.float64, generator=g)
    res = {'config': 'LinearResponseCovariances D=4096, 2048 moments'}
    t_f, fac = timed(lambda: ops.potrf(H), reps=3)
    res['potrf'] = {'ms': t_f, 'tflops': D ** 3 / 3.0 / t_f / 1e9, 'frac_fp64_peak': D ** 3 / 3.0 / t_f / 1e9 / PEAK}
    Jt = J.T.contiguous()
    t_s, X = timed(lambda: fac.solve(Jt), reps=3)
    res['potrs_2048rhs'] = {'ms': t_s, 'tflops': 2.0 * D * D * Kmom / t_s / 1e9,
                            'frac_fp64_peak': 2.0 * D * D * Kmom / t_s / 1e9 / PEAK}
    t_g, C = timed(lambda: ops.gemm(J, X, 'KC', 'KS'), reps=3)
    res['gemm_J_X'] = {'ms': t_g, 'tflops': 2.0 * Kmom * Kmom * D / t_g / 1e9,
                       'frac_fp64_peak': 2.0 * Kmom * Kmom * D / t_g / 1e9 / PEAK}
    lr = vt.LinearResponseCovariances(lambda p: p.sum(), torch.zeros(D, device=dev, dtype=torch.float64), hessian_at_opt=H)
    t_all, cov = timed(lambda: lr.get_lr_covariance_from_jacobians(J, J), reps=2)
    res['get_lr_covariance_from_jacobians'] = {'ms': t_all}
    ref = J[:64] @ torch.linalg.solve(H, J[:64].T)
    res['rel_err_vs_cusolver_64x64'] = float((cov[:64, :64] - ref).abs().max() / ref.abs().max())
    return res


def config5(N=1_000_000, D=2048):
    """Order-3 Taylor expansion in a prior hyperparameter with get_cg_solver over the fused HVP."""
    X = ops.synth_design(5, 0, N, D, dev)
    theta_star = 0.5 * ops.synth_theta(5, D, dev)
    z = ops.glm_stats(X, theta_star, torch.zeros(N, device=dev, dtype=torch.float64), None, 'logistic', want_grad=False)[0]
    y = ops.synth_bernoulli(5, 0, z)
    obj = vt.objectives.GLMPriorObjective(X, y)
    eps0 = torch.tensor([np.log(2.0), 0.0], device=dev, dtype=torch.float64)
    theta = torch.zeros(D, device=dev, dtype=torch.float64)
    for _ in range(20):                       # Newton set-up on the same kernels (untimed)
        gth = obj.vt_grad(theta, eps0)
        step = ops.potrf(obj.vt_hessian(theta, eps0), overwrite=True).solve(gth)
        theta = theta - step
        if float(step.norm()) < 1e-11:
            break
    res = {'config': 'Taylor order 3, GLM + Gaussian prior (log tau, mu), D=2048, N=1M, get_cg_solver over fused HVP',
           'grad_norm_at_opt': float(obj.vt_grad(theta, eps0).norm())}
    hvp = obj.vt_hvp_fn(theta, eps0)
    v = torch.randn(D, device=dev, dtype=torch.float64)
    t_hvp, _ = timed(lambda: hvp(v), reps=10)
    gb = 8.0 * N * D / 1e9
    res['hvp'] = {'ms': t_hvp, 'gb_per_s': gb / t_hvp * 1e3, 'frac_hbm': gb / t_hvp * 1e3 / HBM}
    nmv = {'n': 0}

    def counted(vv):
        nmv['n'] += 1
        return hvp(vv)
    eps1 = eps0 + torch.tensor([0.3, -0.2], device=dev, dtype=torch.float64)
    for tol, key in [(1e-5, 'taylor3_cg_default_tol_1e-5'), (1e-10, 'taylor3_cg_tol_1e-10')]:
        solver = vt.solver_lib.get_cg_solver(counted, D, cg_opts={'tol': tol})
        te = vt.ParametricSensitivityTaylorExpansion(obj, theta, eps0, order=3, hess_solver=solver)
        nmv['n'] = 0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        series = te.evaluate_taylor_series(eps1)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        passes = nmv['n'] + 10                 # 10 directional-derivative sweeps for order 3
        res[key] = {'ms': ms, 'hvp_calls': nmv['n'], 'passes_over_X': passes, 'gb_per_s': passes * gb / ms * 1e3,
                    'frac_hbm': passes * gb / ms * 1e3 / HBM}
    te_chol = vt.ParametricSensitivityTaylorExpansion.optimization_objective(obj, theta, eps0, order=3)
    t_ch, series_chol = timed(lambda: te_chol.evaluate_taylor_series(eps1), reps=2)
    res['taylor3_cholesky_solver'] = {'ms': t_ch}
    res['cg_vs_cholesky_rel'] = float((series - series_chol).abs().max() / series_chol.abs().max())
    # how good is the expansion: distance to the re-optimised optimum
    th1 = theta.clone()
    for _ in range(20):
        step = ops.potrf(obj.vt_hessian(th1, eps1), overwrite=True).solve(obj.vt_grad(th1, eps1))
        th1 = th1 - step
        if float(step.norm()) < 1e-11:
            break
    res['taylor_error_vs_reoptimised'] = {'order0': float((theta - th1).norm()), 'order3': float((series_chol - th1).norm())}
    return res


if __name__ == '__main__':
    which = [int(a) for a in sys.argv[1:]] or [3, 4, 5]
    for c in which:
        out = {3: config3, 4: config4, 5: config5}[c]()
        out['fp64_peak_tflops'] = PEAK
        out['hbm_peak_gbs'] = HBM
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache()
