#!/usr/bin/env python3
"""BASELINE configs 3, 4 and 5 (the headline config 2 is bench.py): one dict per config with the step time, the
per-kernel times against the bounding roofline (HBM copy bandwidth from MEASURED_PEAKS.json, FP64 tensor peak from
the in-run DMMA probe) and a size-independent correctness property.  bench.py calls run_all() (all ranks of a
torchrun job take part: config 3 and 5 shard the observations, config 4 splits the right-hand-side columns);

    python tools/bench_configs.py [3] [4] [5]            # stand-alone, one GPU, one JSON line per config
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class Ctx:
    def __init__(self, dev, group=None, hbm=None, peak=None):
        from vittles_b200 import ops
        self.dev, self.group = dev, group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if group is not None else 0
        if hbm is None:
            try:
                hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
            except Exception:
                hbm = 6650.0                                  # the profiling recipe's stated fallback
        self.hbm = hbm
        self.peak = ops.fp64_peak_probe(0.2) if peak is None else peak

    def max_over_ranks(self, ms):
        if self.group is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def timed(self, fn, reps=3, warm=1):
        """mean device time (CUDA events on the launching stream) of `reps` calls after `warm` warm-up calls,
        max over the ranks"""
        out = None
        for _ in range(warm):
            out = None                     # drop the previous result first: its blocks go back to the caching allocator
            out = fn()                     # and no cudaMalloc of a multi-GB result lands in the timed region
        torch.cuda.synchronize()
        if self.group is not None:
            dist.barrier(self.group)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = None
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return self.max_over_ranks(a.elapsed_time(b) / reps), out


class PassCounter:
    """Counts the sweeps over the design matrix that the library performs (every fused-pass entry point of ops)."""
    NAMES = ('glm_stats', 'glm_hvp', 'glm_dirderiv')

    def __init__(self, ops):
        self.ops, self.n, self.saved = ops, 0, {}

    def __enter__(self):
        for name in self.NAMES:
            fn = getattr(self.ops, name)
            self.saved[name] = fn

            def wrapped(*a, _fn=fn, **k):
                self.n += 1
                return _fn(*a, **k)
            setattr(self.ops, name, wrapped)
        multi = self.ops.glm_hvp_multi
        self.saved['glm_hvp_multi'] = multi

        def wrapped_multi(X, s, V, ridge=0.0):
            K = V.shape[0]
            self.n += 2 if K >= self.ops.HVP_GEMM_MIN else -(-K // self.ops.HVP_MULTI_MAX)
            return multi(X, s, V, ridge)
        self.ops.glm_hvp_multi = wrapped_multi
        return self

    def __exit__(self, *exc):
        for name, fn in self.saved.items():
            setattr(self.ops, name, fn)


def config3(cx, N=1_000_000, K=20, d=16):
    """GMM mean-field VB with per-observation local parameters (SparseBlockHessian path): closed-form block
    assembly, batched block Cholesky, Z = L^-1 C, tensor-core Schur complement (all-reduced over the ranks), dense
    Schur factorisation; then solves with 1 and 8 right-hand sides.  Observations are sharded over the ranks."""
    import vittles_b200 as vt
    from vittles_b200 import ops
    from vittles_b200.block_solver import BlockArrowSolver
    dev = cx.dev
    M, Dg = K - 1, K * d
    n0, n1 = (N * cx.rank) // cx.world, (N * (cx.rank + 1)) // cx.world
    nl = n1 - n0
    g = torch.Generator(device=dev).manual_seed(3)
    centers = 0.35 * torch.randn(K, d, device=dev, dtype=torch.float64, generator=g)       # same on every rank
    g2 = torch.Generator(device=dev).manual_seed(1000 + cx.rank)
    lab = torch.randint(0, K, (nl,), device=dev, generator=g2)
    X = centers[lab] + torch.randn(nl, d, device=dev, dtype=torch.float64, generator=g2)
    obj = vt.objectives.GMMVBObjective(X, K, prior_prec=0.5, group=cx.group)
    # set-up (untimed): a few coordinate-ascent sweeps so that the Hessian is positive definite
    m = centers.clone()
    for _ in range(10):
        c = 0.5 * ((X[:, None, :] - m[None]) ** 2).sum(-1) + np.log(K)
        r = torch.softmax(-c, dim=1)
        num, den = r.T @ X, r.sum(0)
        if cx.group is not None:
            dist.all_reduce(num, group=cx.group)
            dist.all_reduce(den, group=cx.group)
        m = num / (den[:, None] + 0.5)
    c = 0.5 * ((X[:, None, :] - m[None]) ** 2).sum(-1) + np.log(K)
    rho = (-(c - c[:, -1:]))[:, :M].contiguous()
    del c, r
    x = torch.cat([m.reshape(-1), rho.reshape(-1)])
    sa = torch.as_tensor(obj.sparsity_array(), dtype=torch.int64, device=dev)
    res = {'config': 'GMM-VB K={} d={} N={} (M={} local, Dg={} global), observations over {} rank(s)'.format(
        K, d, N, M, Dg, cx.world), 'n_obs': N}
    t_asm, h = cx.timed(lambda: obj.vt_block_hessian(x, sa, which='full'), reps=2)
    bytes_asm = 8.0 * nl * (M * M + M * Dg + d + M + 2 * K)
    res['assemble_blocks'] = {'ms': t_asm, 'gb_per_s': bytes_asm / t_asm / 1e6, 'frac_hbm': bytes_asm / t_asm / 1e6 / cx.hbm}
    # the factorisation through the public solver object (overwrite=True: no room for a second 6 GB copy per
    # million observations); its pieces are timed one by one on a copy of the small parts
    blocks0 = h.blocks.clone()
    t_potrf, Lb = cx.timed(lambda: ops.block_potrf(blocks0.clone()), reps=3)
    t_clone, _ = cx.timed(lambda: blocks0.clone(), reps=3)
    t_potrf = max(t_potrf - t_clone, 1e-3)
    res['block_potrf'] = {'ms': t_potrf, 'gb_per_s': 16.0 * nl * M * M / t_potrf / 1e6,
                          'frac_hbm': 16.0 * nl * M * M / t_potrf / 1e6 / cx.hbm}
    del blocks0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    solver = BlockArrowSolver(h, overwrite=True)          # in place: timed once, clocks warm from the pieces above
    e1.record()
    torch.cuda.synchronize()
    t_factor = cx.max_over_ranks(e0.elapsed_time(e1))
    res['factorise_total'] = {'ms': t_factor, 'what': 'batched Cholesky + Z = L^-1 C in place + Z^T Z (FP64 tensor core) '
                              '+ all-reduce of the {0}x{0} Schur complement + its Cholesky'.format(Dg)}
    Z2 = solver.Z.reshape(nl * M, Dg)
    t_schur, _ = cx.timed(lambda: ops.syrk_weighted(Z2, precision='auto'), reps=2)
    fl = float(nl) * M * Dg * (Dg + 1)
    res['schur_syrk'] = {'ms': t_schur, 'engine': ops.resolve_precision('auto', nl * M, Dg), 'fp64_equiv_tflops_algorithmic': fl / t_schur / 1e9,
                         'ratio_to_fp64_pipe_peak': fl / t_schur / 1e9 / cx.peak}
    trsm_ms = max(t_factor - t_potrf - t_schur, 1e-3)
    res['block_trsm_estimate'] = {'ms': trsm_ms, 'gb_per_s': 16.0 * nl * M * Dg / trsm_ms / 1e6,
                                  'frac_hbm': 16.0 * nl * M * Dg / trsm_ms / 1e6 / cx.hbm,
                                  'note': 'factorise_total minus the batched Cholesky and the Schur SYRK'}
    gb = torch.Generator(device=dev).manual_seed(77)
    b_glob = torch.randn(Dg, 8, device=dev, dtype=torch.float64, generator=gb)               # replicated part
    b = torch.cat([b_glob, torch.randn(nl * M, 8, device=dev, dtype=torch.float64, generator=g2)], dim=0)
    t_solve1, x1 = cx.timed(lambda: solver.solve(b[:, 0].contiguous()), reps=3)
    z_bytes = 8.0 * nl * M * Dg
    res['solve_1_rhs'] = {'ms': t_solve1, 'gb_per_s': 2 * z_bytes / t_solve1 / 1e6, 'frac_hbm': 2 * z_bytes / t_solve1 / 1e6 / cx.hbm}
    t_solve8, x8 = cx.timed(lambda: solver.solve(b), reps=2)
    res['solve_8_rhs'] = {'ms': t_solve8, 'ms_per_rhs': t_solve8 / 8, 'passes_over_Z': 2,
                          'note': 'all columns together: Z is read twice per solve whatever the number of columns'}
    # residual without forming H: r = H x - b, block-arrow mat-vec with freshly assembled blocks
    h2 = obj.vt_block_hessian(x, sa, which='full')
    gi = h2.global_inds
    xl = x8[sa]
    rl = torch.einsum('gij,gjk->gik', h2.blocks, xl) + torch.einsum('gij,jk->gik', h2.cross, x8[gi]) - b[sa]
    rg = torch.einsum('gij,gik->jk', h2.cross, xl)
    if cx.group is not None:
        dist.all_reduce(rg, group=cx.group)
    rg = rg + h2.hgg @ x8[gi] - b[gi]
    rr = torch.stack([rl.abs().max(), rg.abs().max(), b.abs().max()])
    if cx.group is not None:
        dist.all_reduce(rr, op=dist.ReduceOp.MAX, group=cx.group)
    res['residual_rel_8_rhs'] = float(max(rr[0], rr[1]) / rr[2])
    res['solve_1_vs_8_rhs_max_diff'] = float((x1 - x8[:, 0]).abs().max())
    step_ms = t_asm + t_factor + t_solve1
    res['step'] = {'what': 'assemble + factorise + one solve', 'ms': step_ms, 'value': N / (step_ms * 1e-3), 'unit': 'obs/s'}
    return res


def config4(cx, D=4096, Kmom=2048):
    """Dense LR covariance: potrf(4096) + potrs with the moment columns + J H^-1 J^T.  With several ranks the
    columns of J^T are split over the ranks (every rank factorises; no communication)."""
    import vittles_b200 as vt
    from vittles_b200 import ops
    dev = cx.dev
    g = torch.Generator(device=dev).manual_seed(4)
    A = torch.randn(D, D + 64, device=dev, dtype=torch.float64, generator=g)
    H = ops.gemm(A, A, 'KC', 'KC', alpha=1.0 / D)
    H.diagonal().add_(1.0)
    del A
    J = torch.randn(Kmom, D, device=dev, dtype=torch.float64, generator=g)
    k0, k1 = (Kmom * cx.rank) // cx.world, (Kmom * (cx.rank + 1)) // cx.world
    Jl = J[k0:k1].contiguous()
    res = {'config': 'LinearResponseCovariances D={}, {} moments, moment columns over {} rank(s)'.format(D, Kmom, cx.world)}
    Lbuf = torch.empty_like(H)

    def factor():                                   # into a preallocated buffer (the 134 MB copy of H is 0.05 ms)
        Lbuf.copy_(H)
        return ops.potrf(Lbuf, overwrite=True)
    t_f, fac = cx.timed(factor, reps=5, warm=2)
    res['potrf'] = {'ms': t_f, 'tflops': D ** 3 / 3.0 / t_f / 1e9, 'frac_fp64_peak': D ** 3 / 3.0 / t_f / 1e9 / cx.peak}
    Jt = Jl.T.contiguous()
    t_s, X = cx.timed(lambda: fac.solve(Jt), reps=3)
    kl = k1 - k0
    res['potrs'] = {'ms': t_s, 'rhs_columns_per_rank': kl, 'tflops': 2.0 * D * D * kl / t_s / 1e9,
                    'frac_fp64_peak': 2.0 * D * D * kl / t_s / 1e9 / cx.peak}
    t_g, C = cx.timed(lambda: ops.gemm(J, X, 'KC', 'KS'), reps=3)
    res['gemm_J_Hinv_Jt'] = {'ms': t_g, 'tflops': 2.0 * Kmom * kl * D / t_g / 1e9,
                             'frac_fp64_peak': 2.0 * Kmom * kl * D / t_g / 1e9 / cx.peak}
    lr = vt.LinearResponseCovariances(lambda p: p.sum(), torch.zeros(D, device=dev, dtype=torch.float64), hessian_at_opt=H)
    t_all, cov = cx.timed(lambda: lr.get_lr_covariance_from_jacobians(J, Jl), reps=2)
    res['get_lr_covariance_from_jacobians'] = {'ms': t_all, 'what': 'solve + GEMM with the factor held by the object'}
    # residual of the solve itself: H X = J^T
    rel = (ops.gemm(H, X.T.contiguous(), 'KC', 'KC') - Jt).abs().max() / Jt.abs().max()
    res['solve_residual_rel'] = float(rel)
    step_ms = t_f + t_all
    res['step'] = {'what': 'factorise + covariance of all moments', 'ms': step_ms, 'value': Kmom / (step_ms * 1e-3),
                   'unit': 'moments/s'}
    return res


def config5(cx, N=1_000_000, D=2048):
    """Order-3 Taylor expansion in a prior hyperparameter with get_cg_solver over the fused HVP; observations
    sharded over the ranks (one D-vector all-reduce per Hessian-vector product / directional derivative)."""
    import vittles_b200 as vt
    from vittles_b200 import ops
    dev = cx.dev
    n0, n1 = (N * cx.rank) // cx.world, (N * (cx.rank + 1)) // cx.world
    nl = n1 - n0
    X = ops.synth_design(5, n0, nl, D, dev)
    theta_star = 0.5 * ops.synth_theta(5, D, dev)
    z = ops.glm_stats(X, theta_star, torch.zeros(nl, device=dev, dtype=torch.float64), None, 'logistic', want_grad=False)[0]
    y = ops.synth_bernoulli(5, n0, z)
    obj = vt.objectives.GLMPriorObjective(X, y, group=cx.group)
    eps0 = torch.tensor([np.log(2.0), 0.0], device=dev, dtype=torch.float64)
    theta = torch.zeros(D, device=dev, dtype=torch.float64)
    for _ in range(20):                       # Newton set-up on the same kernels (untimed)
        gth = obj.vt_grad(theta, eps0)
        step = ops.potrf(obj.vt_hessian(theta, eps0), overwrite=True).solve(gth)
        theta = theta - step
        if float(step.norm()) < 1e-11:
            break
    res = {'config': 'Taylor order 3, GLM + Gaussian prior (log tau, mu), D={}, N={}, get_cg_solver over the fused HVP, '
                     'observations over {} rank(s)'.format(D, N, cx.world),
           'grad_norm_at_opt': float(obj.vt_grad(theta, eps0).norm())}
    hvp = obj.vt_hvp_fn(theta, eps0)
    v = torch.randn(D, device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(9))
    t_hvp, _ = cx.timed(lambda: hvp(v), reps=10, warm=2)
    gb = 8.0 * nl * D / 1e9
    res['hvp'] = {'ms': t_hvp, 'gb_per_s': gb / t_hvp * 1e3, 'frac_hbm': gb / t_hvp * 1e3 / cx.hbm}
    V4 = torch.randn(D, 4, device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(10))
    t_hvp4, _ = cx.timed(lambda: hvp(V4), reps=10, warm=2)
    res['hvp_4_columns_one_pass'] = {'ms': t_hvp4, 'gb_per_s': gb / t_hvp4 * 1e3, 'frac_hbm': gb / t_hvp4 * 1e3 / cx.hbm}
    eps1 = eps0 + torch.tensor([0.3, -0.2], device=dev, dtype=torch.float64)
    series = None
    for tol, key in [(1e-5, 'taylor3_cg_default_tol_1e-5'), (1e-10, 'taylor3_cg_tol_1e-10')]:
        solver = vt.solver_lib.get_cg_solver(hvp, D, cg_opts={'tol': tol})

        def run():
            te = vt.ParametricSensitivityTaylorExpansion(obj, theta, eps0, order=3, hess_solver=solver)
            return te.evaluate_taylor_series(eps1)
        with PassCounter(ops) as pc:
            run()                                           # warm-up, and the pass count of one evaluation
            passes = pc.n
        ms, series = cx.timed(run, reps=2, warm=0)
        res[key] = {'ms': ms, 'passes_over_X': passes, 'gb_per_s': passes * gb / ms * 1e3,
                    'frac_hbm': passes * gb / ms * 1e3 / cx.hbm,
                    'note': 'passes counted at the fused-pass entry points of ops (HVPs of the three CG solves, '
                            'directional-derivative sweeps, statistics); a read-only stream can exceed the copy peak'}
    te_chol = vt.ParametricSensitivityTaylorExpansion.optimization_objective(obj, theta, eps0, order=3)
    t_ch, series_chol = cx.timed(lambda: te_chol.evaluate_taylor_series(eps1), reps=2)
    res['taylor3_cholesky_solver'] = {'ms': t_ch}
    res['cg_vs_cholesky_rel'] = float((series - series_chol).abs().max() / series_chol.abs().max())
    th1 = theta.clone()                        # how good is the expansion: distance to the re-optimised optimum
    for _ in range(20):
        step = ops.potrf(obj.vt_hessian(th1, eps1), overwrite=True).solve(obj.vt_grad(th1, eps1))
        th1 = th1 - step
        if float(step.norm()) < 1e-11:
            break
    res['taylor_error_vs_reoptimised'] = {'order0': float((theta - th1).norm()), 'order3': float((series_chol - th1).norm())}
    ms = res['taylor3_cg_default_tol_1e-5']['ms']
    res['step'] = {'what': 'order-3 Taylor series with the CG solver at the reference default tolerance', 'ms': ms,
                   'value': N / (ms * 1e-3), 'unit': 'obs/s'}
    return res


CONFIGS = {3: config3, 4: config4, 5: config5}


def run_all(dev, group=None, which=(3, 4, 5), peak=None):
    """{'config3': ..., ...}; a failure of one config is reported in its slot, never raised."""
    import gc
    from vittles_b200 import ops
    cx = Ctx(dev, group, peak=peak)
    out = {'fp64_peak_tflops': cx.peak, 'hbm_peak_gbs': cx.hbm}
    for c in which:
        try:
            out['config{}'.format(c)] = CONFIGS[c](cx)
        except Exception as exc:                     # report, never hide
            out['config{}'.format(c)] = {'error': repr(exc)[:400]}
        gc.collect()
        ops.free_workspaces()
        torch.cuda.empty_cache()
    return out


if __name__ == '__main__':
    which = [int(a) for a in sys.argv[1:] if a.isdigit()] or [3, 4, 5]
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        group = dist.group.WORLD
    res = run_all(torch.device('cuda', local_rank), group, which)
    if int(os.environ.get('RANK', '0')) == 0:
        for k, v in res.items():
            print(json.dumps({k: v}), flush=True)
    if world > 1:
        dist.destroy_process_group()
