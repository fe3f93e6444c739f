"""Structured objective families whose derivatives have fused CUDA kernels.

The reference differentiates arbitrary autograd objectives; at N = 10^7,
D = 10^3 that is infeasible (D reverse sweeps over an N x D tape).  For the
benchmark families the derivatives are closed-form contractions over the
design matrix, and the API classes recognise these objects and dispatch to the
kernels (anything else goes through ``torch.func`` to a dense Hessian and then
the GPU solve - never silently to a slower structured path).

Every class here is ALSO a plain torch callable ``f(theta, hyper) -> scalar``
so that the generic autodiff path (and the oracle) can evaluate the same
objective on small instances.

Observation sharding: ``X``/``y`` hold this rank's rows; pass
``group=<torch.distributed process group>`` and the D-sized reductions
(gradient, Hessian, Hessian-vector products, directional derivatives) are
all-reduced, while per-observation outputs stay sharded.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops
from ._arrays import to_device


class StructuredObjective:
    """Marker base class: objectives with ``vt_*`` kernel hooks."""
    group = None

    @property
    def device(self):
        """The GPU that holds this objective's data (parameters are placed there, whatever device is current)."""
        X = getattr(self, 'X', None)
        return X.device if isinstance(X, torch.Tensor) else None

    def _allreduce(self, t):
        if self.group is not None and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def _b(z, family):
    if family == 'logistic':
        return torch.nn.functional.softplus(z)
    if family == 'poisson':
        return torch.exp(z)
    if family == 'gaussian':
        return 0.5 * z * z
    raise ValueError('unknown GLM family {!r}'.format(family))


class GLMObjective(StructuredObjective):
    """Weighted GLM negative log-likelihood with the per-observation weights as
    the hyperparameter (the infinitesimal-jackknife set-up of the reference's
    notebook, ``mle_weight_sensitivity_example.ipynb:345-371``):

        f(theta, w) = sum_n w_n [ b(x_n . theta) - y_n x_n . theta ] + l2/2 |theta|^2

    ``family``: 'logistic' (b = softplus), 'poisson' (b = exp), 'gaussian'.

    ``precision`` selects the engine of the two contractions (Hessian assembly
    and the H^{-1} G^T apply).  The two FP64-grade engines are held to the same
    rtol 1e-8 parity bar: 'f64' (FP64 DMMA) and 'f64_ozaki' (error-free slicing
    on the INT8 tensor cores, ~2.5x faster at N = 1e7, D = 1024); 'auto' (the
    default) takes the INT8 engine when N D^2 >= 1e11 and the weights are
    non-negative, the DMMA engine otherwise (``ops.resolve_precision``).
    Optional reduced precision: 'tf32' / 'tf32x3' (relative error about 5e-4 /
    2e-5 of the operand scale).  Statistics, factorisation and the inverse are
    FP64 on every engine."""

    def __init__(self, X, y, family='logistic', l2=0.0, device=None, group=None, stream_chunks=16,
                 precision='auto'):
        ops._split(precision)
        self.precision = precision
        self._pending, self._host_src = [], None
        if (isinstance(X, torch.Tensor) and not X.is_cuda and X.dim() == 2 and X.dtype == torch.float64
                and X.is_contiguous() and X.is_pinned() and X.shape[0] >= 64 * stream_chunks):
            self.X = self._stream_from_host(X, device, stream_chunks)
        else:
            self.X = to_device(X, device)
        if self.X.dim() != 2:
            raise ValueError('X must be (N, D)')
        self.X = self.X if self.X.stride(1) == 1 else self.X.contiguous()
        self.y = to_device(y, self.X.device).reshape(-1).contiguous()
        if self.y.numel() != self.X.shape[0]:
            raise ValueError('y must have one entry per row of X')
        if family not in ('logistic', 'poisson', 'gaussian'):
            raise ValueError('unknown GLM family {!r}'.format(family))
        self.family = family
        self.l2 = float(l2)
        self.group = group
        self.n_obs, self.dim = self.X.shape

    # -- host -> device streaming -----------------------------------------------
    def _stream_from_host(self, X_host, device, nchunks):
        """Pinned host design matrix: allocate the device copy now, transfer it in
        chunks on a side stream when the first sweep starts, with one event per
        chunk, so that the statistics pass and the Hessian assembly of chunk i
        overlap the transfer of chunk i+1 (the 82 GB transfer, not the arithmetic,
        bounds the end-to-end time).  The copies are NOT issued here: the H2D copy
        engine is FIFO, and small blocking copies issued later (y, theta, w) would
        wait behind the whole design matrix."""
        from ._arrays import default_device
        dev = default_device() if device is None else torch.device(device)
        self._host_src, self._nchunks = X_host, nchunks
        return torch.empty(X_host.shape, dtype=torch.float64, device=dev)

    def _start_copies(self):
        if self._host_src is None:
            return
        X_host, nchunks, Xd = self._host_src, self._nchunks, self.X
        self._host_src = None
        n = X_host.shape[0]
        copy_stream = torch.cuda.Stream(device=Xd.device)
        copy_stream.wait_stream(torch.cuda.current_stream(Xd.device))
        Xd.record_stream(copy_stream)
        with torch.cuda.stream(copy_stream):
            for c in range(nchunks):
                r0, r1 = (n * c) // nchunks, (n * (c + 1)) // nchunks
                Xd[r0:r1].copy_(X_host[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                self._pending.append((r0, r1, ev))

    def _wait_resident(self):
        """Make the current stream wait for the whole of X (no-op once resident)."""
        self._start_copies()
        if self._pending:
            torch.cuda.current_stream(self.X.device).wait_event(self._pending[-1][2])
            self._pending = []

    def vt_stats_and_hessian(self, theta, w):
        """(stats, H) in one sweep.  While a host transfer is in flight the sweep
        runs chunk by chunk behind the copies; otherwise it is vt_stats + vt_hessian."""
        if self._host_src is None and not self._pending:
            stats = self.vt_stats(theta, w, for_hessian=True)
            return stats, self.vt_hessian(theta, w, stats)
        dev = self.X.device
        theta = to_device(theta, dev)
        w = None if w is None else to_device(w, dev).contiguous()
        self._start_copies()          # after the small operands: the copy engine is FIFO
        n, d = self.X.shape
        z = torch.empty(n, dtype=torch.float64, device=dev)
        resid, s = torch.empty_like(z), torch.empty_like(z)
        grad = torch.zeros(d, dtype=torch.float64, device=dev)
        H = torch.zeros((d, d), dtype=torch.float64, device=dev)
        Hc = torch.empty_like(H)
        cur = torch.cuda.current_stream(dev)
        pending, self._pending = self._pending, []
        for r0, r1, ev in pending:
            cur.wait_event(ev)
            wc = None if w is None else w[r0:r1]
            fused = self._wants_colmax(r1 - r0, wc)
            st = ops.glm_stats(self.X[r0:r1], theta, self.y[r0:r1], wc, self.family, want_grad=True,
                               out=(z[r0:r1], resid[r0:r1], s[r0:r1]), want_colmax=fused)
            grad += st[3]
            ops.syrk_weighted(self.X[r0:r1], s[r0:r1], out=Hc, precision='f64_ozaki' if fused else self.precision,
                              colmax=st[4] if fused else None)
            H += Hc
        self._allreduce(grad)
        self._allreduce(H)
        if self.l2 != 0.0:
            grad = grad + self.l2 * theta
            H.diagonal().add_(self.l2)
        return dict(z=z, resid=resid, s=s, grad=grad), H

    # -- generic torch objective (small problems, autodiff checks) ----------
    def __call__(self, theta, w):
        self._wait_resident()
        z = self.X @ theta
        val = torch.sum(w * (_b(z, self.family) - self.y * z))
        if self.group is not None and dist.is_initialized():
            raise RuntimeError('the generic torch path of GLMObjective is single-process only')
        return val + 0.5 * self.l2 * torch.sum(theta * theta)

    # -- kernel hooks ---------------------------------------------------------
    def _wants_colmax(self, n_rows, w):
        """True when the Hessian assembly of these rows will run on the INT8 slicing engine and the statistics pass
        can hand it the per-feature scales (one sweep over X less)."""
        d = self.X.shape[1]
        return d <= ops.COLMAX_MAX_DIM and ops.resolve_precision(self.precision, n_rows, d, w) == 'f64_ozaki'

    def vt_stats(self, theta, w=None, want_grad=True, for_hessian=False):
        """z, resid = b'(z) - y, s = w b''(z) and the (all-reduced) gradient.  ``for_hessian``: the caller will
        assemble the Hessian from these statistics - on the INT8 slicing engine the pass then also produces the
        per-feature scales of that assembly (``colmax`` entry), sparing it a sweep over X."""
        self._wait_resident()
        theta = to_device(theta, self.X.device)
        w = None if w is None else to_device(w, self.X.device).contiguous()
        fused = for_hessian and self._wants_colmax(self.X.shape[0], w)
        st = ops.glm_stats(self.X, theta, self.y, w, self.family, l2=0.0, want_grad=want_grad, want_colmax=fused)
        z, resid, s, grad = st[:4]
        if grad is not None:
            self._allreduce(grad)
            if self.l2 != 0.0:
                grad = grad + self.l2 * theta
        out = dict(z=z, resid=resid, s=s, grad=grad)
        if fused:
            out['colmax'] = st[4]
        return out

    def vt_grad(self, theta, w):
        return self.vt_stats(theta, w)['grad']

    def vt_hessian(self, theta, w, stats=None):
        """H = X^T diag(w b''(z)) X + l2 I, all-reduced over the group."""
        if stats is None:
            stats = self.vt_stats(theta, w, want_grad=False, for_hessian=True)
        self._wait_resident()
        cm = stats.get('colmax')
        H = ops.syrk_weighted(self.X, stats['s'], l2=0.0, precision='f64_ozaki' if cm is not None else self.precision,
                              colmax=cm)
        self._allreduce(H)
        if self.l2 != 0.0:
            H.diagonal().add_(self.l2)
        return H

    def vt_ij_sensitivity(self, hinv, stats, out=None):
        """-H^{-1} G^T for this rank's observations, (D, N_local)."""
        self._wait_resident()
        return ops.ij_apply(hinv, self.X, stats['resid'], out=out, precision=self.precision)

    def vt_ij_sensitivity_by_substitution(self, factor, stats, out=None):
        """The same (D, N_local) matrix by forward / backward substitution with the Cholesky factor - what the
        reference's ``cho_solve`` does (``solver_lib.py:29``): -G^T is formed once (the transposing GEMM engine
        against the identity, the residuals as the column scale) and solved in place.  Three GEMM-sized passes
        instead of one, backward stable whatever the conditioning of H."""
        self._wait_resident()
        n, d = self.X.shape
        eye = torch.eye(d, dtype=torch.float64, device=self.X.device)
        gt = ops.gemm(eye, self.X, 'KC', 'KC', alpha=-1.0, colscale=stats['resid'], out=out)
        return factor.solve(gt, overwrite=True)

    def vt_hvp_fn(self, theta, w):
        """mat_times_vec for get_cg_solver: v -> H v, one fused pass over X."""
        s = self.vt_stats(theta, w, want_grad=False)['s']

        def hvp(v):
            """(D,) -> (D,), or (D, K) -> (D, K): the K columns share the passes over X (four per fused pass)."""
            v = to_device(v, self.X.device)
            if v.dim() == 2:
                out = ops.glm_hvp_multi(self.X, s, v.T.contiguous(), ridge=0.0).T.contiguous()
            else:
                out = ops.glm_hvp(self.X, s, v, ridge=0.0)
            self._allreduce(out)
            if self.l2 != 0.0:
                out = out + self.l2 * v
            return out
        hvp.batched = True
        return hvp

    def vt_directional_derivative(self, theta, w, eta_dirs, eps_dirs, cache=None):
        """d^{m+n} g / d theta^m d w^n contracted with the directions, where
        g = grad_theta f.  g is linear in w, so n >= 2 gives zero."""
        self._wait_resident()
        dev = self.X.device
        theta = to_device(theta, dev)
        m, n = len(eta_dirs), len(eps_dirs)
        if n >= 2:
            return torch.zeros_like(theta)
        src = w if n == 0 else eps_dirs[0]
        wts = None if src is None else to_device(src, dev).contiguous()
        if m == 0:
            st = ops.glm_stats(self.X, theta, self.y, wts, self.family, want_grad=True, want_z=False)
            out = self._allreduce(st[3])
            return out + self.l2 * theta if (n == 0 and self.l2 != 0.0) else out
        if cache is not None and 'z' in cache and torch.equal(cache['theta'], theta):
            z = cache['z']
        else:
            z = ops.glm_stats(self.X, theta, self.y, None, self.family, want_grad=False)[0]
            if cache is not None:
                cache['z'], cache['theta'] = z, theta.clone()
        dirs = torch.stack([to_device(v, dev) for v in eta_dirs])
        out = self._allreduce(ops.glm_dirderiv(self.X, z, dirs, wts, self.family))
        if m == 1 and n == 0 and self.l2 != 0.0:
            out = out + self.l2 * dirs[0]
        return out


class GLMPriorObjective(StructuredObjective):
    """GLM with a Gaussian prior whose log-precision and mean are the
    hyperparameter eps = (log tau, mu) - the "prior hyperparameter of a
    hierarchical model" family of benchmark config 5:

        f(theta, eps) = sum_n [ b(x_n . theta) - y_n x_n . theta ]
                        + exp(eps_0)/2 * |theta - eps_1|^2
    """

    def __init__(self, X, y, family='logistic', device=None, group=None):
        self._glm = GLMObjective(X, y, family=family, l2=0.0, device=device, group=group)
        self.X, self.y, self.family, self.group = self._glm.X, self._glm.y, family, group
        self.n_obs, self.dim = self.X.shape

    def __call__(self, theta, eps):
        z = self.X @ theta
        return torch.sum(_b(z, self.family) - self.y * z) + \
            0.5 * torch.exp(eps[0]) * torch.sum((theta - eps[1]) ** 2)

    def vt_grad(self, theta, eps):
        theta = to_device(theta, self.X.device)
        eps = to_device(eps, self.X.device)
        return self._glm.vt_stats(theta, None)['grad'] + torch.exp(eps[0]) * (theta - eps[1])

    def vt_hessian(self, theta, eps, stats=None):
        eps = to_device(eps, self.X.device)
        H = self._glm.vt_hessian(theta, None, stats)
        H.diagonal().add_(torch.exp(eps[0]))
        return H

    def vt_hvp_fn(self, theta, eps):
        tau = float(torch.exp(to_device(eps, self.X.device)[0]).item())
        s = self._glm.vt_stats(theta, None, want_grad=False)['s']

        def hvp(v):
            """(D,) -> (D,), or (D, K) -> (D, K): the K columns share the passes over X (four per fused pass)."""
            v = to_device(v, self.X.device)
            if v.dim() == 2:
                out = self._allreduce(ops.glm_hvp_multi(self.X, s, v.T.contiguous(), ridge=0.0).T.contiguous())
                return out + tau * v
            if self.group is None:
                return ops.glm_hvp(self.X, s, v, ridge=tau)
            out = self._allreduce(ops.glm_hvp(self.X, s, v, ridge=0.0))
            return out + tau * v
        hvp.batched = True
        return hvp

    def vt_directional_derivative(self, theta, eps, eta_dirs, eps_dirs, cache=None):
        dev = self.X.device
        theta, eps = to_device(theta, dev), to_device(eps, dev)
        m, n = len(eta_dirs), len(eps_dirs)
        out = torch.zeros_like(theta)
        # data part: independent of eps
        if n == 0:
            out = out + self._glm.vt_directional_derivative(theta, None, eta_dirs, [], cache=cache)
        # prior part: g_p = tau (theta - mu 1), tau = exp(eps_0)
        tau = torch.exp(eps[0])
        d0 = [float(to_device(d, dev)[0]) for d in eps_dirs]
        d1 = [float(to_device(d, dev)[1]) for d in eps_dirs]
        p0 = float(np.prod(d0)) if n > 0 else 1.0
        if m == 0:
            cross = sum(d1[i] * float(np.prod(d0[:i] + d0[i + 1:])) for i in range(n))
            out = out + tau * (p0 * (theta - eps[1]) - cross)
        elif m == 1:
            out = out + tau * p0 * to_device(eta_dirs[0], dev)
        return out


class GMMVBObjective(StructuredObjective):
    """Mean-field VB for a Gaussian mixture with unit covariances, with
    per-observation local parameters (benchmark config 3).  Flat parameter

        x = ( m (K*d global means),  rho_1 .. rho_N (K-1 free logits each) )

        r_n = softmax([rho_n, 0]),  c_nk = |x_n - m_k|^2 / 2 - log pi_k
        f(x) = sum_n sum_k r_nk (c_nk + log r_nk) + prior_prec/2 * |m|^2

    The Hessian is block-arrow: (K-1)x(K-1) local blocks, (K-1)x(K*d) cross
    blocks, and a diagonal global block.  ``vt_block_hessian`` assembles all of
    it in closed form in one kernel pass (``vt_gmm_blocks``)."""

    def __init__(self, X, K, log_pi=None, prior_prec=1e-2, device=None, group=None):
        self.X = to_device(X, device).contiguous()
        self.n_obs, self.d = self.X.shape
        self.K = int(K)
        self.log_pi = (torch.full((self.K,), -float(np.log(self.K)), dtype=torch.float64, device=self.X.device)
                       if log_pi is None else to_device(log_pi, self.X.device))
        self.prior_prec = float(prior_prec)
        self.group = group
        self.n_global = self.K * self.d
        self.dim = self.n_global + self.n_obs * (self.K - 1)

    def sparsity_array(self):
        return self.n_global + np.arange(self.n_obs * (self.K - 1)).reshape(self.n_obs, self.K - 1)

    def _split(self, x):
        x = to_device(x, self.X.device)
        return x[:self.n_global].reshape(self.K, self.d), x[self.n_global:].reshape(self.n_obs, self.K - 1)

    def __call__(self, x):
        m, rho = x[:self.n_global].reshape(self.K, self.d), x[self.n_global:].reshape(self.n_obs, self.K - 1)
        logits = torch.cat([rho, torch.zeros(self.n_obs, 1, dtype=x.dtype, device=x.device)], dim=1)
        logr = torch.log_softmax(logits, dim=1)
        r = torch.exp(logr)
        c = 0.5 * ((self.X[:, None, :] - m[None, :, :]) ** 2).sum(-1) - self.log_pi[None, :]
        return torch.sum(r * (c + logr)) + 0.5 * self.prior_prec * torch.sum(m * m)

    def vt_grad(self, x):
        m, rho = self._split(x)
        out = ops.gmm_blocks(self.X, m, rho, self.log_pi, want_blocks=False, want_cross=False)
        r = out['r']
        # d f / d m_k = sum_n r_nk (m_k - x_n) + prior_prec m_k: only the data term is summed over the ranks
        rsum = r.sum(0)
        gm = self._allreduce(rsum[:, None] * m - ops.gemm(r, self.X, 'KS', 'KS')) + self.prior_prec * m
        return torch.cat([gm.reshape(-1), out['grad_rho'].reshape(-1)])

    def vt_hessian(self, x):
        raise NotImplementedError('the dense Hessian of a GMM-VB objective is never formed; use SparseBlockHessian')

    def vt_block_hessian(self, x, sparsity_array, which='full', global_inds=None):
        from .sparse_hessian_lib import BlockArrowHessian
        m, rho = self._split(x)
        want_b = which in ('full', 'block')
        want_g = which in ('full', 'global')
        out = ops.gmm_blocks(self.X, m, rho, self.log_pi, want_blocks=want_b, want_cross=want_g)
        dev = self.X.device
        gi = torch.arange(self.n_global, dtype=torch.int64, device=dev) if want_g else \
            torch.empty(0, dtype=torch.int64, device=dev)
        hgg = None
        if want_g:
            rsum = self._allreduce(out['r'].sum(0))
            hgg = torch.diag((rsum + self.prior_prec).repeat_interleave(self.d))
        return BlockArrowHessian(self.dim, sparsity_array, gi, blocks=out['blocks'], cross=out['cross'], hgg=hgg,
                                 group=self.group)
