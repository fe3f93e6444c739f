"""Solvers returning ``solve(v) -> H^{-1} v`` closures - the drop-in for
``vittles/solver_lib.py`` with the arithmetic on the GPU.

Same names, argument meaning and error behaviour as the reference
(SURVEY.md section 8b); ``v`` may be ``(D,)`` or ``(D, K)``, numpy or torch, and
the result comes back in the same kind.
"""
import warnings

import numpy as np
import scipy as sp
import scipy.sparse
import torch

from . import ops
from ._arrays import to_device, kind_of, as_kind


def get_dense_cholesky_solver(h, h_chol=None):
    """Dense Cholesky solver (reference: ``solver_lib.py:7-30``).

    ``cho_factor`` (``:27``) becomes ``vt_potrf``; the closure's ``cho_solve``
    (``:29``) becomes ``vt_potrs``.  ``h_chol`` may be an
    ``ops.CholeskyFactor`` (reused as is) or a ``(c, lower)`` tuple from
    ``scipy.linalg.cho_factor``; as upstream, ``h`` is then ignored.  Raises
    ``numpy.linalg.LinAlgError`` if ``h`` is not positive definite."""
    if h_chol is None:
        factor = ops.potrf(to_device(h))
    elif isinstance(h_chol, ops.CholeskyFactor):
        factor = h_chol
    else:
        c, lower = h_chol
        c = to_device(c)
        tri = torch.tril(c) if lower else torch.triu(c).T.contiguous()
        # rebuild H = L L^T with the GEMM engine and refactor: gives the blocked
        # factor layout (inverted diagonal blocks) the GPU solve needs
        factor = ops.potrf(ops.gemm(tri, tri, 'KC', 'KC'), overwrite=True)

    def solve(v):
        return as_kind(factor.solve(to_device(v, factor.L.device)), kind_of(v))
    solve.factor = factor
    return solve


def get_sparse_cholesky_solver(h):
    """Solver for a sparse ``h`` (reference: ``solver_lib.py:33-48``, which is
    SuperLU through ``scipy.sparse.linalg.factorized``).

    ``ValueError`` unless ``h`` is sparse (``:46-47``).  A block-arrow Hessian
    produced by :class:`vittles_b200.SparseBlockHessian` is factorised with the
    batched block-Cholesky + Schur-complement kernels; so is a scipy sparse
    matrix in which that structure is recognised
    (:meth:`BlockArrowHessian.from_sparse`, e.g. the reference's own
    ``coo_matrix`` Hessians).  Any other scipy sparse matrix is densified and
    factorised with the dense GPU Cholesky (the matrix must be symmetric
    positive definite, as the reference's name promises)."""
    from .sparse_hessian_lib import BlockArrowHessian
    if isinstance(h, BlockArrowHessian):
        return h.get_solver()
    if not sp.sparse.issparse(h):
        raise ValueError('`h` must be sparse.')
    arrow = BlockArrowHessian.from_sparse(h)
    if arrow is not None:
        factorised = arrow.get_solver()

        def solve(v):
            return factorised(v)
        solve.block_arrow = arrow      # (a bound method cannot carry attributes)
        return solve
    if h.shape[0] > 32768:
        raise ValueError('get_sparse_cholesky_solver: a general sparse matrix of dimension {} is too large to '
                         'densify; build it with SparseBlockHessian to use the block-arrow solver.'.format(h.shape[0]))
    return get_dense_cholesky_solver(np.asarray(h.todense()))


def get_cholesky_solver(h):
    """Dispatch on sparsity (reference: ``solver_lib.py:51-67``)."""
    from .sparse_hessian_lib import BlockArrowHessian
    if sp.sparse.issparse(h) or isinstance(h, BlockArrowHessian):
        return get_sparse_cholesky_solver(h)
    return get_dense_cholesky_solver(h)


class JacobiPreconditioner:
    """Diagonal preconditioner M = diag(1 / hess_diag) for ``cg_opts['M']``: applied inside the fused CG update
    kernel (no extra pass).  Also a valid scipy ``M`` (it has ``shape``, ``dtype``, ``matvec`` and ``@``)."""

    def __init__(self, hess_diag):
        d = to_device(hess_diag).reshape(-1)
        self.inv_diag = (1.0 / d).contiguous()
        self.shape = (d.numel(), d.numel())
        self.dtype = np.dtype(np.float64)

    def matvec(self, v):
        return as_kind(self.inv_diag * to_device(v, self.inv_diag.device), kind_of(v))

    __matmul__ = matvec
    __call__ = matvec


def _preconditioner(M, dim, dev):
    """(minv | None, apply | None) from scipy's ``M`` argument: an approximate INVERSE of the matrix, given as a
    JacobiPreconditioner, a diagonal scipy-sparse / dense matrix (fused in the update kernel), any dense (dim, dim)
    array (one GEMM per iteration), or a LinearOperator / callable (called per column like ``mat_times_vec``)."""
    if M is None:
        return None, None
    if isinstance(M, JacobiPreconditioner):
        return M.inv_diag.to(dev), None
    if sp.sparse.issparse(M):
        if M.shape != (dim, dim):
            raise ValueError('preconditioner M has shape {}, expected ({}, {})'.format(M.shape, dim, dim))
        c = M.tocoo()
        if np.all(c.row == c.col):
            return to_device(M.diagonal(), dev).contiguous(), None
        M = np.asarray(M.todense())
    if isinstance(M, (np.ndarray, torch.Tensor)):
        Md = to_device(M, dev)
        if tuple(Md.shape) != (dim, dim):
            raise ValueError('preconditioner M has shape {}, expected ({}, {})'.format(tuple(Md.shape), dim, dim))
        if bool((Md == torch.diag(torch.diagonal(Md))).all()):
            return torch.diagonal(Md).contiguous(), None
        Md = Md.contiguous()
        return None, lambda R: ops.gemm(R, Md, 'KC', 'KC')          # rows of R times M^T, (K, dim)
    fn = M.matvec if hasattr(M, 'matvec') else M
    if not callable(fn):
        raise TypeError('cg_opts["M"] must be a matrix, a LinearOperator or a callable')

    def apply(R):
        return torch.stack([to_device(fn(as_kind(R[k], 'numpy')), dev).reshape(-1) for k in range(R.shape[0])])
    return None, apply


def get_cg_solver(mat_times_vec, dim, cg_opts={}):
    """Conjugate-gradient solver (reference: ``solver_lib.py:70-98``).

    Follows ``scipy.sparse.linalg.cg`` step for step with scipy's *legacy*
    stopping rule the reference asks for (``atol='legacy'``, ``:93``): stop when
    ``||r|| < max(atol, tol * ||b||)``, ``tol`` defaulting to 1e-5, at most
    ``maxiter = 10 * dim`` iterations.  If the iteration does not converge a
    ``UserWarning('CG exited with error code ...')`` is issued and the current
    iterate is still returned (``:94-97``).  The vector updates, reductions and
    the convergence test run in the ``vt_cg_batch_*`` kernels with all scalars on
    the device; the host only polls a device flag every few iterations.
    ``mat_times_vec`` is called with the same array kind as ``v`` (a float64
    CUDA tensor if ``v`` is one, numpy if ``v`` is numpy).

    Supported ``cg_opts`` (scipy's names): ``tol`` / ``rtol``, ``atol``,
    ``maxiter``, ``x0``, ``callback``, ``M`` (see :func:`_preconditioner`;
    :class:`JacobiPreconditioner` is fused into the update kernel).

    Extension: a ``(dim, K)`` right-hand side is
    accepted and returned as ``(dim, K)``, so the closure can stand in for the
    Cholesky one at the matrix-RHS call sites (``sensitivity_lib.py:226``,
    ``lr_cov_lib.py:172``); scipy's CG, hence the reference, accepts vectors
    only.  The K columns are K independent CG iterations run side by side -
    each stops at the iteration scipy would stop at - that SHARE the
    matrix-vector product: a ``mat_times_vec`` with the attribute
    ``batched = True`` (``GLMObjective.vt_hvp_fn``) is called once per
    iteration with the ``(dim, K)`` matrix of search directions (one fused pass
    over X serves four columns), any other one column by column."""
    opts = dict(cg_opts)
    rtol = opts.get('rtol', opts.get('tol', 1e-5))
    atol = opts.get('atol', 0.0)
    if atol == 'legacy' or atol is None:
        atol = 0.0
    maxiter = opts.get('maxiter', None)
    x0 = opts.get('x0', None)
    callback = opts.get('callback', None)
    M_opt = opts.get('M', None)
    batched = bool(getattr(mat_times_vec, 'batched', False))

    def solve(v):
        kind = kind_of(v)
        vd = to_device(v)
        as_matrix = vd.dim() == 2                            # any 2-d v, (dim, 1) included: shape preserved
        if vd.dim() not in (1, 2) or vd.shape[0] != dim:
            raise ValueError('right-hand side has shape {}, expected ({},) or ({}, K)'.format(tuple(vd.shape), dim, dim))
        dev = vd.device
        B = (vd.T if as_matrix else vd.reshape(1, -1)).contiguous()          # (K, dim): one row per column
        K = B.shape[0]
        minv, papply = _preconditioner(M_opt, dim, dev)

        def matvec_rows(P):
            """rows of P -> rows of A P^T"""
            if batched and K > 1:
                Q = mat_times_vec(as_kind(P.T.contiguous(), kind))            # (dim, K) in, (dim, K) out
                return to_device(Q, dev).reshape(dim, K).T.contiguous()
            return torch.stack([to_device(mat_times_vec(as_kind(P[k].contiguous(), kind)), dev).reshape(-1)
                                for k in range(K)])

        X = torch.empty_like(B)
        R = torch.empty_like(B)
        P = torch.zeros_like(B)
        state = torch.zeros((K, 8), dtype=torch.float64, device=dev)
        if x0 is not None:
            x0d = to_device(x0, dev)
            X.copy_((x0d.T if x0d.dim() == 2 else x0d.reshape(1, -1)).expand_as(B))
            R.copy_(B - matvec_rows(X))
        ops.cg_batch_init(B, X, R, state, rtol, atol, keep_xr=x0 is not None)
        iters = dim * 10 if maxiter is None else int(maxiter)
        poll = 1
        for it in range(iters + 1):
            Z = papply(R) if papply is not None else None
            ops.cg_batch_update_p(R, P, state, iters, Z=Z, minv=minv)
            # all scalars are on the device; the host asks "is any column still running?" at iterations
            # 0, 1, 2, 4, 8, ... 32, 64, 96, ... (or every iteration for a callback): converged columns are frozen by
            # the kernels themselves, so polling late costs idle passes, never a different answer
            if callback is not None or it == poll or it == 0:
                if it == poll:
                    poll = poll * 2 if poll < 32 else poll + 32
                if not bool((state[:, 6] == 1.0).any()):
                    break
            Q = matvec_rows(P)
            ops.cg_batch_update_xr(P, Q, X, R, state)
            if callback is not None:
                callback(as_kind(X[0] if not as_matrix else X.T, kind))
        st = state.cpu()
        solve.last_iterations = int(st[:, 7].max().item())            # matrix-vector products of the slowest column
        solve.iterations_per_column = [int(i) for i in st[:, 7].tolist()]
        bad = st[:, 6] != 0.0
        if bool(bad.any()):
            warnings.warn('CG exited with error code {}'.format(iters))
        out = X.T.contiguous() if as_matrix else X.reshape(-1)
        return as_kind(out, kind)
    solve.last_iterations = None
    solve.iterations_per_column = None
    return solve
