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
from ._arrays import to_device, kind_of, as_kind, default_device


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


def get_cg_solver(mat_times_vec, dim, cg_opts={}):
    """Conjugate-gradient solver (reference: ``solver_lib.py:70-98``).

    Follows ``scipy.sparse.linalg.cg`` step for step with scipy's *legacy*
    stopping rule the reference asks for (``atol='legacy'``, ``:93``): stop when
    ``||r|| < max(atol, tol * ||b||)``, ``tol`` defaulting to 1e-5, at most
    ``maxiter = 10 * dim`` iterations.  If the iteration does not converge a
    ``UserWarning('CG exited with error code ...')`` is issued and the current
    iterate is still returned (``:94-97``).  The vector updates and reductions
    run in the ``vt_cg_*`` kernels with all scalars on the device;
    ``mat_times_vec`` is called with the same array kind as ``v`` (a float64
    CUDA tensor if ``v`` is one, numpy if ``v`` is numpy).

    Supported ``cg_opts``: ``tol`` / ``rtol``, ``atol``, ``maxiter``, ``x0``,
    ``callback``.  A preconditioner ``M`` is not implemented.

    Extension: a ``(dim, K)`` right-hand side is
    solved column by column, so the closure can stand in for the Cholesky one
    at the matrix-RHS call sites (``sensitivity_lib.py:226``,
    ``lr_cov_lib.py:172``); scipy's CG, hence the reference, accepts vectors
    only."""
    opts = dict(cg_opts)
    if 'M' in opts and opts['M'] is not None:
        raise NotImplementedError('get_cg_solver: preconditioner `M` is not implemented on the GPU path')
    rtol = opts.get('rtol', opts.get('tol', 1e-5))
    atol = opts.get('atol', 0.0)
    if atol == 'legacy' or atol is None:
        atol = 0.0
    maxiter = opts.get('maxiter', None)
    x0 = opts.get('x0', None)
    callback = opts.get('callback', None)

    def solve(v):
        if getattr(v, 'ndim', 1) == 2:                       # any 2-d v, (dim, 1) included, is a matrix: shape preserved
            if v.shape[0] != dim:
                raise ValueError('right-hand side has shape {}, expected ({}, K)'.format(tuple(v.shape), dim))
            vd = to_device(v)
            cols = [to_device(solve(vd[:, k].contiguous()), vd.device) for k in range(vd.shape[1])]
            return as_kind(torch.stack(cols, dim=1), kind_of(v))
        kind = kind_of(v)
        b = to_device(v).reshape(-1).contiguous()
        if b.numel() != dim:
            raise ValueError('right-hand side has {} entries, expected {}'.format(b.numel(), dim))
        dev = b.device

        def matvec(p):
            q = mat_times_vec(as_kind(p, kind))
            return to_device(q, dev).reshape(-1).contiguous()

        x = torch.empty_like(b)
        r = torch.empty_like(b)
        p = torch.empty_like(b)
        state = torch.zeros(8, dtype=torch.float64, device=dev)
        ops.cg_init(b, x, r, state)
        bnrm2 = float(state[4].item()) ** 0.5
        if bnrm2 == 0.0:
            return as_kind(b.clone(), kind)
        if x0 is not None:
            x.copy_(to_device(x0, dev).reshape(-1))
            r.copy_(b - matvec(x))
            state[3] = torch.dot(r, r)
        tol = max(float(atol), float(rtol) * bnrm2)
        iters = dim * 10 if maxiter is None else int(maxiter)
        info = iters
        nmv = 0
        for it in range(iters):
            if float(state[3].item()) ** 0.5 < tol:
                info = 0
                break
            ops.cg_update_p(r, p, state, first=(it == 0))
            q = matvec(p)
            nmv += 1
            ops.cg_update_xr(p, q, x, r, state)
            if callback is not None:
                callback(as_kind(x, kind))
        solve.last_iterations = nmv
        if info != 0:
            warnings.warn('CG exited with error code {}'.format(info))
        return as_kind(x, kind)
    solve.last_iterations = None
    return solve
