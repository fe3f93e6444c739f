"""Oracle restatement of ``vittles/solver_lib.py`` (scipy, float64, CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import warnings

import numpy as np
import scipy as sp
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg


def get_dense_cholesky_solver(h, h_chol=None):
    """``solver_lib.py:7-30``: ``cho_factor`` once (``:27``), closure
    ``solve(v) = cho_solve(h_chol, v)`` (``:29``); ``h`` ignored if ``h_chol``
    is given.  ``v`` may be ``(D,)`` or ``(D, K)``."""
    if h_chol is None:
        h_chol = sp.linalg.cho_factor(h)

    def solve(v):
        return sp.linalg.cho_solve(h_chol, v)
    return solve


def get_sparse_cholesky_solver(h):
    """``solver_lib.py:33-48``: ``ValueError`` unless sparse (``:46-47``);
    SuperLU LU through ``scipy.sparse.linalg.factorized`` (``:48``)."""
    if not sp.sparse.issparse(h):
        raise ValueError('`h` must be sparse.')
    return sp.sparse.linalg.factorized(sp.sparse.csc_matrix(h))


def get_cholesky_solver(h):
    """``solver_lib.py:51-67``: dispatch on sparsity."""
    if sp.sparse.issparse(h):
        return get_sparse_cholesky_solver(h)
    return get_dense_cholesky_solver(h)


def _legacy_cg_opts(cg_opts):
    """The reference calls ``cg(linop, v, **cg_opts, atol='legacy')``
    (``solver_lib.py:93``).  scipy >= 1.14 removed ``'legacy'`` and renamed
    ``tol`` to ``rtol``.  Legacy semantics: stop when ``||r|| <= tol * ||b||``,
    default ``tol = 1e-5`` - i.e. today's ``rtol=tol, atol=0``."""
    opts = dict(cg_opts)
    if 'tol' in opts:
        opts['rtol'] = opts.pop('tol')
    opts.setdefault('rtol', 1e-5)
    opts['atol'] = 0.0
    return opts


def get_cg_solver(mat_times_vec, dim, cg_opts={}):
    """``solver_lib.py:70-98``: ``LinearOperator`` (``:91``), scipy ``cg``
    (``:93``), ``warnings.warn`` on non-zero info and still return ``x``
    (``:94-97``)."""
    linop = sp.sparse.linalg.LinearOperator((dim, dim), mat_times_vec, dtype=np.float64)
    opts = _legacy_cg_opts(cg_opts)

    def solve(v):
        x, info = sp.sparse.linalg.cg(linop, v, **opts)
        if info != 0:
            warnings.warn('CG exited with error code {}'.format(info))
        return x
    return solve


def cg_reference_iterates(mat_times_vec, b, rtol=1e-5, maxiter=None, x0=None):
    """Plain restatement of scipy's un-preconditioned ``cg`` loop
    (scipy/sparse/linalg/_isolve/iterative.py, 1.18.1) used to pin the
    ITERATION COUNT and stopping rule of the GPU CG:

        r = b - A x0;  atol = rtol*||b||;  for it in range(maxiter):
            if ||r|| < atol: return x, 0
            z = r; rho = r.z; p = z + (rho/rho_prev) p; q = A p
            alpha = rho / (p.q); x += alpha p; r -= alpha q
        return x, maxiter

    Returns ``(x, info, n_matvec)``."""
    b = np.asarray(b, dtype=np.float64)
    n = b.shape[0]
    if maxiter is None:
        maxiter = n * 10
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    bnrm2 = np.linalg.norm(b)
    if bnrm2 == 0:
        return b.copy(), 0, 0
    atol = rtol * bnrm2
    r = b - mat_times_vec(x) if x.any() else b.copy()
    rho_prev, p = None, None
    nmv = 0
    for _ in range(maxiter):
        if np.linalg.norm(r) < atol:
            return x, 0, nmv
        z = r
        rho_cur = np.dot(r, z)
        if rho_prev is not None:
            beta = rho_cur / rho_prev
            p = p * beta + z
        else:
            p = z.copy()
        q = mat_times_vec(p)
        nmv += 1
        alpha = rho_cur / np.dot(p, q)
        x = x + alpha * p
        r = r - alpha * q
        rho_prev = rho_cur
    return x, maxiter, nmv
