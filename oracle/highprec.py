"""Extended-precision (x87 80-bit ``numpy.longdouble``, eps 1.1e-19) restatement of the infinitesimal-jackknife
solve, used to measure the FORWARD error of float64 algorithms on ill-conditioned Hessians.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows ``sensitivity_lib.py:226`` (``-hess_solver(cross)``)
with ``solver_lib.py:27-29`` (Cholesky factor + two triangular solves) in a wider type: neither the reference's
``cho_solve`` nor this build's kernels can be told apart at kappa(H) ~ 1e9 by comparing them with each other (both
carry ~kappa eps of forward error); against this oracle each one's own error is visible.
"""
import numpy as np

LD = np.longdouble


def cholesky_ld(H):
    """Lower Cholesky factor of a small SPD matrix in longdouble (column by column)."""
    H = np.array(H, dtype=LD)
    d = H.shape[0]
    L = np.zeros((d, d), dtype=LD)
    for j in range(d):
        v = H[j:, j] - L[j:, :j] @ L[j, :j]
        L[j:, j] = v / np.sqrt(v[0])
    return L


def solve_lower_ld(L, B, transpose=False):
    """L^{-1} B or L^{-T} B by substitution in longdouble."""
    d = L.shape[0]
    X = np.array(B, dtype=LD).reshape(d, -1).copy()
    if not transpose:
        for i in range(d):
            X[i] = (X[i] - L[i, :i] @ X[:i]) / L[i, i]
    else:
        for i in range(d - 1, -1, -1):
            X[i] = (X[i] - L[i + 1:, i] @ X[i + 1:]) / L[i, i]
    return X


def logistic_ij_ld(X, y, theta, w):
    """(H, S = -H^{-1} G^T) of the weighted logistic objective in longdouble from float64 inputs."""
    X, y, theta, w = (np.asarray(a, dtype=LD) for a in (X, y, theta, w))
    z = X @ theta
    p = 1 / (1 + np.exp(-z))
    s = w * p * (1 - p)
    H = X.T @ (s[:, None] * X)
    G = (p - y)[:, None] * X                       # d^2 f / d w_n d theta = (p_n - y_n) x_n
    L = cholesky_ld(H)
    S = -solve_lower_ld(L, solve_lower_ld(L, G.T), transpose=True)
    return H, S
