"""Oracle restatement of ``vittles/lr_cov_lib.py``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np
import torch
from torch import func as tf

from . import solver_lib


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


def base_values(objective, opt_par, hessian=None, validate=False, grad_tol=1e-8):
    """``set_base_values`` (``lr_cov_lib.py:88-119``): H = hessian(f)(opt)
    unless given (``:101-104``); Cholesky solver (``:106``) - the
    ``factorize_hessian`` flag is accepted and ignored upstream; validation is
    on the NEWTON STEP norm ``||H^{-1} grad||`` (``:108-119``)."""
    opt_par = np.asarray(opt_par, dtype=np.float64)
    if hessian is None:
        hessian = tf.hessian(objective)(_t(opt_par)).detach().numpy().copy()
    solve = solver_lib.get_cholesky_solver(hessian)
    if validate:
        g = tf.grad(objective)(_t(opt_par)).detach().numpy()
        step = -1 * solve(g)
        nrm = np.linalg.norm(step)
        if nrm > grad_tol:
            raise ValueError(
                'The gradient is not zero at the proposed optimal values.  '
                '||newton_step|| = {} > {} = grad_tol'.format(nrm, grad_tol))
    return hessian, solve


def lr_covariance_from_jacobians(solve, dim, j1, j2):
    """``get_lr_covariance_from_jacobians`` (``lr_cov_lib.py:124-172``): four
    ``ValueError`` checks (``:152-170``), then ``J1 @ solve(J2.T)`` (``:172``)."""
    if j1.ndim != 2:
        raise ValueError('moment_jacobian1 must be a 2d array.')
    if j2.ndim != 2:
        raise ValueError('moment_jacobian2 must be a 2d array.')
    if j1.shape[1] != dim:
        raise ValueError('The number of rows of moment_jacobian1 must match the dimension '
                         'of the optimization parameter.')
    if j2.shape[1] != dim:
        raise ValueError('The number of rows of moment_jacobian2 must match the dimension '
                         'of the optimization parameter.')
    return j1 @ solve(j2.T)


def moment_jacobian(moments, opt_par):
    """``get_moment_jacobian`` (``lr_cov_lib.py:174-192``)."""
    return tf.jacrev(moments)(_t(opt_par)).detach().numpy().copy()


def lr_covariance(objective, opt_par, moments, hessian=None):
    """``get_lr_covariance`` (``lr_cov_lib.py:194-216``)."""
    _, solve = base_values(objective, opt_par, hessian)
    j = moment_jacobian(moments, opt_par)
    return lr_covariance_from_jacobians(solve, len(opt_par), j, j)
