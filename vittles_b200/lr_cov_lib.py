"""Linear-response covariances - the drop-in for ``vittles/lr_cov_lib.py``.

``Cov_LR(g1, g2) = J1 H^{-1} J2^T``: the Hessian is factorised by the GPU
Cholesky (``vt_potrf``), ``H^{-1} J2^T`` is a multi-right-hand-side ``vt_potrs``
and the final product runs on the FP64 tensor-core GEMM engine.

``factorize_hessian=False`` - documented upstream (``lr_cov_lib.py:67-70``) but
never implemented there (``:106`` always factorises) - is honoured here
(SURVEY.md section 8f item 2): the Hessian is not formed and every solve is a
conjugate-gradient run (``vt_cg_*`` kernels) over Hessian-vector products.
"""
from copy import deepcopy

import torch
from torch import func as tf

from . import ops, solver_lib
from ._arrays import to_device, kind_of, as_kind
from .objectives import StructuredObjective


class LinearResponseCovariances:
    """Reference: ``lr_cov_lib.py:9-216``.  ``objective_fun`` is a torch
    callable of the flat parameter (or a structured objective exposing
    ``vt_hessian`` / ``vt_grad``)."""

    #: options of the conjugate-gradient solver used when ``factorize_hessian=False``
    #: (see :func:`vittles_b200.solver_lib.get_cg_solver`); tighter than scipy's default
    #: 1e-5 because a covariance is a difference-sensitive quantity
    cg_opts = {'tol': 1e-10}

    def __init__(self, objective_fun, opt_par_value, validate_optimum=False, hessian_at_opt=None,
                 factorize_hessian=True, grad_tol=1e-8):
        self._obj_fun = objective_fun
        self._structured = isinstance(objective_fun, StructuredObjective)
        if self._structured:
            self._obj_fun_grad = objective_fun.vt_grad
            self._obj_fun_hessian = objective_fun.vt_hessian
        else:
            self._obj_fun_grad = tf.grad(objective_fun)
            self._obj_fun_hessian = tf.hessian(objective_fun)
        self._grad_tol = grad_tol
        self.set_base_values(opt_par_value, hessian_at_opt, factorize_hessian, validate=validate_optimum)

    def set_base_values(self, opt_par_value, hessian_at_opt, factorize_hessian=True, validate=True, grad_tol=None):
        """Reference ``:88-119``.  With ``factorize_hessian=True`` (default) the
        Hessian is formed (unless given) and factorised by the GPU Cholesky, as
        upstream (``:106``).  With ``False`` the solver is conjugate gradients
        over Hessian-vector products and no Hessian is formed (a supplied
        ``hessian_at_opt`` is used as the operator).  Validation is on the
        Newton step ``||H^{-1} grad||`` (``:108-119``)."""
        if grad_tol is None:
            grad_tol = self._grad_tol
        self._kind = kind_of(opt_par_value)
        self._opt0 = to_device(deepcopy(opt_par_value)).reshape(-1)
        self._hess_kind = self._kind if hessian_at_opt is None else None
        if factorize_hessian:
            self._hess0 = self._obj_fun_hessian(self._opt0) if hessian_at_opt is None else hessian_at_opt
            self.hess_solver = solver_lib.get_cholesky_solver(self._hess0)
        else:
            self._hess0 = hessian_at_opt                 # None: formed only if get_hessian_at_opt() is called
            self.hess_solver = solver_lib.get_cg_solver(self._hvp_operator(hessian_at_opt), len(self._opt0),
                                                        cg_opts=dict(self.cg_opts))
        if validate:
            grad0 = self._obj_fun_grad(self._opt0)
            newton_step = -1 * to_device(self.hess_solver(grad0), self._opt0.device)
            newton_step_norm = float(torch.linalg.vector_norm(newton_step))
            if newton_step_norm > grad_tol:
                raise ValueError(
                    'The gradient is not zero at the proposed optimal values.  '
                    '||newton_step|| = {} > {} = grad_tol'.format(newton_step_norm, grad_tol))

    def _hvp_operator(self, hessian_at_opt):
        """``v -> H v`` on device tensors without forming H: the structured
        objective's fused pass, a GEMV with a supplied Hessian, or forward-over-
        reverse autodiff of the generic objective."""
        dev = self._opt0.device
        if hessian_at_opt is not None:
            H = to_device(hessian_at_opt, dev).contiguous()
            return lambda v: ops.gemv(H, to_device(v, dev))
        if self._structured and hasattr(self._obj_fun, 'vt_hvp_fn'):
            return self._obj_fun.vt_hvp_fn(self._opt0)
        x0 = self._opt0
        return lambda v: tf.jvp(self._obj_fun_grad, (x0,), (to_device(v, dev),))[1]

    def get_hessian_at_opt(self):
        if self._hess0 is None:                           # factorize_hessian=False: formed on request only
            self._hess0 = self._obj_fun_hessian(self._opt0)
        return self._hess0 if self._hess_kind is None else as_kind(self._hess0, self._hess_kind)

    def get_lr_covariance_from_jacobians(self, moment_jacobian1, moment_jacobian2):
        """``J1 @ solve(J2^T)`` (reference ``:124-172``) with the same four
        ``ValueError`` checks (``:152-170``)."""
        if moment_jacobian1.ndim != 2:
            raise ValueError('moment_jacobian1 must be a 2d array.')
        if moment_jacobian2.ndim != 2:
            raise ValueError('moment_jacobian2 must be a 2d array.')
        dim = len(self._opt0)
        if moment_jacobian1.shape[1] != dim:
            raise ValueError(('The number of rows of moment_jacobian1 must match the dimension of the '
                              'optimization parameter. Expected {} rows, but got shape = {}').format(
                                  dim, tuple(moment_jacobian1.shape)))
        if moment_jacobian2.shape[1] != dim:
            raise ValueError(('The number of rows of moment_jacobian2 must match the dimension of the '
                              'optimization parameter. Expected {} rows, but got shape = {}').format(
                                  dim, tuple(moment_jacobian2.shape)))
        kind = kind_of(moment_jacobian1)
        dev = self._opt0.device
        j1 = to_device(moment_jacobian1, dev).contiguous()
        j2 = to_device(moment_jacobian2, dev)
        solved = to_device(self.hess_solver(j2.T.contiguous()), dev)        # (dim, K2) = H^{-1} J2^T
        return as_kind(ops.gemm(j1, solved, 'KC', 'KS'), kind)               # (K1, K2)

    def get_moment_jacobian(self, calculate_moments):
        """Jacobian of the moment map at the optimum (reference ``:174-192``);
        generic autodiff, off the hot path."""
        return as_kind(tf.jacrev(calculate_moments)(self._opt0), self._kind)

    def get_lr_covariance(self, calculate_moments):
        """Reference ``:194-216``."""
        moment_jacobian = self.get_moment_jacobian(calculate_moments)
        return self.get_lr_covariance_from_jacobians(moment_jacobian, moment_jacobian)
