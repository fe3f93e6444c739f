"""Sensitivity of optima to hyperparameters - the drop-in for
``vittles/sensitivity_lib.py`` with the arithmetic on a B200.

Same class names, constructor keywords, methods and exceptions as the reference
(SURVEY.md section 8b).  Two evaluation paths, chosen by the TYPE of the
objective, never silently:

* a :class:`vittles_b200.objectives.StructuredObjective` (GLM families ...):
  Hessian, cross-Hessian apply and directional derivatives run in the fused
  sm_100a kernels; the (D, M) cross-Hessian is never materialised;
* any other torch callable ``f(theta, lam) -> scalar``: derivatives come from
  ``torch.func`` on the GPU ("off the hot path") and only the linear algebra -
  Cholesky factorisation, multi-RHS solve, prediction GEMV - uses the kernels.
"""
import warnings
from copy import deepcopy
from math import factorial

import numpy as np
import torch
from torch import func as tf

from . import ops, solver_lib
from ._arrays import to_device, kind_of, as_kind, default_device
from .objectives import StructuredObjective


# ---------------------------------------------------------------------------
# A differentiable function with a prescribed value and Jacobian
# ---------------------------------------------------------------------------

def get_linear_function(return_val0, arg_val0, dreturn_darg):
    """Differentiable ``arg -> return_val0`` with Jacobian ``dreturn_darg``
    (reference: ``sensitivity_lib.py:18-82``).

    The returned function must be evaluated at ``arg_val0`` (max-abs tolerance
    1e-8, ``ValueError`` otherwise, ``:38-42``).  Its reverse-mode derivative is
    ``g^T J`` (``:66-67``) and its forward-mode derivative ``J g`` (``:74-76``),
    both run by the GEMV kernel; second and higher derivatives raise
    ``NotImplementedError`` on purpose (``:63-65,71-72``)."""
    ret0 = to_device(return_val0)
    arg0 = to_device(arg_val0, ret0.device)
    jac = to_device(dreturn_darg, ret0.device)
    jac = jac if jac.stride(-1) == 1 else jac.contiguous()

    def check_arg(arg, tolerance=1e-8):
        a = arg.detach() if isinstance(arg, torch.Tensor) else arg
        if float(torch.max(torch.abs(to_device(a, ret0.device) - arg0))) > tolerance:
            raise ValueError('get_return_par must be evaluated at ', arg0, ' != ', arg)

    class _NoHigherOrder(torch.autograd.Function):
        """g -> g^T J, itself non-differentiable."""
        @staticmethod
        def forward(ctx, g):
            # g^T J as a (1 x N) x (N x M) product on the GEMM engine: no transposed copy of J
            return ops.gemm(g.reshape(1, -1).contiguous(), jac, 'KC', 'KS').reshape(-1)

        @staticmethod
        def backward(ctx, gg):
            raise NotImplementedError('second-order derivatives of the optimum are not available from a '
                                      'first-order approximation')

    class _LinearFunction(torch.autograd.Function):
        @staticmethod
        def forward(arg):
            return ret0.clone()

        @staticmethod
        def setup_context(ctx, inputs, output):
            pass

        @staticmethod
        def backward(ctx, g):
            return _NoHigherOrder.apply(g)

        @staticmethod
        def jvp(ctx, tangent):
            if isinstance(tangent, torch.Tensor) and tangent.requires_grad:
                raise NotImplementedError('second-order derivatives of the optimum are not available')
            return ops.gemv(jac, tangent.contiguous())

    def get_return_par(arg_par):
        check_arg(arg_par)
        if isinstance(arg_par, torch.Tensor):
            return _LinearFunction.apply(arg_par.to(device=ret0.device, dtype=torch.float64))
        return as_kind(ret0.clone(), kind_of(arg_par))
    return get_return_par


# ---------------------------------------------------------------------------
# Linear approximation
# ---------------------------------------------------------------------------

class EstimatingEquationLinearApproximation:
    """Linear dependence of the root of an estimating equation on a
    hyperparameter (reference: ``sensitivity_lib.py:85-254``).

    ``estimating_equation(input_par, hyper_par) -> (N,)`` is a torch callable;
    ``hess_solver(v)`` applies the inverse Jacobian with respect to
    ``input_par`` to a vector or to the columns of a matrix."""

    def __init__(self, estimating_equation, input_val0, hyper_val0, hess_solver,
                 validate_solution=False, estimating_equation_jac0=None,
                 hyper_par_estimating_equation=None, solution_tol=1e-8):
        self._estimating_equation = estimating_equation
        self._hyper_par_estimating_equation = \
            estimating_equation if hyper_par_estimating_equation is None else hyper_par_estimating_equation
        self._hess_solver = hess_solver
        self._solution_tol = solution_tol
        self.set_location(input_val0, hyper_val0, estimating_equation_jac0,
                          validate_solution=validate_solution, solution_tol=solution_tol)

    # -- hooks overridden by the structured path --------------------------
    def _home_device(self, value=None):
        """Where the parameters live: a structured objective's own GPU, else the device of a CUDA input, else the
        current device."""
        obj = getattr(self, '_objective_fun', None)
        dev = getattr(obj, 'device', None)
        if dev is None and isinstance(value, torch.Tensor) and value.is_cuda:
            dev = value.device
        return dev

    def _eval_estimating_equation(self, x, h):
        return self._estimating_equation(x, h)

    def _eval_cross_jacobian(self, x, h):
        """d estimating_equation / d hyper, shape (N, M) (reference ``:180-181,217-219``)."""
        return tf.jacrev(self._hyper_par_estimating_equation, argnums=1)(x, h)

    def set_location(self, input_val0, hyper_val0, estimating_equation_jac0,
                     validate_solution=True, solution_tol=None):
        """Reference: ``sensitivity_lib.py:192-226``."""
        self._kind = kind_of(input_val0)
        self._hyper_kind = kind_of(hyper_val0)
        self._input_val0 = to_device(deepcopy(input_val0), self._home_device(input_val0)).reshape(-1)
        self._hyper_val0 = to_device(deepcopy(hyper_val0), self._input_val0.device).reshape(-1)

        if validate_solution:
            if solution_tol is None:
                solution_tol = self._solution_tol
            ee_val = self._eval_estimating_equation(self._input_val0, self._hyper_val0)
            ee_norm = float(torch.linalg.vector_norm(to_device(ee_val)))
            if ee_norm > solution_tol:
                raise ValueError(
                    'The estimating equation is not zero at the proposed  values.  '
                    '||ee|| = {} > {} = solution_tol'.format(ee_norm, solution_tol))

        self._set_sens_mat(estimating_equation_jac0)

    def _set_sens_mat(self, estimating_equation_jac0):
        if estimating_equation_jac0 is None:
            jac0 = self._eval_cross_jacobian(self._input_val0, self._hyper_val0)
        else:
            jac0 = to_device(estimating_equation_jac0, self._input_val0.device)
        self._estimating_equation_jac0 = jac0
        if tuple(jac0.shape) != (len(self._input_val0), len(self._hyper_val0)):
            raise ValueError('``_estimating_equation_jac0`` is the wrong shape.')
        self._sens_mat = -1 * to_device(self._hess_solver(jac0), self._input_val0.device)

    def get_dinput_dhyper(self):
        """The (N, M) sensitivity matrix, in the array kind of ``hyper_val0``
        (a CUDA tensor stays on the device; it is returned by reference, as
        upstream ``:230-231``)."""
        return as_kind(self._sens_mat, self._hyper_kind)

    def hess_solver(self):
        return self._hess_solver

    def _predict_device(self, new_hyper):
        delta = to_device(new_hyper, self._input_val0.device).reshape(-1) - self._hyper_val0
        sm = self._sens_mat if self._sens_mat.stride(1) == 1 else self._sens_mat.contiguous()
        return ops.gemv(sm, delta, alpha=1.0, y0=self._input_val0, beta=1.0)

    def predict_input_par_from_hyper_par(self, new_hyper_par_value):
        """``input_val0 + sens_mat @ (new_hyper - hyper_val0)`` (reference
        ``:236-247``), one streaming pass over the sensitivity matrix."""
        return as_kind(self._predict_device(new_hyper_par_value), kind_of(new_hyper_par_value))

    def get_input_par_function(self):
        """Reference ``:250-254``."""
        return get_linear_function(self._input_val0, self._hyper_val0, self._sens_mat)


# kappa(H) above which the fused IJ path stops multiplying by an explicit inverse: kappa * 2^-53 ~ 1e-10 leaves two
# orders of magnitude to the rtol 1e-8 parity bar (tests/test_gpu_ij.py::test_conditioning_sweep)
EXPLICIT_INVERSE_MAX_COND = 1e6


class HyperparameterSensitivityLinearApproximation(EstimatingEquationLinearApproximation):
    """Linear dependence of an optimum on a hyperparameter:
    d theta_hat / d lambda = -H^{-1} d^2 f / d theta d lambda
    (reference: ``sensitivity_lib.py:258-429``).

    With a :class:`~vittles_b200.objectives.GLMObjective` (hyperparameter :=
    per-observation weights) this is the all-observation infinitesimal
    jackknife: H = X^T diag(s) X is assembled by the FP64 tensor-core SYRK,
    factorised on the GPU, and the (D, N) result is produced by one fused GEMM
    that never materialises the per-observation gradient matrix."""

    def __init__(self, objective_fun, opt_par_value, hyper_par_value,
                 validate_optimum=False, hessian_at_opt=None, cross_hess_at_opt=None,
                 hyper_par_objective_fun=None, grad_tol=1e-8):
        self._objective_fun = objective_fun
        self._structured = isinstance(objective_fun, StructuredObjective)
        if hyper_par_objective_fun is None:
            hyper_par_objective_fun = objective_fun
        if self._structured:
            obj_fun_grad = objective_fun.vt_grad
            hyper_obj_fun_grad = objective_fun.vt_grad
        else:
            obj_fun_grad = tf.grad(objective_fun, argnums=0)
            hyper_obj_fun_grad = tf.grad(hyper_par_objective_fun, argnums=0)

        hess_solver = self._get_hessian_solver(opt_par_value, hyper_par_value, hessian_at_opt)

        EstimatingEquationLinearApproximation.__init__(
            self,
            estimating_equation=obj_fun_grad,
            input_val0=opt_par_value,
            hyper_val0=hyper_par_value,
            hess_solver=hess_solver,
            validate_solution=validate_optimum,
            estimating_equation_jac0=cross_hess_at_opt,
            hyper_par_estimating_equation=hyper_obj_fun_grad,
            solution_tol=grad_tol)

    def _get_hessian_solver(self, opt_par_value, hyper_par_value, hessian_at_opt):
        """Reference ``:376-390``: Hessian by differentiation unless supplied,
        shape check, then ALWAYS the Cholesky solver."""
        self._stats = None
        if hessian_at_opt is None:
            theta = to_device(opt_par_value, self._home_device(opt_par_value)).reshape(-1)
            lam = to_device(hyper_par_value, theta.device).reshape(-1)
            if self._structured and hasattr(self._objective_fun, 'vt_stats_and_hessian'):
                self._stats, self._hess0 = self._objective_fun.vt_stats_and_hessian(theta, lam)
            elif self._structured:
                self._stats = self._objective_fun.vt_stats(theta, lam)
                self._hess0 = self._objective_fun.vt_hessian(theta, lam, self._stats)
            else:
                self._hess0 = tf.hessian(self._objective_fun, argnums=0)(theta, lam)
            self._hess_kind = kind_of(opt_par_value)
        else:
            self._hess0 = hessian_at_opt
            self._hess_kind = None
        if tuple(self._hess0.shape) != (len(opt_par_value), len(opt_par_value)):
            raise ValueError('``hessian_at_opt`` is the wrong shape.')
        return solver_lib.get_cholesky_solver(self._hess0)

    def _eval_estimating_equation(self, x, h):
        if self._structured and self._stats is not None:
            return self._stats['grad']
        return self._estimating_equation(x, h)

    def _set_sens_mat(self, estimating_equation_jac0):
        is_ij = self._structured and estimating_equation_jac0 is None and \
            hasattr(self._objective_fun, 'vt_ij_sensitivity') and \
            len(self._hyper_val0) == self._objective_fun.n_obs
        if not is_ij:
            if self._structured and estimating_equation_jac0 is None:
                raise NotImplementedError(
                    'the structured linear approximation is implemented for hyperparameter := observation weights; '
                    'pass `cross_hess_at_opt` or use ParametricSensitivityTaylorExpansion for other hyperparameters')
            return super()._set_sens_mat(estimating_equation_jac0)
        # fused infinitesimal-jackknife path: S = -H^{-1} G^T, G^T never formed
        if self._stats is None:
            self._stats = self._objective_fun.vt_stats(self._input_val0, self._hyper_val0, want_grad=False)
        factor = getattr(self._hess_solver, 'factor', None)
        if factor is None:
            factor = ops.potrf(to_device(self._hess0, self._input_val0.device))
        self._estimating_equation_jac0 = None
        # The fused path multiplies by an explicit H^{-1} (one GEMM); the reference substitutes with the Cholesky
        # factor (``solver_lib.py:29``).  The two agree to ~kappa(H) eps, so the explicit inverse is used only while
        # a lower bound on kappa(H) keeps that below the parity tolerance; otherwise substitution, like upstream.
        self.hessian_cond_lower_bound = factor.cond_lower_bound()
        self.used_explicit_inverse = self.hessian_cond_lower_bound <= EXPLICIT_INVERSE_MAX_COND or \
            not hasattr(self._objective_fun, 'vt_ij_sensitivity_by_substitution')
        if self.used_explicit_inverse:
            self._hinv = factor.inverse()
            self._sens_mat = self._objective_fun.vt_ij_sensitivity(self._hinv, self._stats)
        else:
            self._hinv = None
            self._sens_mat = self._objective_fun.vt_ij_sensitivity_by_substitution(factor, self._stats)

    def set_base_values(self, opt_par_value, hyper_par_value, hessian_at_opt, cross_hess_at_opt,
                        validate_optimum=True, grad_tol=None):
        """Reference ``:392-405``."""
        self._hess_solver = self._get_hessian_solver(opt_par_value, hyper_par_value, hessian_at_opt)
        EstimatingEquationLinearApproximation.set_location(
            self, input_val0=opt_par_value, hyper_val0=hyper_par_value,
            estimating_equation_jac0=cross_hess_at_opt,
            validate_solution=validate_optimum, solution_tol=grad_tol)

    def get_dopt_dhyper(self):
        return super().get_dinput_dhyper()

    def get_hessian_at_opt(self):
        if self._hess_kind is None:
            return self._hess0
        return as_kind(self._hess0, self._hess_kind)

    def predict_opt_par_from_hyper_par(self, new_hyper_par_value):
        if self._structured and self._objective_fun.group is not None:
            # each rank holds the columns of its own observations
            delta = to_device(new_hyper_par_value, self._input_val0.device).reshape(-1) - self._hyper_val0
            part = ops.gemv(self._sens_mat, delta, alpha=1.0)
            self._objective_fun._allreduce(part)
            return as_kind(self._input_val0 + part, kind_of(new_hyper_par_value))
        return super().predict_input_par_from_hyper_par(new_hyper_par_value)

    def get_opt_par_function(self):
        return super().get_input_par_function()


# ---------------------------------------------------------------------------
# Higher-order directional derivatives
# ---------------------------------------------------------------------------

def _append_jvp(fun, num_base_args=1, argnum=0):
    """Append one forward-mode direction to ``fun`` (reference:
    ``sensitivity_lib.py:440-492``): the returned function takes the base
    arguments, the directions already appended, and one more direction ``v``,
    and returns d/d x_argnum [fun(..., previous directions)] . v."""
    assert argnum < num_base_args

    def with_direction(*argv):
        base, vecs = list(argv[:num_base_args]), list(argv[num_base_args:])
        v, earlier = vecs[-1], vecs[:-1]

        def of_arg(x):
            args = list(base)
            args[argnum] = x
            return fun(*args, *earlier)
        return tf.jvp(of_arg, (base[argnum],), (v,))[1]
    return with_direction


class DerivativeTerm:
    """One term of d^k/d eps^k g(eta(eps), eps): ``prefactor`` times the partial
    derivative of g of order ``eps_order`` in eps and ``sum(eta_orders)`` in eta,
    contracted with ``eta_orders[i]`` copies of d^{i+1} eta / d eps^{i+1}
    (reference: ``sensitivity_lib.py:495-688``)."""

    def __init__(self, eps_order, eta_orders, prefactor):
        self.eps_order = eps_order
        self.eta_orders = eta_orders
        self.prefactor = prefactor
        self.total_eta_order = int(np.sum(self.eta_orders))
        self._order = int(self.eps_order + sum((i + 1) * c for i, c in enumerate(self.eta_orders)))
        assert isinstance(self.eps_order, int) and self.eps_order >= 0
        assert len(self.eta_orders) == self._order
        for c in self.eta_orders:
            assert isinstance(c, int) and c >= 0

    def __str__(self):
        return 'Order: {}\t{} * eta{} * eps[{}]'.format(self._order, self.prefactor, self.eta_orders, self.eps_order)

    def order(self):
        return self._order

    def differentiate(self):
        """Product and chain rule in eps (reference ``:638-673``): one term from
        d/d eps of the partial of g, one from d/d eta of it (times eta'), and one
        per distinct eta-derivative factor that gets differentiated."""
        grown = list(self.eta_orders) + [0]
        out = [DerivativeTerm(self.eps_order + 1, list(grown), self.prefactor)]
        bumped = list(grown)
        bumped[0] += 1
        out.append(DerivativeTerm(self.eps_order, bumped, self.prefactor))
        for i, count in enumerate(self.eta_orders):
            if count > 0:
                moved = list(grown)
                moved[i] -= 1
                moved[i + 1] += 1
                out.append(DerivativeTerm(self.eps_order, moved, self.prefactor * count))
        return out

    def check_similarity(self, term):
        return (self.eps_order == term.eps_order) & (self.eta_orders == term.eta_orders)

    def combine_with(self, term):
        assert self.check_similarity(term)
        return DerivativeTerm(self.eps_order, self.eta_orders, self.prefactor + term.prefactor)


def _consolidate_terms(dterms):
    """Merge terms with the same derivative signature (reference ``:980-1004``).
    First-appearance order is kept; every like term is merged."""
    merged = {}
    for t in dterms:
        key = (t.eps_order, tuple(t.eta_orders))
        merged[key] = merged[key].combine_with(t) if key in merged else t
    return list(merged.values())


def _get_taylor_base_terms():
    """dg/deps and dg/deta * eta'  (reference ``:1008-1018``)."""
    return [DerivativeTerm(1, [0], 1.0), DerivativeTerm(0, [1], 1.0)]


def _evaluate_term_fwd(term, eta0, eps0, deps, eta_derivs, eval_directional_derivative, validate=False):
    """Evaluate one term in forward mode (reference ``:691-734``): the eta
    directions are ``eta_orders[i]`` copies of ``eta_derivs[i]``, followed by
    ``eps_order`` copies of ``deps``."""
    if validate and len(eta_derivs) < term.order() - 1:
        raise ValueError('Not enough derivatives in ``eta_derivs``.')
    eta_directions = []
    for i, count in enumerate(term.eta_orders):
        eta_directions += [eta_derivs[i]] * count if count > 0 else []
    eps_directions = [deps] * term.eps_order
    return term.prefactor * eval_directional_derivative(eta0, eps0, eta_directions, eps_directions, validate=validate)


class ForwardModeDerivativeArray:
    """Directional partial derivatives of ``fun(x1, x2)`` of order up to
    ``(order1, order2)`` (reference: ``sensitivity_lib.py:766-807``).

    For a generic torch callable the table of nested forward-mode JVPs is built
    like the reference's.  If ``fun`` carries a ``vt_directional_derivative``
    hook (a structured estimating equation) every derivative is ONE fused pass
    over the design matrix instead of a tower of JVPs."""

    def __init__(self, fun, order1, order2):
        self._order1 = order1
        self._order2 = order2
        self._fun = fun
        self._hook = getattr(fun, 'vt_directional_derivative', None)
        self._cache = {}
        if self._hook is None:
            table = [[fun]]
            for i1 in range(order1 + 1):
                if i1 > 0:
                    table.append([_append_jvp(table[i1 - 1][0], num_base_args=2, argnum=0)])
                for i2 in range(order2):
                    table[i1].append(_append_jvp(table[i1][i2], num_base_args=2, argnum=1))
            self._eval_fun_derivs = table

    def eval_directional_derivative(self, x1, x2, dx1s, dx2s, validate=True):
        order1, order2 = len(dx1s), len(dx2s)
        if validate:
            if order1 > self._order1:
                raise ValueError('The number of `dx1s` ({}) must be <= order1 = {}'.format(order1, self._order1))
            if order2 > self._order2:
                raise ValueError('The number of `dx2s` ({}) must be <= order2 = {}'.format(order2, self._order2))
            if (not isinstance(dx1s, list)) or (not isinstance(dx2s, list)):
                raise ValueError('`dx1s` and `dx2s` must be lists of vectors.')
        if self._hook is not None:
            return self._hook(x1, x2, dx1s, dx2s, cache=self._cache)
        return self._eval_fun_derivs[order1][order2](x1, x2, *[*dx1s, *dx2s])


def _contract_tensor(deriv_array, dx1s, dx2s):
    """Contract every axis after the first with the given vectors, x1
    directions first (reference ``:737-763``)."""
    vecs = list(dx1s) + list(dx2s)
    if len(vecs) >= 26:
        raise ValueError('You cannot use _contract_tensor with so many vectors.')
    out = deriv_array
    for v in reversed(vecs):
        out = torch.tensordot(out, v, dims=([out.dim() - 1], [0]))
    return out


class ReverseModeDerivativeArray:
    """Dense partial-derivative tensors of ``fun(x1, x2)`` by repeated
    reverse-mode Jacobians, contracted on demand (reference:
    ``sensitivity_lib.py:810-918``; flagged experimental upstream).  Only for
    low-dimensional problems: the same 100 000-element guard applies."""

    def __init__(self, fun, order1, order2):
        self._order1, self._order2 = order1, order2
        table = [[fun]]
        for i1 in range(order1 + 1):
            if i1 > 0:
                table.append([tf.jacrev(table[i1 - 1][0], argnums=0)])
            for i2 in range(order2):
                table[i1].append(tf.jacrev(table[i1][i2], argnums=1))
        self._eval_deriv_arrays = table

    def set_evaluation_location(self, x1, x2, force=False, verbose=False):
        x1 = torch.atleast_1d(to_device(x1))
        x2 = torch.atleast_1d(to_device(x2, x1.device))
        if x1.dim() != 1 or x2.dim() != 1:
            raise ValueError('x1 and x2 must be 1d arrays.')
        base = torch.atleast_1d(self._eval_deriv_arrays[0][0](x1, x2))
        dim0, dim1, dim2 = len(base), len(x1), len(x2)
        total_size = sum(dim0 * dim1 ** i1 * dim2 ** i2
                         for i1 in range(self._order1 + 1) for i2 in range(self._order2 + 1))
        max_allowed_size = 100000
        if total_size > max_allowed_size and not force:
            raise ValueError(
                ('With len(x1) = {}, len(x2) = {}, order1 = {}, and order2 = {}, this will create a partial '
                 'derivative array of size {} > {}.  To force the creation of these arrays, set`force=True`.'
                 ).format(dim1, dim2, self._order1, self._order2, total_size, max_allowed_size))
        if self._order1 > 2 and self._order2 > 2 and not force:
            raise ValueError('With both orders greater than two, reverse mode can be slow even in '
                             'low-dimensional problems.  To force the creation of the arrays, set force=True.')
        self._x1, self._x2 = x1.clone(), x2.clone()
        if base.dim() != 1:
            raise ValueError('The base function is expected to evaluate to a 1d vector.')
        self._deriv_arrays = [[None] * (self._order2 + 1) for _ in range(self._order1 + 1)]
        for i1 in range(self._order1 + 1):
            for i2 in range(self._order2 + 1):
                if verbose:
                    print('Evaluating the derivative {}, {}'.format(i1, i2))
                self._deriv_arrays[i1][i2] = self._eval_deriv_arrays[i1][i2](x1, x2)

    def deriv_arrays(self, order1, order2):
        return self._deriv_arrays[order1][order2]

    def _check_location(self, x1, x2, tol=1e-8):
        x1, x2 = to_device(x1, self._x1.device), to_device(x2, self._x1.device)
        if float(torch.max(torch.abs(x1 - self._x1))) > tol or float(torch.max(torch.abs(x2 - self._x2))) > tol:
            raise ValueError('You must use the x1 and x2 set in `set_evaluation_location`.')

    def eval_directional_derivative(self, x1, x2, dx1s, dx2s, validate=True):
        order1, order2 = len(dx1s), len(dx2s)
        if validate:
            self._check_location(x1, x2)
            if order1 > self._order1:
                raise ValueError('The number of `dx1s` ({}) must be <= order1 = {}'.format(order1, self._order1))
            if order2 > self._order2:
                raise ValueError('The number of `dx2s` ({}) must be <= order2 = {}'.format(order2, self._order2))
        dev = self._x1.device
        return _contract_tensor(self._deriv_arrays[order1][order2],
                                [to_device(v, dev) for v in dx1s], [to_device(v, dev) for v in dx2s])


class ReorderedReverseModeDerivativeArray:
    """Reverse-mode arrays that differentiate with respect to the larger
    argument last (reference: ``sensitivity_lib.py:921-977``)."""

    def __init__(self, fun, order1, order2, swapped=None):
        self._swapped = swapped
        self._orderz1, self._orderz2 = (order2, order1) if swapped else (order1, order2)
        self._fun = lambda z1, z2: fun(*self._swapped_args(z1, z2))
        self._rmda = ReverseModeDerivativeArray(self._fun, self._orderz1, self._orderz2)

    def _swapped_args(self, x1, x2):
        return (x2, x1) if self._swapped else (x1, x2)

    def set_evaluation_location(self, x1, x2, force=False, verbose=False):
        z1, z2 = self._swapped_args(x1, x2)
        return self._rmda.set_evaluation_location(x1=z1, x2=z2, force=force, verbose=verbose)

    def eval_directional_derivative(self, x1, x2, dx1s, dx2s, validate=True):
        z1, z2 = self._swapped_args(x1, x2)
        dz1s, dz2s = self._swapped_args(dx1s, dx2s)
        return self._rmda.eval_directional_derivative(z1, z2, dz1s, dz2s, validate=validate)

    def deriv_arrays(self, order1, order2):
        if self._swapped:
            arr = self._rmda.deriv_arrays(order2, order1)
            axes2 = [a + 1 for a in range(order2)]
            return torch.movedim(arr, axes2, [-a for a in axes2])
        return self._rmda.deriv_arrays(order1, order2)


# ---------------------------------------------------------------------------
# Taylor expansion of the optimum
# ---------------------------------------------------------------------------

class ParametricSensitivityTaylorExpansion(object):
    """Taylor series of eta_hat(eps) solving g(eta, eps) = 0 (reference:
    ``sensitivity_lib.py:1021-1363``).

    The term tables (Faa di Bruno bookkeeping) are built on the host exactly as
    upstream and used as a launch schedule: each term is one directional
    derivative of g, each order ends with one ``hess_solver`` call.  With a
    structured objective every directional derivative is a single fused pass
    over the design matrix and ``hess_solver`` may be ``get_cg_solver`` over the
    fused Hessian-vector product."""

    @classmethod
    def optimization_objective(cls, objective_function, input_val0, hyper_val0, order, hess0=None,
                               forward_mode=True, max_input_order=None, max_hyper_order=None, force=False):
        """Reference ``:1032-1081``: estimating equation = gradient of the
        objective; Hessian by differentiation unless ``hess0`` is given; always
        the Cholesky solver."""
        if isinstance(objective_function, StructuredObjective):
            estimating_equation = StructuredEstimatingEquation(objective_function)
            if hess0 is None:
                hess0 = objective_function.vt_hessian(to_device(input_val0), to_device(hyper_val0))
        else:
            estimating_equation = tf.grad(objective_function, argnums=0)
            if hess0 is None:
                x = to_device(input_val0).reshape(-1)
                hess0 = tf.hessian(objective_function, argnums=0)(x, to_device(hyper_val0, x.device).reshape(-1))
        hess_solver = solver_lib.get_cholesky_solver(hess0)
        return cls(estimating_equation=estimating_equation, input_val0=input_val0, hyper_val0=hyper_val0,
                   order=order, hess_solver=hess_solver, forward_mode=forward_mode,
                   max_input_order=max_input_order, max_hyper_order=max_hyper_order, force=force)

    def __init__(self, estimating_equation, input_val0, hyper_val0, order, hess_solver,
                 forward_mode=True, max_input_order=None, max_hyper_order=None, force=False):
        self._kind = kind_of(input_val0)
        self._input_val0 = to_device(deepcopy(input_val0)).reshape(-1)
        self._hyper_val0 = to_device(deepcopy(hyper_val0), self._input_val0.device).reshape(-1)
        if isinstance(estimating_equation, StructuredObjective):
            estimating_equation = StructuredEstimatingEquation(estimating_equation)
        self._objective_function_eta_grad = estimating_equation
        self._set_order(order, max_input_order, max_hyper_order, forward_mode)
        self.hess_solver = hess_solver
        if not self._forward_mode:
            self._deriv_array.set_evaluation_location(self._input_val0, self._hyper_val0, force=force)

    def _set_order(self, order, max_input_order, max_hyper_order, forward_mode):
        """Reference ``:1143-1203``."""
        self._max_input_order = max_input_order
        self._max_hyper_order = max_hyper_order
        self._forward_mode = forward_mode
        if not self._forward_mode:
            warnings.warn('Reverse mode Taylor expansions are experimental.')
        if self._max_input_order is not None or self._max_hyper_order is not None:
            warnings.warn('Setting _max_hyper_order or _max_input_order is experimental.')
        if self._max_input_order is not None and self._max_input_order < 1:
            raise ValueError('max_input_order must be >= 1.')
        if self._max_hyper_order is not None and self._max_hyper_order < 1:
            raise ValueError('max_hyper_order must be >= 1.')
        self._order = order
        order1 = self._order if self._max_input_order is None else min(self._order, self._max_input_order)
        order2 = self._order if self._max_hyper_order is None else min(self._order, self._max_hyper_order)
        if self._forward_mode:
            self._deriv_array = ForwardModeDerivativeArray(self._objective_function_eta_grad, order1, order2)
        else:
            swapped = len(self._input_val0) > len(self._hyper_val0)
            self._deriv_array = ReorderedReverseModeDerivativeArray(
                self._objective_function_eta_grad, order1, order2, swapped=swapped)
        self._taylor_terms_list = [_get_taylor_base_terms()]
        for k in range(1, self._order):
            derived = []
            for term in self._taylor_terms_list[k - 1]:
                derived += term.differentiate()
            self._taylor_terms_list.append(_consolidate_terms(derived))

    def get_max_order(self):
        return self._order

    def _evaluate_dkinput_dhyperk(self, dhyper, input_derivs, k):
        """d^k input / d hyper^k along ``dhyper`` (reference ``:1208-1260``)."""
        if k <= 0:
            raise ValueError('k must be at least one.')
        if k > self._order:
            raise ValueError('k must be no greater than the declared order={}'.format(self._order))
        if len(input_derivs) < k - 1:
            raise ValueError('Not enough eta_derivs provided.')
        vec = torch.zeros_like(self._input_val0)
        for term in self._taylor_terms_list[k - 1]:
            if term.eta_orders[-1] > 0:
                continue      # holds the unknown d^k eta itself
            if self._max_hyper_order is not None and term.eps_order > self._max_hyper_order:
                continue
            if self._max_input_order is not None and term.total_eta_order > self._max_input_order:
                continue
            vec = vec + _evaluate_term_fwd(
                term=term, eta0=self._input_val0, eps0=self._hyper_val0, deps=dhyper, eta_derivs=input_derivs,
                eval_directional_derivative=self._deriv_array.eval_directional_derivative)
        return -1 * to_device(self.hess_solver(vec), vec.device)

    def _get_default_max_order(self, max_order):
        if max_order is None:
            return self._order
        if max_order <= 0:
            raise ValueError('max_order must be greater than zero.')
        if max_order > self._order:
            raise ValueError('max_order must be no greater than the order={}'.format(self._order))
        return max_order

    def _input_derivs_device(self, dhyper, max_order):
        dhyper = to_device(dhyper, self._input_val0.device).reshape(-1)
        input_derivs = []
        for k in range(1, max_order + 1):
            input_derivs.append(self._evaluate_dkinput_dhyperk(dhyper=dhyper, input_derivs=input_derivs, k=k))
        return input_derivs

    def evaluate_input_derivs(self, dhyper, max_order=None):
        """[d^k input / d hyper^k dhyper^k for k = 1..max_order] (reference ``:1274-1286``)."""
        max_order = self._get_default_max_order(max_order)
        return [as_kind(d, self._kind) for d in self._input_derivs_device(dhyper, max_order)]

    def _terms_device(self, new_hyper_val, add_offset, max_order):
        max_order = self._get_default_max_order(max_order)
        dhyper = to_device(new_hyper_val, self._input_val0.device).reshape(-1) - self._hyper_val0
        derivs = self._input_derivs_device(dhyper, max_order)
        terms = [self._input_val0 if add_offset else torch.zeros_like(self._input_val0)]
        for k in range(1, max_order + 1):
            terms.append(derivs[k - 1] / float(factorial(k)))
        return terms

    def evaluate_taylor_series_terms(self, new_hyper_val, add_offset=True, max_order=None):
        """Reference ``:1289-1304``."""
        return [as_kind(t, self._kind) for t in self._terms_device(new_hyper_val, add_offset, max_order)]

    def evaluate_taylor_series(self, new_hyper_val, add_offset=True, max_order=None, sum_terms=True):
        """Reference ``:1307-1343`` (``sum_terms`` is accepted and, as upstream, ignored)."""
        terms = self._terms_device(new_hyper_val, add_offset, max_order)
        return as_kind(torch.sum(torch.stack(terms), dim=0), self._kind)

    def print_terms(self, k=None):
        """Reference ``:1346-1363``."""
        if k is not None and k > self._order:
            raise ValueError('k must be no greater than order={}'.format(self._order))
        for order in range(self._order):
            if k is None or order == (k - 1):
                print('\nTerms for order {}:'.format(order + 1))
                for term in self._taylor_terms_list[order]:
                    print(term)


class StructuredEstimatingEquation:
    """g = grad_theta f for a structured objective: callable like a generic
    estimating equation and carrying the fused directional-derivative hook."""

    def __init__(self, objective):
        self.objective = objective

    def __call__(self, eta, eps):
        return self.objective.vt_grad(eta, eps)

    def vt_directional_derivative(self, eta, eps, eta_dirs, eps_dirs, cache=None):
        return self.objective.vt_directional_derivative(eta, eps, eta_dirs, eps_dirs, cache=cache)
