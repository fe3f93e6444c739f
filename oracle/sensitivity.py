"""Oracle restatement of ``vittles/sensitivity_lib.py`` (hot-path parts).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Objectives are torch
callables ``f(theta, lam) -> scalar`` evaluated in float64 on the CPU;
derivatives come from ``torch.func`` where the reference used ``autograd``.
Inputs and outputs at this module's surface are numpy arrays.
"""
from math import factorial

import numpy as np
import torch
from torch import func as tf

from . import solver_lib


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


def _n(a):
    return a.detach().numpy().copy() if isinstance(a, torch.Tensor) else np.asarray(a)


# --------------------------------------------------------------------------
# Linear approximation  (sensitivity_lib.py:85-429)
# --------------------------------------------------------------------------

def linear_sensitivity(objective, theta0, lam0, hessian=None, cross_hessian=None,
                       hyper_objective=None, validate_optimum=False, grad_tol=1e-8,
                       hess_solver=None):
    """d theta_hat / d lambda = -H^{-1} d2f/dtheta dlambda.

    Follows ``HyperparameterSensitivityLinearApproximation.__init__``
    (``sensitivity_lib.py:303-374``): H = hessian(f, 0)(theta0, lam0) unless
    given (``:381-385``), shape check (``:386-387``), Cholesky solver (``:389``);
    then ``set_location`` (``:192-226``): optional ``||grad|| <= tol`` check
    (``:203-215``), cross term = jacobian(grad_theta f_hyper, 1) unless given
    (``:217-221``), shape check (``:222-224``), ``sens = -solve(cross)``
    (``:226``).  Returns a dict with ``hessian``, ``cross_hessian``, ``sens``
    and the ``solve`` closure.
    """
    theta0 = np.asarray(theta0, dtype=np.float64)
    lam0 = np.asarray(lam0, dtype=np.float64)
    grad_theta = tf.grad(objective, argnums=0)
    if hess_solver is None:
        if hessian is None:
            hessian = _n(tf.hessian(objective, argnums=0)(_t(theta0), _t(lam0)))
        if hessian.shape != (len(theta0), len(theta0)):
            raise ValueError('``hessian_at_opt`` is the wrong shape.')
        hess_solver = solver_lib.get_cholesky_solver(hessian)
    if validate_optimum:
        g = _n(grad_theta(_t(theta0), _t(lam0)))
        gnorm = np.linalg.norm(g)
        if gnorm > grad_tol:
            raise ValueError(
                'The estimating equation is not zero at the proposed  values.  '
                '||ee|| = {} > {} = solution_tol'.format(gnorm, grad_tol))
    if cross_hessian is None:
        hyper_grad = tf.grad(objective if hyper_objective is None else hyper_objective, argnums=0)
        cross_hessian = _n(tf.jacrev(hyper_grad, argnums=1)(_t(theta0), _t(lam0)))
    if cross_hessian.shape != (len(theta0), len(lam0)):
        raise ValueError('``_estimating_equation_jac0`` is the wrong shape.')
    sens = -1 * hess_solver(cross_hessian)
    return dict(hessian=hessian, cross_hessian=cross_hessian, sens=sens, solve=hess_solver)


def predict_from_hyper(theta0, lam0, sens, lam1):
    """``predict_input_par_from_hyper_par`` (``sensitivity_lib.py:236-247``)."""
    return np.asarray(theta0) + sens @ (np.asarray(lam1) - np.asarray(lam0))


# --------------------------------------------------------------------------
# Forward-mode directional derivatives  (sensitivity_lib.py:440-492, 766-807)
# --------------------------------------------------------------------------

def directional_derivative(g, eta0, eps0, eta_dirs, eps_dirs):
    """``ForwardModeDerivativeArray.eval_directional_derivative``
    (``sensitivity_lib.py:788-807``): g differentiated ``len(eta_dirs)`` times
    in argument 0, then ``len(eps_dirs)`` times in argument 1, each derivative
    contracted with its direction - built, like ``_append_jvp``
    (``:440-492``), from nested forward-mode JVPs.  ``g`` is a torch callable
    ``(eta, eps) -> (D,)``."""
    fun = g
    dirs = [(0, _t(v)) for v in eta_dirs] + [(1, _t(v)) for v in eps_dirs]

    def wrap(f, argnum, v):
        def f2(eta, eps):
            args = [eta, eps]

            def f_of(x):
                a = list(args)
                a[argnum] = x
                return f(*a)
            return tf.jvp(f_of, (args[argnum],), (v,))[1]
        return f2
    for argnum, v in dirs:
        fun = wrap(fun, argnum, v)
    return _n(fun(_t(eta0), _t(eps0)))


# --------------------------------------------------------------------------
# Taylor-term bookkeeping  (sensitivity_lib.py:495-688, 980-1018)
# --------------------------------------------------------------------------

def taylor_term_table(order):
    """Terms of d^k/d eps^k g(eta(eps), eps) for k = 1..order.

    Each term is ``(prefactor, eps_order, eta_orders)`` with ``eta_orders[i]``
    the number of factors d^{i+1} eta / d eps^{i+1}
    (``DerivativeTerm``, ``sensitivity_lib.py:495-688``).  Order 1 is
    ``dg/deps + dg/deta * eta'`` (``_get_taylor_base_terms`` ``:1008-1018``);
    order k+1 applies the product/chain rule to every term
    (``differentiate`` ``:638-673``) and merges like terms
    (``_consolidate_terms`` ``:980-1004``).  Terms are merged fully here, keyed
    by their derivative signature, so the SUM of the terms is what is pinned
    against the reference, not the list order."""
    tables = [{(1, (0,)): 1.0, (0, (1,)): 1.0}]
    for _ in range(1, order):
        nxt = {}

        def add(eps_order, eta_orders, pref):
            key = (eps_order, tuple(eta_orders))
            nxt[key] = nxt.get(key, 0.0) + pref
        for (eps_order, eta_orders), pref in tables[-1].items():
            base = list(eta_orders) + [0]
            add(eps_order + 1, base, pref)                 # d/d eps of the g partial
            e = list(base); e[0] += 1
            add(eps_order, e, pref)                        # d/d eta of the g partial times eta'
            for i, cnt in enumerate(eta_orders):           # derivative of each eta^(i+1) factor
                if cnt > 0:
                    e = list(base); e[i] -= 1; e[i + 1] += 1
                    add(eps_order, e, pref * cnt)
        tables.append(nxt)
    return [[(p, k[0], list(k[1])) for k, p in t.items()] for t in tables]


def taylor_input_derivs(g, eta0, eps0, deps, order, solve,
                        max_input_order=None, max_hyper_order=None):
    """``evaluate_input_derivs`` (``sensitivity_lib.py:1274-1286``) over
    ``_evaluate_dkinput_dhyperk`` (``:1208-1260``): for k = 1..order,
    ``d^k eta = -solve(sum_terms prefactor * d^m g[eta dirs..., deps...])``
    skipping the term holding the unknown (``:1238``) and truncated orders
    (``:1242-1249``); each term is expanded as in ``_evaluate_term_fwd``
    (``:691-734``)."""
    tables = taylor_term_table(order)
    eta0 = np.asarray(eta0, dtype=np.float64)
    derivs = []
    for k in range(1, order + 1):
        vec = np.zeros_like(eta0)
        for pref, eps_order, eta_orders in tables[k - 1]:
            if eta_orders[-1] > 0:
                continue
            if max_hyper_order is not None and eps_order > max_hyper_order:
                continue
            if max_input_order is not None and sum(eta_orders) > max_input_order:
                continue
            eta_dirs = []
            for i, cnt in enumerate(eta_orders):
                if cnt > 0:
                    eta_dirs += [derivs[i]] * cnt
            eps_dirs = [deps] * eps_order
            vec = vec + pref * directional_derivative(g, eta0, eps0, eta_dirs, eps_dirs)
        derivs.append(-1 * solve(vec))
    return derivs


def taylor_series_terms(g, eta0, eps0, eps1, order, solve, add_offset=True, **kw):
    """``evaluate_taylor_series_terms`` (``sensitivity_lib.py:1289-1304``)."""
    eta0 = np.asarray(eta0, dtype=np.float64)
    deps = np.asarray(eps1, dtype=np.float64) - np.asarray(eps0, dtype=np.float64)
    derivs = taylor_input_derivs(g, eta0, eps0, deps, order, solve, **kw)
    terms = [eta0 if add_offset else np.zeros_like(eta0)]
    for k in range(1, order + 1):
        terms.append(derivs[k - 1] / float(factorial(k)))
    return terms


def taylor_series(g, eta0, eps0, eps1, order, solve, add_offset=True, **kw):
    """``evaluate_taylor_series`` (``sensitivity_lib.py:1307-1343``)."""
    return np.sum(taylor_series_terms(g, eta0, eps0, eps1, order, solve, add_offset, **kw), axis=0)
