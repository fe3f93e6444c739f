"""Second-order cross sensitivities and the incomplete-optimisation check - the
drop-in for ``vittles/bivariate_sensitivity_lib.py``.

Same class names, constructor keywords and methods as the reference.  The
solve seam is unchanged: ``solver`` is any closure ``v -> H^{-1} v`` (the GPU
Cholesky / CG closures of :mod:`vittles_b200.solver_lib`, or a user's own).

Two evaluation paths, chosen by the TYPE of the estimating equation:

* a torch callable ``g(input, hyper1, hyper2) -> vector``: the directional
  derivatives are nested ``torch.func.jvp`` on the GPU, built with the same
  ``_append_jvp`` chain as the reference (``bivariate_sensitivity_lib.py:47-53``);
* :class:`OptimumChecker` on a :class:`~vittles_b200.objectives.StructuredObjective`
  (meaning "the gradient of this objective"): every term is one fused pass over
  the design matrix (``vt_glm_stats`` / ``vt_glm_dirderiv``), the same kernels as
  the Taylor expansion's order-2 terms.

One deliberate difference: the reference's ``evaluate`` reads a misspelt
attribute ``_term_i12`` (``:73,77``) and raises ``AttributeError`` whenever
``term_ii`` is switched off; here the test is on the terms that actually need
each direction.
"""
import warnings
from copy import deepcopy

import torch

from .sensitivity_lib import _append_jvp
from ._arrays import to_device, kind_of, as_kind
from .objectives import StructuredObjective


def _solve(solver, v, dev):
    """Apply a solver closure to a device vector; accept numpy-only closures."""
    try:
        out = solver(v)
    except TypeError:
        out = solver(v.detach().cpu().numpy())
    return to_device(out, dev)


class CrossSensitivity():
    """Directional second derivative of an optimum in two hyperparameters,

        d^2 theta_hat / d eps1 d eps2 [dh1, dh2]
            = -H^{-1} (g_ii[di1, di2] + g_i1[di2, dh1] + g_i2[di1, dh2] + g_12[dh1, dh2]),

    for an estimating equation ``g(theta, eps1, eps2)`` (reference:
    ``bivariate_sensitivity_lib.py:8-115``).  Flagged experimental upstream; the
    same ``UserWarning`` is raised here."""

    def __init__(self, estimating_equation, solver, input_base, hyper1_base, hyper2_base,
                 term_ii=True, term_i1=True, term_i2=True, term_12=True):
        warnings.warn('The CrossSensitivity class is very experimental and untested.')
        self._g = estimating_equation
        self._solver = solver
        self._kind = kind_of(input_base)
        # copies: the solver is only valid at these values (reference :40-43)
        self._input_base = to_device(deepcopy(input_base))
        dev = self._input_base.device
        self._hyper1_base = to_device(deepcopy(hyper1_base), dev)
        self._hyper2_base = to_device(deepcopy(hyper2_base), dev)

        self._g_i = _append_jvp(self._g, num_base_args=3, argnum=0)
        self._g_ii = _append_jvp(self._g_i, num_base_args=3, argnum=0)
        self._g_i1 = _append_jvp(self._g_i, num_base_args=3, argnum=1)
        self._g_i2 = _append_jvp(self._g_i, num_base_args=3, argnum=2)
        self._g_1 = _append_jvp(self._g, num_base_args=3, argnum=1)
        self._g_2 = _append_jvp(self._g, num_base_args=3, argnum=2)
        self._g_12 = _append_jvp(self._g_1, num_base_args=3, argnum=2)

        self._term_ii = term_ii
        self._term_i1 = term_i1
        self._term_i2 = term_i2
        self._term_12 = term_12

    def _base(self):
        return self._input_base, self._hyper1_base, self._hyper2_base

    def _dev(self, v):
        return to_device(v, self._input_base.device)

    def _di1(self, dh1):
        return -1 * _solve(self._solver, self._g_1(*self._base(), self._dev(dh1)), self._input_base.device)

    def _di2(self, dh2):
        return -1 * _solve(self._solver, self._g_2(*self._base(), self._dev(dh2)), self._input_base.device)

    def get_di1(self, dh1):
        """``-H^{-1} dg/d eps1 . dh1`` (reference ``:57-62``)."""
        return as_kind(self._di1(dh1), self._kind)

    def get_di2(self, dh2):
        """``-H^{-1} dg/d eps2 . dh2`` (reference ``:64-69``)."""
        return as_kind(self._di2(dh2), self._kind)

    def _evaluate(self, dh1, dh2, di1=None, di2=None, debug=False):
        dh1, dh2 = self._dev(dh1), self._dev(dh2)
        if (self._term_ii or self._term_i2) and di1 is None:
            di1 = self._di1(dh1)
        if (self._term_ii or self._term_i1) and di2 is None:
            di2 = self._di2(dh2)
        di1 = None if di1 is None else self._dev(di1)
        di2 = None if di2 is None else self._dev(di2)
        base = self._base()
        g_ii = g_i1 = g_i2 = g_12 = 0
        if self._term_ii:
            g_ii = self._g_ii(*base, di1, di2)
        if self._term_i1:
            g_i1 = self._g_i1(*base, di2, dh1)
        if self._term_i2:
            g_i2 = self._g_i2(*base, di1, dh2)
        if self._term_12:
            g_12 = self._g_12(*base, dh1, dh2)
        if debug:
            print('g_ii: ', g_ii)
            print('g_i1: ', g_i1)
            print('g_i2: ', g_i2)
            print('g_12: ', g_12)
            print('di1: ', di1)
            print('di2: ', di2)
        total = g_ii + g_i1 + g_i2 + g_12
        if not isinstance(total, torch.Tensor):
            total = torch.zeros_like(self._input_base)
        return -1 * _solve(self._solver, total, self._input_base.device)

    def evaluate(self, dh1, dh2, di1=None, di2=None, debug=False):
        """Reference ``:71-115``."""
        return as_kind(self._evaluate(dh1, dh2, di1=di1, di2=di2, debug=debug), self._kind)


class _StructuredLagrangeTerms:
    """The derivative closures of ``g(theta, w) + lam`` for a structured
    objective, each one fused pass over the data: what ``CrossSensitivity``
    builds with nested JVPs for generic callables."""

    def __init__(self, objective):
        self._obj = objective
        self._cache = {}

    def _dd(self, theta, w, eta_dirs, eps_dirs):
        return self._obj.vt_directional_derivative(theta, w, eta_dirs, eps_dirs, cache=self._cache)

    def g(self, theta, w):
        return self._dd(theta, w, [], [])

    def g_1(self, theta, w, lam, dh1):
        return self._dd(theta, w, [], [dh1])

    def g_2(self, theta, w, lam, dlam):
        return dlam

    def g_ii(self, theta, w, lam, di1, di2):
        return self._dd(theta, w, [di1, di2], [])

    def g_i1(self, theta, w, lam, di2, dh1):
        return self._dd(theta, w, [di2], [dh1])


class OptimumChecker():
    """Estimate the error in a sensitivity due to incomplete optimisation
    (reference: ``bivariate_sensitivity_lib.py:118-205``).

    ``estimating_equation`` is a torch callable ``(input, hyper) -> vector`` or a
    :class:`~vittles_b200.objectives.StructuredObjective` (standing for its
    gradient in the input).  ``solver`` solves with the Hessian at
    ``input_base``, ``hyper_base``."""

    def __init__(self, estimating_equation, solver, input_base, hyper_base):
        self._kind = kind_of(input_base)
        self._input_base = to_device(deepcopy(input_base))
        dev = self._input_base.device
        self._hyper_base = to_device(deepcopy(hyper_base), dev)
        self._solver = solver
        self._structured = isinstance(estimating_equation, StructuredObjective)

        if self._structured:
            terms = _StructuredLagrangeTerms(estimating_equation)
            g0 = terms.g(self._input_base, self._hyper_base)

            def estimating_equation_lagrange(ipar, hpar, lam):
                return terms.g(ipar, hpar) + lam
        else:
            g0 = estimating_equation(self._input_base, self._hyper_base)

            def estimating_equation_lagrange(ipar, hpar, lam):
                return estimating_equation(ipar, hpar) + lam
        self.estimating_equation_lagrange = estimating_equation_lagrange

        self._lam_base = -1 * g0
        self._dlam = -1 * self._lam_base

        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            self._cross_sens = CrossSensitivity(
                estimating_equation=self.estimating_equation_lagrange,
                solver=self._solver,
                input_base=self._input_base,
                hyper1_base=self._hyper_base,
                hyper2_base=self._lam_base,
                term_i2=False,
                term_12=False)
        if self._structured:
            # same contract as the nested-JVP closures, evaluated by the fused kernels
            cs = self._cross_sens
            cs._g_1, cs._g_2, cs._g_ii, cs._g_i1 = terms.g_1, terms.g_2, terms.g_ii, terms.g_i1

    def get_newton_step(self):
        """A Newton step towards the optimum (reference ``:167-170``)."""
        return as_kind(self._cross_sens._di2(self._dlam), self._kind)

    def get_dinput_dhyper(self, dhyper):
        """First directional derivative of the optimum in the direction
        ``dhyper`` (reference ``:172-176``)."""
        return as_kind(self._cross_sens._di1(dhyper), self._kind)

    def _correction(self, hyper_new, dinput_dhyper=None, newton_step=None):
        dev = self._input_base.device
        dhyper = to_device(hyper_new, dev) - self._hyper_base
        if dinput_dhyper is None:
            dinput_dhyper = self._cross_sens._di1(dhyper)
        if newton_step is None:
            newton_step = self._cross_sens._di2(self._dlam)
        corr = self._cross_sens._evaluate(dhyper, self._dlam, di1=dinput_dhyper, di2=newton_step)
        return to_device(dinput_dhyper, dev), corr

    def correction(self, hyper_new, dinput_dhyper=None, newton_step=None):
        """First-order correction to ``dinput_dhyper`` from taking the Newton
        step (reference ``:178-193``)."""
        return as_kind(self._correction(hyper_new, dinput_dhyper, newton_step)[1], self._kind)

    def evaluate(self, hyper_new, dinput_dhyper=None, newton_step=None):
        """``input_base + dinput_dhyper + correction`` (reference ``:195-205``)."""
        dinput, corr = self._correction(hyper_new, dinput_dhyper, newton_step)
        return as_kind(self._input_base + dinput + corr, self._kind)
