"""Oracle restatement of ``vittles/bivariate_sensitivity_lib.py``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Estimating equations
are torch callables evaluated in float64 on the CPU; the nested forward-mode
directions the reference builds with ``_append_jvp`` (``sensitivity_lib.py:440-492``)
are nested ``torch.func.jvp`` calls here.  Inputs and outputs are numpy arrays.
"""
import numpy as np
import torch
from torch import func as tf


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


def _n(a):
    return a.detach().numpy().copy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)


def _directional(g, base, argnums, dirs):
    """d^k g / d x_{argnums[0]} ... d x_{argnums[k-1]} contracted with ``dirs``,
    at ``base`` (a list of tensors): the ``_append_jvp`` chain of
    ``bivariate_sensitivity_lib.py:47-53``."""
    def nest(f, argnum, v):
        def f2(*args):
            def of_arg(x):
                a = list(args)
                a[argnum] = x
                return f(*a)
            return tf.jvp(of_arg, (args[argnum],), (v,))[1]
        return f2
    f = g
    for argnum, v in zip(argnums, dirs):
        f = nest(f, argnum, _t(v))
    return _n(f(*[_t(b) for b in base]))


def cross_sensitivity(g, solver, input_base, hyper1_base, hyper2_base, dh1, dh2, di1=None, di2=None,
                      term_ii=True, term_i1=True, term_i2=True, term_12=True):
    """``CrossSensitivity.evaluate`` (``bivariate_sensitivity_lib.py:71-115``):

        d^2 theta / d eps1 d eps2 [dh1, dh2]
            = -H^{-1} ( g_ii[di1, di2] + g_i1[di2, dh1] + g_i2[di1, dh2] + g_12[dh1, dh2] )

    with ``di1 = -H^{-1} g_1 dh1`` (``:57-62``) and ``di2 = -H^{-1} g_2 dh2``
    (``:64-69``).  Returns (result, di1, di2)."""
    base = [input_base, hyper1_base, hyper2_base]
    if di1 is None:
        di1 = -1 * np.asarray(solver(_directional(g, base, [1], [dh1])))
    if di2 is None:
        di2 = -1 * np.asarray(solver(_directional(g, base, [2], [dh2])))
    total = 0
    if term_ii:
        total = total + _directional(g, base, [0, 0], [di1, di2])
    if term_i1:
        total = total + _directional(g, base, [0, 1], [di2, dh1])
    if term_i2:
        total = total + _directional(g, base, [0, 2], [di1, dh2])
    if term_12:
        total = total + _directional(g, base, [1, 2], [dh1, dh2])
    return -1 * np.asarray(solver(total)), di1, di2


def optimum_checker(g2, solver, input_base, hyper_base, hyper_new):
    """``OptimumChecker`` (``bivariate_sensitivity_lib.py:118-202``) for an estimating
    equation ``g2(input, hyper)``: the Lagrange-shifted equation ``g2 + lam``
    (``:146-148``) at ``lam_base = -g2(input_base, hyper_base)`` (``:154-156``),
    cross sensitivity with ``term_i2 = term_12 = False`` (``:158-165``).
    Returns the Newton step (``:167-170``), the first-order change of the
    optimum (``:172-176``), the correction (``:178-193``) and ``evaluate``
    (``:195-205``)."""
    lam_base = -1 * _n(g2(_t(input_base), _t(hyper_base)))
    dlam = -1 * lam_base

    def g3(ipar, hpar, lam):
        return g2(ipar, hpar) + lam
    base = [input_base, hyper_base, lam_base]
    newton_step = -1 * np.asarray(solver(_directional(g3, base, [2], [dlam])))
    dhyper = np.asarray(hyper_new) - np.asarray(hyper_base)
    dinput = -1 * np.asarray(solver(_directional(g3, base, [1], [dhyper])))
    correction, _, _ = cross_sensitivity(g3, solver, input_base, hyper_base, lam_base, dhyper, dlam,
                                         di1=dinput, di2=newton_step, term_i2=False, term_12=False)
    return dict(newton_step=newton_step, dinput_dhyper=dinput, correction=correction,
                evaluate=np.asarray(input_base) + dinput + correction)
