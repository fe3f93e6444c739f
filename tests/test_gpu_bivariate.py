"""GPU parity: CrossSensitivity and OptimumChecker (SURVEY.md section 8f item 1;
reference ``bivariate_sensitivity_lib.py``) against the golden values produced
by the unmodified reference (``oracle/make_golden.py::golden_bivariate``)."""
import warnings

import numpy as np
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def vt():
    import vittles_b200
    return vittles_b200


def _dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device='cuda')


def _model(g):
    xt, yt = _dev(g['x']), _dev(g['y'])

    def w_obj(theta, w):
        resid = yt - torch.exp(xt @ theta)
        return 0.5 * torch.sum(w * resid ** 2)

    def pert_obj(theta, lam, w):
        return w_obj(theta, w) - torch.dot(lam, theta)
    return w_obj, pert_obj


def test_cross_sensitivity_vs_golden(vt, golden):
    from vittles_b200.bivariate_sensitivity_lib import CrossSensitivity
    g = golden('bivariate')
    _, pert_obj = _model(g)
    solver = vt.solver_lib.get_cholesky_solver(g['hess_base'])            # GPU Cholesky closure
    w_base = np.ones(len(g['y']))
    with pytest.warns(UserWarning, match='experimental'):
        cs = CrossSensitivity(estimating_equation=torch.func.grad(pert_obj, argnums=0), solver=solver,
                              input_base=g['theta_base'], hyper1_base=g['lam_base'], hyper2_base=w_base)
    dlambda, dw = -1 * g['lam_base'], g['new_w'] - w_base
    di1 = cs.get_di1(dlambda)
    assert isinstance(di1, np.ndarray)                                   # numpy in -> numpy out
    assert_close(di1, g['di1'])
    assert_close(cs.get_di2(dw), g['di2'])
    assert_close(cs.evaluate(dlambda, dw), g['cross'])
    # supplying the first-order directions gives the same answer
    assert_close(cs.evaluate(dlambda, dw, di1=g['di1'], di2=g['di2']), g['cross'])
    # a numpy-only solver closure is accepted as well
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        cs_np = CrossSensitivity(estimating_equation=torch.func.grad(pert_obj, argnums=0),
                                 solver=lambda v: np.linalg.solve(g['hess_base'], v),
                                 input_base=g['theta_base'], hyper1_base=g['lam_base'], hyper2_base=w_base)
    assert_close(cs_np.evaluate(dlambda, dw), g['cross'])


def test_cross_sensitivity_term_switches(vt, golden):
    """The reference raises AttributeError (`_term_i12`, :73,77) when term_ii is off; here the
    switches work and the four terms add up to the full answer."""
    from vittles_b200.bivariate_sensitivity_lib import CrossSensitivity
    g = golden('bivariate')
    _, pert_obj = _model(g)
    solver = vt.solver_lib.get_cholesky_solver(g['hess_base'])
    w_base = np.ones(len(g['y']))
    dlambda, dw = -1 * g['lam_base'], g['new_w'] - w_base
    total = 0.0
    for k in range(4):
        flags = dict(term_ii=False, term_i1=False, term_i2=False, term_12=False)
        flags[list(flags)[k]] = True
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            cs = CrossSensitivity(torch.func.grad(pert_obj, argnums=0), solver, g['theta_base'], g['lam_base'],
                                  w_base, **flags)
        total = total + cs.evaluate(dlambda, dw)
    assert_close(total, g['cross'], rtol=1e-8, atol_scale=1e-11)


def test_optimum_checker_vs_golden(vt, golden):
    from vittles_b200.bivariate_sensitivity_lib import OptimumChecker
    g = golden('bivariate')
    w_obj, _ = _model(g)
    solver = vt.solver_lib.get_cholesky_solver(g['hess_base'])
    w_base = np.ones(len(g['y']))
    oc = OptimumChecker(estimating_equation=torch.func.grad(w_obj, argnums=0), solver=solver,
                        input_base=g['theta_base'], hyper_base=w_base)
    # the reference test's assertions (tests/test_bivariate_sensitivity_lib.py:249-257)
    assert_close(oc.get_newton_step(), -np.linalg.solve(g['hess_base'], g['lam_base']))
    assert_close(oc.get_newton_step(), g['oc_newton_step'])
    assert_close(oc.get_dinput_dhyper(g['new_w'] - w_base), g['oc_dinput_dhyper'])
    assert_close(oc.correction(g['new_w']), g['oc_correction'], atol_scale=1e-11)
    assert_close(oc.evaluate(g['new_w']), g['oc_evaluate'])


def test_optimum_checker_structured_glm(vt, golden):
    """Poisson GLM through the fused kernels (vt_glm_stats / vt_glm_dirderiv) vs the reference."""
    from vittles_b200.bivariate_sensitivity_lib import OptimumChecker
    g = golden('bivariate')
    obj = vt.objectives.GLMObjective(g['p_X'], g['p_y'], family='poisson', l2=float(g['p_l2']))
    hess = obj.vt_hessian(_dev(g['p_theta']), _dev(g['p_w']))
    assert_close(hess, g['p_hess'], rtol=1e-10)
    solver = vt.solver_lib.get_cholesky_solver(hess)
    oc = OptimumChecker(estimating_equation=obj, solver=solver, input_base=g['p_theta'], hyper_base=g['p_w'])
    assert_close(oc.get_newton_step(), g['p_oc_newton_step'])
    assert_close(oc.get_dinput_dhyper(g['p_new_w'] - g['p_w']), g['p_oc_dinput_dhyper'])
    assert_close(oc.correction(g['p_new_w']), g['p_oc_correction'], atol_scale=1e-11)
    assert_close(oc.evaluate(g['p_new_w']), g['p_oc_evaluate'])
    # device tensors in -> device tensors out
    oc_d = OptimumChecker(estimating_equation=obj, solver=solver, input_base=_dev(g['p_theta']),
                          hyper_base=_dev(g['p_w']))
    out = oc_d.evaluate(_dev(g['p_new_w']))
    assert isinstance(out, torch.Tensor) and out.is_cuda
    assert_close(out, g['p_oc_evaluate'])
