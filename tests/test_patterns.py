"""CPU tests of ``vittles_b200.patterns`` (paragami-style flatten / fold on torch).
Pure torch - no kernel is involved - so they run without a GPU.  The GPU half (the reference's QuadraticModel
set-up written with these patterns, through the sensitivity classes, against the golden values of the reference) is
``tests/test_gpu_ij.py::test_quadratic_model_through_patterns_vs_golden``."""
import numpy as np
import pytest
import torch
from torch import func as tf

from vittles_b200 import patterns as pg


@pytest.mark.parametrize('lb,ub', [(-np.inf, np.inf), (-20.0, np.inf), (-np.inf, 3.0), (-1.0, 2.5)])
@pytest.mark.parametrize('free', [False, True])
def test_numeric_array_pattern_round_trip(lb, ub, free):
    pat = pg.NumericArrayPattern(shape=(2, 3), lb=lb, ub=ub)
    rng = np.random.RandomState(0)
    lo = lb if np.isfinite(lb) else -5.0
    hi = ub if np.isfinite(ub) else 5.0
    x = lo + (hi - lo) * rng.uniform(0.05, 0.95, size=(2, 3))
    flat = pat.flatten(x, free=free)
    assert isinstance(flat, np.ndarray) and flat.shape == (pat.flat_length(free),) == (6,)
    np.testing.assert_allclose(pat.fold(flat, free=free), x, rtol=1e-13, atol=1e-13)
    # tensors in, tensors out, differentiable
    ft = torch.as_tensor(flat)
    assert isinstance(pat.fold(ft, free=free), torch.Tensor)
    jac = tf.jacrev(lambda v: pat.fold(v, free=free).reshape(-1))(ft)
    if not free or not (np.isfinite(lb) or np.isfinite(ub)):
        np.testing.assert_allclose(jac.numpy(), np.eye(6), atol=1e-14)
    elif np.isfinite(lb) and not np.isfinite(ub):
        np.testing.assert_allclose(np.diag(jac.numpy()), (x - lb).reshape(-1), rtol=1e-12)   # d(exp(f)+lb)/df = x - lb
    if np.isfinite(lb):
        with pytest.raises(ValueError):
            pat.flatten(np.full((2, 3), lb - 1.0), free=free)
    with pytest.raises(ValueError):
        pat.fold(np.zeros(5), free=free)


def test_psd_pattern_matches_the_fixture_map_and_round_trips():
    from oracle.fixtures import _psd_from_free
    pat = pg.PSDSymmetricMatrixPattern(size=3)
    assert pat.flat_length(True) == 6 and pat.flat_length(False) == 9
    v = torch.as_tensor(np.random.RandomState(1).normal(size=6))
    A = pat.fold(v, free=True)
    torch.testing.assert_close(A, _psd_from_free(v), rtol=1e-14, atol=1e-14)
    assert torch.all(torch.linalg.eigvalsh(A) > 0)
    torch.testing.assert_close(pat.flatten(A, free=True), v, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(pat.fold(pat.flatten(A, free=False), free=False), A)
    # second derivatives through the fold (what SparseBlockHessian needs)
    h = tf.hessian(lambda f: torch.trace(pat.fold(f, free=True)))(v)
    assert h.shape == (6, 6) and torch.isfinite(h).all()


def test_pattern_dict_and_array():
    d = pg.PatternDict()
    d['mean'] = pg.NumericVectorPattern(4)
    d['var'] = pg.NumericArrayPattern(shape=(4,), lb=0.0)
    d['cov'] = pg.PatternArray((2,), pg.PSDSymmetricMatrixPattern(3))
    assert d.flat_length(False) == 4 + 4 + 18 and d.flat_length(True) == 4 + 4 + 12
    rng = np.random.RandomState(2)
    a = rng.normal(size=(2, 3, 3))
    val = {'mean': rng.normal(size=4), 'var': rng.uniform(0.5, 2.0, size=4),
           'cov': a @ a.transpose(0, 2, 1) + np.eye(3)}
    for free in (False, True):
        flat = d.flatten(val, free=free)
        assert flat.shape == (d.flat_length(free),)
        back = d.fold(flat, free=free)
        assert list(back.keys()) == ['mean', 'var', 'cov']
        for k in val:
            np.testing.assert_allclose(back[k], val[k], rtol=1e-12, atol=1e-12)
    mask = d.empty_bool(False) if hasattr(d, 'empty_bool') else None
    assert mask.shape == (26,)
    sel = {'mean': np.array([True, False, False, True]), 'var': np.zeros(4, bool), 'cov': np.zeros((2, 3, 3), bool)}
    np.testing.assert_array_equal(d.flat_indices(sel, free=False), [0, 3])
    d.lock()
    with pytest.raises(ValueError):
        d['other'] = pg.NumericVectorPattern(1)


@pytest.mark.parametrize('theta_free,lambda_free', [(False, False), (True, False), (False, True), (True, True)])
def test_flatten_function_input_reproduces_the_quadratic_fixture(theta_free, lambda_free):
    """The reference's QuadraticModel (``test_utils.py:23-75``) written with these patterns equals the oracle's
    hand-flattened restatement of it (``oracle/fixtures.py``) - values and derivatives."""
    from oracle.fixtures import QuadraticModel
    dim = 3
    ref = QuadraticModel(dim)
    theta_pattern = pg.NumericArrayPattern(shape=(dim,), lb=-20.0)
    lambda_pattern = pg.NumericArrayPattern(shape=(dim,), lb=-20.0)
    A = torch.as_tensor(ref.matrix)

    def objective(theta, lam):
        return 0.5 * theta @ A @ theta + lam @ theta

    def true_optimal_theta(lam):
        return -1 * torch.linalg.solve(A, lam)
    flat_obj = pg.FlattenFunctionInput(objective, free=[theta_free, lambda_free], argnums=[0, 1],
                                       patterns=[theta_pattern, lambda_pattern])
    flat_opt = pg.FlattenFunctionInputAndOutput(true_optimal_theta, input_free=lambda_free, output_free=theta_free,
                                                input_patterns=lambda_pattern, output_patterns=theta_pattern,
                                                input_argnums=[0], output_retnums=[0])
    lam0 = ref.get_default_lambda()
    theta0 = ref.get_true_optimal_theta(lam0)
    t_flat, l_flat = theta_pattern.flatten(theta0, theta_free), lambda_pattern.flatten(lam0, lambda_free)
    rt, rl = ref.get_default_flat_values(theta_free, lambda_free)
    np.testing.assert_allclose(t_flat, rt, rtol=1e-14)
    np.testing.assert_allclose(l_flat, rl, rtol=1e-14)
    tt, lt = torch.as_tensor(t_flat), torch.as_tensor(l_flat)
    ref_obj = ref.get_flat_objective(theta_free, lambda_free)
    torch.testing.assert_close(flat_obj(tt, lt), ref_obj(tt, lt), rtol=1e-14, atol=1e-14)
    torch.testing.assert_close(tf.hessian(flat_obj)(tt, lt), tf.hessian(ref_obj)(tt, lt), rtol=1e-12, atol=1e-12)
    assert float(torch.linalg.vector_norm(tf.grad(flat_obj)(tt, lt))) < 1e-10          # theta0 is the optimum
    ref_opt = ref.get_flat_true_optimal_theta(theta_free, lambda_free)
    torch.testing.assert_close(flat_opt(lt), ref_opt(lt), rtol=1e-13, atol=1e-13)
    torch.testing.assert_close(tf.jacrev(flat_opt)(lt), tf.jacrev(ref_opt)(lt), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(flat_opt(l_flat), ref_opt(lt).numpy(), rtol=1e-13)       # numpy in, numpy out
