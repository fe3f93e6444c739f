"""GPU parity: LinearResponseCovariances."""
import numpy as np
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def vt():
    import vittles_b200
    return vittles_b200


def test_lr_cov_vs_golden(vt, golden):
    """tests/test_lr_cov_lib.py:16-135: the LR covariance of the mean of an
    MVN target is exactly the target covariance; shape errors; optimum check."""
    from oracle import models
    g = golden('lr_cov')
    f = models.mvn_kl_objective(g['true_mean'], g['true_info'])

    def f_dev(par):
        tm = torch.as_tensor(g['true_mean'], device=par.device)
        ti = torch.as_tensor(g['true_info'], device=par.device)
        mean, var = par[:4], par[4:]
        tc = mean - tm
        return -1 * (0.5 * torch.sum(torch.log(var)) - 0.5 * (torch.sum(torch.diagonal(ti) * var) + tc @ ti @ tc))
    for hess in (None, g['hessian']):
        lr = vt.LinearResponseCovariances(f_dev, g['opt'], validate_optimum=True, hessian_at_opt=hess, grad_tol=1e-12)
        assert_close(lr.get_hessian_at_opt(), g['hessian'], rtol=1e-9, atol_scale=1e-12)
        cov = lr.get_lr_covariance(lambda par: par[:4])
        assert_close(cov, g['cov'], rtol=1e-8)
        assert_close(cov, g['true_cov'], rtol=1e-8)
        jac = lr.get_moment_jacobian(lambda par: par[:4])
        assert_close(jac, g['jac'])
        assert_close(lr.get_lr_covariance_from_jacobians(jac[0:2], jac[2:4]), g['cross01_23'], rtol=1e-8,
                     atol_scale=1e-10)
        for a, b in [(jac.T, jac), (jac, jac.T), (jac[:, :, None], jac), (jac, jac[:, :, None])]:
            with pytest.raises(ValueError):
                lr.get_lr_covariance_from_jacobians(a, b)
    with pytest.raises(ValueError):
        vt.LinearResponseCovariances(f_dev, g['opt'] + 0.01, validate_optimum=True, grad_tol=1e-12)


@pytest.mark.parametrize('dim,k', [(96, 17), (700, 300)])
def test_lr_cov_dense_closed_form(vt, dim, k):
    """Config-4 family at sizes that exercise the blocked Cholesky: mean-field
    normal VB of an MVN target with 2*dim parameters (closed-form Hessian),
    LR covariance of k random linear moments vs numpy."""
    from oracle import models
    rng = np.random.RandomState(dim)
    a = rng.normal(size=(dim, dim + 3))
    true_cov = a @ a.T / dim + np.eye(dim)
    true_info = np.linalg.inv(true_cov)
    true_mean = rng.normal(size=dim)
    opt, H = models.mvn_kl_closed_form(true_mean, true_info)
    lr = vt.LinearResponseCovariances(lambda par: par.sum(), opt, hessian_at_opt=H)
    J = rng.normal(size=(k, 2 * dim))
    ref = J @ np.linalg.solve(H, J.T)
    assert_close(lr.get_lr_covariance_from_jacobians(J, J), ref, rtol=1e-8, atol_scale=1e-11)
    # mean moments recover the true covariance exactly
    Jm = np.hstack([np.eye(dim), np.zeros((dim, dim))])
    assert_close(lr.get_lr_covariance_from_jacobians(Jm, Jm), true_cov, rtol=1e-8, atol_scale=1e-11)


def test_lr_cov_without_factorisation(vt, golden):
    """factorize_hessian=False (documented upstream at lr_cov_lib.py:67-70, never implemented
    there): CG over Hessian-vector products, no Hessian formed; same covariances."""
    g = golden('lr_cov')

    def f_dev(par):
        tm = torch.as_tensor(g['true_mean'], device=par.device)
        ti = torch.as_tensor(g['true_info'], device=par.device)
        mean, var = par[:4], par[4:]
        tc = mean - tm
        return -1 * (0.5 * torch.sum(torch.log(var)) - 0.5 * (torch.sum(torch.diagonal(ti) * var) + tc @ ti @ tc))
    for hess in (None, g['hessian']):
        lr = vt.LinearResponseCovariances(f_dev, g['opt'], validate_optimum=True, hessian_at_opt=hess,
                                          factorize_hessian=False, grad_tol=1e-12)
        assert lr._hess0 is hess                      # nothing formed behind the caller's back
        assert_close(lr.get_lr_covariance(lambda par: par[:4]), g['true_cov'], rtol=1e-8)
        assert_close(lr.get_hessian_at_opt(), g['hessian'], rtol=1e-9, atol_scale=1e-12)
    with pytest.raises(ValueError):
        vt.LinearResponseCovariances(f_dev, g['opt'] + 0.01, validate_optimum=True, factorize_hessian=False,
                                     grad_tol=1e-12)
    # larger, supplied Hessian: the operator is the GEMV kernel
    rng = np.random.RandomState(5)
    dim = 300
    a = rng.normal(size=(dim, dim + 3))
    H = a @ a.T / dim + np.eye(dim)
    J = rng.normal(size=(7, dim))
    lr = vt.LinearResponseCovariances(lambda par: par.sum(), np.zeros(dim), hessian_at_opt=H, factorize_hessian=False)
    assert_close(lr.get_lr_covariance_from_jacobians(J, J), J @ np.linalg.solve(H, J.T), rtol=1e-8, atol_scale=1e-10)


def test_lr_cov_config4_full_size_vs_scipy(vt):
    """BASELINE config 4 at its full size: mean-field normal VB of an MVN target with D = 4096 parameters
    (2048 means + 2048 variances, closed-form Hessian), LR covariance of 2048 linear moments.  The dense Cholesky
    factorisation and the 2048-column solve run through the blocked kernels (32 block columns; 256-row steps with
    inverted diagonal blocks in the solve); parity against scipy's cho_factor / cho_solve - what the reference
    runs (``solver_lib.py:27,29``; ``lr_cov_lib.py:106,172``) - at rtol 1e-8."""
    import scipy.linalg
    from oracle import models
    dim, k = 2048, 2048
    rng = np.random.RandomState(4)
    a = rng.normal(size=(dim, dim + 16))
    true_cov = a @ a.T / dim + np.eye(dim)
    true_info = np.linalg.inv(true_cov)
    true_mean = rng.normal(size=dim)
    opt, H = models.mvn_kl_closed_form(true_mean, true_info)
    assert H.shape == (4096, 4096)
    J = rng.normal(size=(k, 2 * dim))
    lr = vt.LinearResponseCovariances(lambda par: par.sum(), opt, hessian_at_opt=H)
    cov = lr.get_lr_covariance_from_jacobians(J, J)
    chol = scipy.linalg.cho_factor(H)
    ref = J @ scipy.linalg.cho_solve(chol, J.T)
    assert_close(cov, ref, rtol=1e-8, atol_scale=1e-11, what='config 4 LR covariance, D = 4096')
    # the solver pieces on their own: factor, 2048-column solve, and a ragged dimension crossing the 256-blocks
    solve = vt.solver_lib.get_cholesky_solver(H)
    assert_close(solve(J.T), scipy.linalg.cho_solve(chol, J.T), rtol=1e-8, atol_scale=1e-11, what='potrs 2048 rhs')
    for d2 in (4096 - 131, 300, 257, 129):
        H2 = H[:d2, :d2]
        B2 = J[:37, :d2].T
        assert_close(vt.solver_lib.get_cholesky_solver(H2)(B2), scipy.linalg.cho_solve(scipy.linalg.cho_factor(H2), B2),
                     rtol=1e-8, atol_scale=1e-11, what='potrs D = {}'.format(d2))
