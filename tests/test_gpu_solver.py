"""GPU parity: solver_lib and the GEMM engine."""
import warnings

import numpy as np
import pytest
import scipy as sp
import scipy.linalg
import scipy.sparse
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def vt():
    import vittles_b200
    return vittles_b200


def _dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device='cuda')


def test_solver_lib_vs_golden(vt, golden):
    """tests/test_solver_lib.py:11-43 inputs; golden outputs from the reference."""
    g = golden('solver_lib')
    h, v, V = g['h'], g['v'], g['V']
    sl = vt.solver_lib
    assert_close(sl.get_dense_cholesky_solver(h)(v), g['dense_v'])
    assert_close(sl.get_cholesky_solver(h)(V), g['dense_V'])
    h_chol = sp.linalg.cho_factor(h)
    assert_close(sl.get_dense_cholesky_solver(None, h_chol)(v), g['dense_v'], rtol=1e-9)
    hs = sp.sparse.csc_matrix(h)
    assert_close(sl.get_cholesky_solver(hs)(v), g['sparse_v'])
    assert_close(sl.get_sparse_cholesky_solver(hs)(V), g['sparse_V'])
    with pytest.raises(ValueError):
        sl.get_sparse_cholesky_solver(h)
    # CG: default legacy tolerance 1e-5 -> same iterates as scipy, same iteration count
    cg = sl.get_cg_solver(lambda x: h @ x, 10)
    assert_close(cg(v), g['cg_v'], rtol=1e-8, atol_scale=1e-10)
    assert cg.last_iterations == int(g['cg_v_iters'])
    assert_close(sl.get_cg_solver(lambda x: hs @ x, 10, cg_opts={'tol': 1e-12})(v), g['cg_tol_v'], rtol=1e-8,
                 atol_scale=1e-10)
    with pytest.warns(UserWarning, match='CG exited with error code 1'):
        out = sl.get_cg_solver(lambda x: hs @ x, 10, cg_opts={'maxiter': 1})(v)
    assert_close(out, g['cg_maxiter1_v'], rtol=1e-8, atol_scale=1e-10)
    # device tensors in -> device tensors out, matvec called with device tensors
    hd = _dev(h)
    xd = sl.get_cg_solver(lambda x: hd @ x, 10, {'tol': 1e-13})(_dev(v))
    assert xd.is_cuda
    assert_close(xd, np.linalg.solve(h, v), rtol=1e-8)


def test_not_positive_definite_raises(vt):
    h = np.eye(5)
    h[3, 3] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        vt.solver_lib.get_dense_cholesky_solver(h)


@pytest.mark.parametrize('d,k', [(1, 1), (7, 3), (128, 5), (129, 130), (300, 1), (1024, 64), (1500, 257)])
def test_cholesky_vs_scipy(vt, d, k):
    rng = np.random.RandomState(d * 7 + k)
    a = rng.normal(size=(d, d + 5))
    h = a @ a.T / d + np.eye(d)
    B = rng.normal(size=(d, k))
    solve = vt.solver_lib.get_cholesky_solver(h)
    ref = sp.linalg.cho_solve(sp.linalg.cho_factor(h), B)
    assert_close(solve(B), ref, rtol=1e-8, atol_scale=1e-11)
    assert_close(solve(B[:, 0]), ref[:, 0], rtol=1e-8, atol_scale=1e-11)
    L = torch.tril(solve.factor.L).cpu().numpy()
    assert_close(L, np.linalg.cholesky(h), rtol=1e-9, atol_scale=1e-12)


@pytest.mark.parametrize('tile', [0, 64, 128])
@pytest.mark.parametrize('am,bm', [('KC', 'KC'), ('KC', 'KS'), ('KS', 'KC'), ('KS', 'KS')])
@pytest.mark.parametrize('m,n,k', [(1, 1, 1), (5, 7, 3), (128, 128, 16), (130, 257, 1000), (300, 64, 4100)])
def test_gemm_engine(vt, am, bm, m, n, k, tile):
    rng = np.random.RandomState(m + n + k)
    A = rng.normal(size=(m, k))
    B = rng.normal(size=(n, k))
    C0 = rng.normal(size=(m, n))
    rs, cs = rng.normal(size=m), rng.normal(size=n)
    Ad = _dev(A if am == 'KC' else A.T.copy())
    Bd = _dev(B if bm == 'KC' else B.T.copy())
    out = _dev(C0.copy())
    vt.ops.gemm(Ad, Bd, am, bm, alpha=-0.5, beta=2.0, out=out, colscale=_dev(cs), rowscale=_dev(rs), tile=tile)
    ref = -0.5 * (rs[:, None] * (A @ B.T) * cs[None, :]) + 2.0 * C0
    assert_close(out, ref, rtol=1e-10, atol_scale=1e-13)


@pytest.mark.parametrize('tile', [0, 64, 128])
@pytest.mark.parametrize('d,k', [(64, 40), (200, 5000), (320, 70000)])
def test_gemm_lower_mirror_splitk(vt, d, k, tile):
    """Symmetric rank-k update: lower tiles only, mirrored, deterministic split-K."""
    rng = np.random.RandomState(d + k)
    Z = rng.normal(size=(k, d))
    ks = rng.uniform(0.5, 1.5, size=k)
    Zd = _dev(Z)
    out = vt.ops.gemm(Zd, Zd, 'KS', 'KS', kscale=_dev(ks), lower=True, mirror=True, tile=tile)
    out2 = vt.ops.gemm(Zd, Zd, 'KS', 'KS', kscale=_dev(ks), lower=True, mirror=True, tile=tile)
    assert torch.equal(out, out2)
    o = out.cpu().numpy()
    assert np.array_equal(o, o.T)
    assert_close(o, (Z * ks[:, None]).T @ Z, rtol=1e-10, atol_scale=1e-13)


def test_cg_solver_matrix_rhs(vt):
    """Extension over scipy/the reference: a (dim, K) right-hand side is solved column by column."""
    rng = np.random.RandomState(11)
    d = 40
    a = rng.normal(size=(d, d + 2))
    h = a @ a.T / d + np.eye(d)
    hd = _dev(h)
    solve = vt.solver_lib.get_cg_solver(lambda v: hd @ v if isinstance(v, torch.Tensor) else h @ v, d,
                                        cg_opts={'tol': 1e-13})
    B = rng.normal(size=(d, 3))
    assert_close(solve(B), np.linalg.solve(h, B), rtol=1e-8, atol_scale=1e-11)
    out = solve(_dev(B))
    assert isinstance(out, torch.Tensor) and out.is_cuda and out.shape == (d, 3)
    with pytest.raises(ValueError):
        solve(rng.normal(size=(d + 1, 3)))
