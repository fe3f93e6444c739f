"""GPU parity: SparseBlockHessian, the block-arrow solver and the GMM-VB
closed-form block assembly (BASELINE config 3
family)."""
import numpy as np
import pytest
import scipy as sp
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


def test_block_hessian_vs_golden(vt, golden):
    """tests/test_sparse_hessian_lib.py:15-53: G=10 blocks of M=9 scattered indices."""
    from oracle import fixtures
    g = golden('sparse_hessian')
    f, x, inds, _ = fixtures.block_quadratic(10, 3, with_scales=False)
    assert np.array_equal(inds, g['bq_inds'])
    sh = vt.SparseBlockHessian(f, inds)
    hb = sh.get_block_hessian(g['bq_x'])
    assert_close(hb.todense(), g['bq_block_hess'], rtol=1e-8, atol_scale=1e-11)
    with pytest.raises(ValueError):
        vt.SparseBlockHessian(f, np.array([[0, 1], [1, 2]]))      # repeated index (:55-57)
    # pure block-diagonal solve
    b = np.random.RandomState(0).normal(size=len(x))
    hpd = hb + hb                                                 # exercise __add__
    dense = hpd.todense() + 0.0
    # make it positive definite block-wise for the solve test
    hb.blocks = hb.blocks @ hb.blocks.transpose(1, 2) + torch.eye(9, device='cuda', dtype=torch.float64)
    sol = vt.solver_lib.get_cholesky_solver(hb)(b)
    assert_close(sol, np.linalg.solve(hb.todense(), b), rtol=1e-8, atol_scale=1e-10)
    assert_close(dense, 2 * g['bq_block_hess'], rtol=1e-8, atol_scale=1e-11)


def test_full_hessian_with_globals_vs_golden(vt, golden):
    """tests/test_sparse_hessian_lib.py:55-113."""
    from oracle import fixtures
    g = golden('sparse_hessian')
    f, x, inds, ginds = fixtures.block_quadratic(10, 3, with_scales=True)
    sh = vt.SparseBlockHessian(f, inds)
    h = sh.get_block_hessian(g['bqs_x']) + sh.get_global_hessian(g['bqs_x'])
    assert_close(h.todense(), g['bqs_full_hess'], rtol=1e-8, atol_scale=1e-11)
    h2 = sh.get_block_hessian(g['bqs_x']) + sh.get_global_hessian(g['bqs_x'], global_inds=g['bqs_ginds'])
    assert_close(h2.todense(), g['bqs_full_hess'], rtol=1e-8, atol_scale=1e-11)
    assert_close(sh.get_global_hessian(g['bqs_x'], global_inds=ginds).todense(), g['bqs_global_hess'], rtol=1e-8,
                 atol_scale=1e-11)
    assert_close(sh.get_hessian(g['bqs_x'], print_every=1).todense(), g['bqs_full_hess'], rtol=1e-8,
                 atol_scale=1e-11)
    with pytest.raises(ValueError):
        sh.get_global_hessian(g['bqs_x'], global_inds=np.array([0, 91]))   # overlaps a local index (:118-122)


def test_gmm_vb_small_vs_golden(vt, golden):
    """GMM-VB (N=12, K=3, d=2): generic colouring path, the closed-form kernel
    and the block-arrow solve against the reference's SparseBlockHessian +
    SuperLU."""
    from oracle import models
    g = golden('sparse_hessian')
    Xobs, K, x = g['gmm_X'], int(g['gmm_K']), g['gmm_x']
    N, d = Xobs.shape
    inds = models.gmm_vb_sparsity(N, K, d)
    obj = vt.objectives.GMMVBObjective(Xobs, K)
    assert np.array_equal(obj.sparsity_array(), inds)
    # generic path: the structured object is also a plain torch callable
    h_generic = vt.SparseBlockHessian(lambda xx: obj(xx), inds).get_hessian(x)
    assert_close(h_generic.todense(), g['gmm_hess'], rtol=1e-8, atol_scale=1e-11)
    # closed-form kernel path
    h_kernel = vt.SparseBlockHessian(obj, inds).get_hessian(x)
    assert_close(h_kernel.todense(), g['gmm_hess'], rtol=1e-8, atol_scale=1e-11)
    for h in (h_generic, h_kernel):
        sol = vt.solver_lib.get_cholesky_solver(h)(g['gmm_b'])
        assert_close(sol, g['gmm_solve'], rtol=1e-8, atol_scale=1e-10)
    # gradient hook against autodiff, away from the optimum (where it is not ~0)
    x_off = x + 0.3 * np.random.RandomState(1).normal(size=len(x))
    gref = torch.func.grad(models.gmm_vb_objective(Xobs, K))(torch.as_tensor(x_off)).numpy()
    assert_close(obj.vt_grad(x_off), gref, rtol=1e-9, atol_scale=1e-12)


def test_gmm_vb_medium_block_arrow_solve(vt):
    """K=20, d=16 (config 3 block shape: M=19, Dg=320) at N=600: kernel
    assembly vs autodiff Hessian, Schur solve vs dense numpy solve, multi-RHS."""
    from oracle import models
    rng = np.random.RandomState(3)
    N, K, d = 200, 20, 16
    centers = 0.35 * rng.normal(size=(K, d))         # overlapping clusters: responsibilities stay away from 0
    Xobs = centers[rng.randint(K, size=N)] + rng.normal(size=(N, d))
    obj = vt.objectives.GMMVBObjective(Xobs, K, prior_prec=0.5)
    x = models.gmm_vb_fit(Xobs, K, prior_prec=0.5, iters=100)      # near the VB optimum: H is positive definite
    sh = vt.SparseBlockHessian(obj, obj.sparsity_array())
    h = sh.get_hessian(x)
    assert h.blocks.shape == (N, K - 1, K - 1) and h.cross.shape == (N, K - 1, K * d)
    href = torch.func.hessian(models.gmm_vb_objective(Xobs, K, prior_prec=0.5))(torch.as_tensor(x)).numpy()
    assert_close(h.todense(), href, rtol=1e-8, atol_scale=1e-11)
    # The logit parametrisation makes H ill conditioned (eigenvalues ~ min_k r_nk), so the
    # forward error of ANY solver is kappa*eps; parity is asserted on the backward error
    # and against the reference route (scipy COO -> SuperLU) at the matching tolerance.
    B = rng.normal(size=(len(x), 3))
    solve = vt.solver_lib.get_sparse_cholesky_solver(h)
    sol = solve(B)
    resid = href @ sol - B
    assert np.max(np.abs(resid)) < 1e-9 * np.max(np.abs(href)) * np.max(np.abs(sol))
    lu = sp.sparse.linalg.factorized(sp.sparse.csc_matrix(h.tocoo()))
    ref1 = lu(B[:, 1])
    kappa = np.linalg.cond(href)
    assert_close(solve(B[:, 1]), ref1, rtol=0.0, atol_scale=max(1e-8, 50 * kappa * 2.2e-16))


def test_block_arrow_solver_well_conditioned(vt):
    """Config-3 block shape (M=19, Dg=320) on a well conditioned synthetic
    block-arrow matrix: full rtol-1e-8 parity with a dense solve, vector and
    multi-RHS, scattered (non-contiguous) index sets."""
    from vittles_b200.sparse_hessian_lib import BlockArrowHessian
    rng = np.random.RandomState(11)
    G, M, Dg = 300, 19, 320
    d = G * M + Dg
    perm = rng.permutation(d)
    sa = perm[:G * M].reshape(G, M)
    gi = perm[G * M:]
    a = rng.normal(size=(G, M, M))
    blocks = a @ a.transpose(0, 2, 1) / M + np.eye(M)
    cross = 0.01 * rng.normal(size=(G, M, Dg))
    g0 = rng.normal(size=(Dg, Dg))
    hgg = g0 @ g0.T / Dg + 3.0 * np.eye(Dg)
    h = BlockArrowHessian(d, torch.as_tensor(sa, device='cuda'), torch.as_tensor(gi, device='cuda'),
                          blocks=_dev(blocks), cross=_dev(cross), hgg=_dev(hgg))
    dense = h.todense()
    assert np.allclose(dense, dense.T)
    B = rng.normal(size=(d, 2))
    solve = vt.solver_lib.get_cholesky_solver(h)
    ref = np.linalg.solve(dense, B)
    assert_close(solve(B), ref, rtol=1e-8, atol_scale=1e-11)
    out = solve(_dev(B[:, 0]))
    assert out.is_cuda
    assert_close(out, ref[:, 0], rtol=1e-8, atol_scale=1e-11)


def test_scipy_sparse_arrow_is_recognised(vt):
    """a scipy COO block-arrow matrix (the form the reference's SparseBlockHessian
    returns and hands to SuperLU) is routed to the batched block + Schur kernels."""
    import scipy.sparse
    from vittles_b200.sparse_hessian_lib import BlockArrowHessian
    rng = np.random.RandomState(3)
    G, M, Dg = 120, 7, 150
    d = G * M + Dg
    perm = rng.permutation(d)
    sa, gi = perm[:G * M].reshape(G, M), perm[G * M:]
    a = rng.normal(size=(G, M, M))
    blocks = a @ a.transpose(0, 2, 1) / M + np.eye(M)
    cross = 0.02 * rng.normal(size=(G, M, Dg))
    g0 = rng.normal(size=(Dg, Dg))
    hgg = g0 @ g0.T / Dg + 3.0 * np.eye(Dg)
    h = BlockArrowHessian(d, torch.as_tensor(sa, device='cuda'), torch.as_tensor(gi, device='cuda'),
                          blocks=_dev(blocks), cross=_dev(cross), hgg=_dev(hgg))
    coo = h.tocoo()
    assert scipy.sparse.issparse(coo)
    solve = vt.solver_lib.get_cholesky_solver(coo)
    assert hasattr(solve, 'block_arrow')                      # not densified
    assert solve.block_arrow.sparsity_array.shape == (G, M)
    B = rng.normal(size=(d, 2))
    ref = np.linalg.solve(coo.toarray(), B)
    assert_close(solve(B), ref, rtol=1e-8, atol_scale=1e-11)
    assert_close(solve(B[:, 1]), ref[:, 1], rtol=1e-8, atol_scale=1e-11)
    # block diagonal only (reference test shape: 10 blocks of 9), and an unstructured matrix
    bd = scipy.sparse.block_diag([blocks[g] for g in range(10)]).tocoo()
    s2 = vt.solver_lib.get_cholesky_solver(bd)
    assert hasattr(s2, 'block_arrow')
    b2 = rng.normal(size=70)
    assert_close(s2(b2), np.linalg.solve(bd.toarray(), b2), rtol=1e-8, atol_scale=1e-11)
    r = scipy.sparse.random(60, 60, density=0.2, random_state=1)
    spd = (r @ r.T + 5.0 * scipy.sparse.eye(60)).tocoo()
    s3 = vt.solver_lib.get_cholesky_solver(spd)
    assert not hasattr(s3, 'block_arrow')
    assert_close(s3(b2[:60]), np.linalg.solve(spd.toarray(), b2[:60]), rtol=1e-8, atol_scale=1e-11)


def test_block_kernels_directly(vt):
    rng = np.random.RandomState(0)
    for (G, M, Dg) in [(1, 1, 1), (37, 19, 320), (1000, 32, 7), (5, 2, 1030)]:
        a = rng.normal(size=(G, M, M + 2))
        Bk = a @ a.transpose(0, 2, 1) + np.eye(M)
        C = rng.normal(size=(G, M, Dg))
        Lb = vt.ops.block_potrf(_dev(Bk))
        assert_close(Lb, np.linalg.cholesky(Bk), rtol=1e-10, atol_scale=1e-13)
        Z = vt.ops.block_trsm(Lb, _dev(C))
        Lnp = np.linalg.cholesky(Bk)
        Zref = np.stack([np.linalg.solve(Lnp[g], C[g]) for g in range(G)])
        assert_close(Z, Zref, rtol=1e-9, atol_scale=1e-12)
        b = rng.normal(size=(G, M))
        y = vt.ops.block_solve(Lb, _dev(b), transpose=False)
        assert_close(y, np.stack([np.linalg.solve(Lnp[g], b[g]) for g in range(G)]), rtol=1e-9, atol_scale=1e-12)
        yt = vt.ops.block_solve(Lb, _dev(b), transpose=True)
        assert_close(yt, np.stack([np.linalg.solve(Lnp[g].T, b[g]) for g in range(G)]), rtol=1e-9, atol_scale=1e-12)
        Zt = vt.ops.block_trsm(Lb, _dev(C), transpose=True)
        assert_close(Zt, np.stack([np.linalg.solve(Lnp[g].T, C[g]) for g in range(G)]), rtol=1e-9, atol_scale=1e-12)
        Z2 = Z.reshape(G * M, Dg)
        xg = rng.normal(size=Dg)
        u = rng.normal(size=G * M)
        assert_close(vt.ops.tall_gemv(Z2, _dev(xg)), Zref.reshape(G * M, Dg) @ xg, rtol=1e-9, atol_scale=1e-12)
        assert_close(vt.ops.tall_colsum(Z2, _dev(u)), Zref.reshape(G * M, Dg).T @ u, rtol=1e-9, atol_scale=1e-12)
    bad = np.eye(3)[None].repeat(4, 0).copy()
    bad[2, 1, 1] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        vt.ops.block_potrf(_dev(bad))


@pytest.mark.parametrize('K', [1, 2, 3, 40, 130])
def test_block_arrow_solver_many_right_hand_sides(vt, K):
    """(d, K) right-hand sides are solved together (SURVEY 8f-2 / the matrix-RHS call sites
    ``sensitivity_lib.py:226``, ``lr_cov_lib.py:172``): one or two columns through the matrix-vector kernels, more
    through the GEMM engine with Z read twice per solve - every K against a dense solve, shape preserved."""
    from vittles_b200.sparse_hessian_lib import BlockArrowHessian
    rng = np.random.RandomState(100 + K)
    G, M, Dg = 257, 19, 96
    d = G * M + Dg
    perm = rng.permutation(d)
    sa, gi = perm[:G * M].reshape(G, M), perm[G * M:]
    a = rng.normal(size=(G, M, M))
    blocks = a @ a.transpose(0, 2, 1) / M + np.eye(M)
    cross = 0.02 * rng.normal(size=(G, M, Dg))
    g0 = rng.normal(size=(Dg, Dg))
    hgg = g0 @ g0.T / Dg + 3.0 * np.eye(Dg)
    h = BlockArrowHessian(d, torch.as_tensor(sa, device='cuda'), torch.as_tensor(gi, device='cuda'),
                          blocks=_dev(blocks), cross=_dev(cross), hgg=_dev(hgg))
    dense = h.todense()
    B = rng.normal(size=(d, K))
    solve = vt.solver_lib.get_cholesky_solver(h)
    out = solve(B)
    assert out.shape == (d, K)
    assert_close(out, np.linalg.solve(dense, B), rtol=1e-8, atol_scale=1e-11, what='K = {}'.format(K))
    # H @ v without densifying, and the device-side dense form
    assert_close(h @ B, dense @ B, rtol=1e-11, atol_scale=1e-13)
    assert_close(h.to_dense_tensor(), dense, rtol=0, atol_scale=1e-15)


def test_blocks_wider_than_the_batched_kernels_fall_back_to_the_dense_solver(vt):
    """The batched kernels hold one block per warp / CTA (M <= 32); the reference's SuperLU path takes any block
    size (``solver_lib.py:46-48``), so wider blocks go through the dense GPU Cholesky instead of failing."""
    from vittles_b200.sparse_hessian_lib import BlockArrowHessian
    rng = np.random.RandomState(5)
    G, M, Dg = 6, 40, 9
    d = G * M + Dg
    a = rng.normal(size=(G, M, M))
    blocks = a @ a.transpose(0, 2, 1) / M + np.eye(M)
    cross = 0.05 * rng.normal(size=(G, M, Dg))
    hgg = 4.0 * np.eye(Dg)
    sa = np.arange(G * M).reshape(G, M)
    gi = G * M + np.arange(Dg)
    h = BlockArrowHessian(d, torch.as_tensor(sa, device='cuda'), torch.as_tensor(gi, device='cuda'),
                          blocks=_dev(blocks), cross=_dev(cross), hgg=_dev(hgg))
    B = rng.normal(size=(d, 3))
    assert_close(vt.solver_lib.get_cholesky_solver(h)(B), np.linalg.solve(h.todense(), B), rtol=1e-8, atol_scale=1e-11)
