"""GPU tests of the optional TF32 path (tcgen05.mma.kind::tf32, TMEM accumulators,
TMA-fed operands; ``csrc/tgemm.cu``).

This is the north_star's "optional FP32/TF32 path": it is NOT held to the rtol
1e-8 bar of the FP64 path.  Stated tolerances, all normwise (|err| relative to
max|reference|), measured on B200 with tools/tgemm_probe.py:

* against a float64 product of the TF32-rounded operands (what the tensor core
  is asked to compute): 2e-5 - only FP32 accumulation separates the two;
* 'tf32'   against the exact float64 result: 2e-3  (one TF32 rounding per operand, 2^-11);
* 'tf32x3' against the exact float64 result: 1e-4  (hi/lo split; limited by the
  tensor core's FP32 accumulation over up to 4096 terms per segment).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_ROUNDED, TOL_TF32, TOL_X3 = 2e-5, 2e-3, 1e-4


@pytest.fixture(scope='module')
def vt():
    import vittles_b200
    return vittles_b200


def _rnd(*shape, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(*shape, device='cuda', dtype=torch.float64, generator=g)


def _nerr(out, ref):
    return float((out - ref).abs().max() / ref.abs().max())


def _tf32_round(x32):
    b = x32.view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def test_convert_rounds_to_tf32_and_pads(vt):
    X = _rnd(300, 102)
    hi, lo = vt.ops.tf32_convert(X, split=3)
    assert hi.dtype == torch.float32 and hi.shape == (300, 104) and lo.shape == hi.shape
    assert torch.equal(hi[:, :102], _tf32_round(X.float()))          # round to nearest, ties away
    assert bool((hi[:, 102:] == 0).all()) and bool((lo[:, 102:] == 0).all())
    assert _nerr(hi[:, :102].double(), X) < 2.0 ** -11
    assert _nerr(hi[:, :102].double() + lo[:, :102].double(), X) < 2.0 ** -21
    # row scaling (the sqrt(s) weights of the Hessian assembly)
    s = torch.rand(300, device='cuda', dtype=torch.float64)
    hs, none = vt.ops.tf32_convert(X, rowscale=s, sqrt_scale=True)
    assert none is None
    assert torch.equal(hs[:, :102], _tf32_round((X * s.sqrt()[:, None]).float()))


@pytest.mark.parametrize('mode', ['KC', 'KS'])
@pytest.mark.parametrize('shape', [(128, 256, 32), (1024, 768, 1024), (200, 300, 100), (130, 515, 102), (1, 1, 4),
                                   (257, 129, 36)])
def test_tf32_gemm_matches_product_of_rounded_operands(vt, mode, shape):
    M, N, K = shape
    A = _rnd(M, K, seed=1) if mode == 'KC' else _rnd(K, M, seed=1)
    B = _rnd(N, K, seed=2) if mode == 'KC' else _rnd(K, N, seed=2)
    out = vt.ops.tf32_gemm(A, B, mode, mode, precision='tf32')
    Ah = vt.ops.tf32_convert(A)[0][:, :A.shape[1]].double()
    Bh = vt.ops.tf32_convert(B)[0][:, :B.shape[1]].double()
    rounded = Ah @ Bh.T if mode == 'KC' else Ah.T @ Bh
    exact = A @ B.T if mode == 'KC' else A.T @ B
    assert out.shape == (M, N)
    assert _nerr(out, rounded) < TOL_ROUNDED
    assert _nerr(out, exact) < TOL_TF32
    x3 = vt.ops.tf32_gemm(A, B, mode, mode, precision='tf32x3')
    assert _nerr(x3, exact) < TOL_X3


def test_tf32_gemm_scales_and_long_k(vt):
    """alpha, row / column scales in the FP64 epilogue; a long inner dimension
    runs as split-K parts and 4096-term segments flushed to FP64."""
    A, B = _rnd(256, 40000, seed=3), _rnd(384, 40000, seed=4)
    rs, cs = _rnd(256, seed=5), _rnd(384, seed=6)
    out = vt.ops.tf32_gemm(A, B, alpha=-0.5, rowscale=rs, colscale=cs, precision='tf32x3')
    ref = -0.5 * rs[:, None] * cs[None, :] * (A @ B.T)
    assert _nerr(out, ref) < TOL_X3
    At, Bt = A.T.contiguous(), B.T.contiguous()
    out_ks = vt.ops.tf32_gemm(At, Bt, 'KS', 'KS', alpha=-0.5, rowscale=rs, colscale=cs, precision='tf32x3')
    assert _nerr(out_ks, ref) < TOL_X3


@pytest.mark.parametrize('precision,tol', [('tf32', TOL_TF32), ('tf32x3', TOL_X3)])
def test_ij_apply_and_syrk_against_the_fp64_engine(vt, precision, tol):
    N, D = 50000, 1024                 # several conversion chunks would need N > 40960: see the ragged case below
    X = vt.ops.synth_design(11, 0, N, D, 'cuda')
    Hm = _rnd(D, D, seed=7)
    Hinv = Hm @ Hm.T / D + torch.eye(D, device='cuda', dtype=torch.float64)
    resid = _rnd(N, seed=8)
    s = torch.rand(N, device='cuda', dtype=torch.float64) * 0.25
    S64 = vt.ops.ij_apply(Hinv, X, resid)
    S = vt.ops.ij_apply(Hinv, X, resid, precision=precision)
    assert _nerr(S, S64) < tol
    H64 = vt.ops.syrk_weighted(X, s, l2=0.5)
    H = vt.ops.syrk_weighted(X, s, l2=0.5, precision=precision)
    assert torch.equal(H, H.T)                                        # mirrored lower triangle: exactly symmetric
    assert _nerr(H, H64) < tol
    # ragged: D not a multiple of anything, N not a multiple of the chunk or the tile
    N2, D2 = 3001, 77
    X2 = vt.ops.synth_design(12, 0, N2, D2, 'cuda')
    Hinv2 = torch.eye(D2, device='cuda', dtype=torch.float64) + 0.1 * _rnd(D2, D2, seed=9)
    r2 = _rnd(N2, seed=10)
    assert _nerr(vt.ops.ij_apply(Hinv2, X2, r2, precision=precision), vt.ops.ij_apply(Hinv2, X2, r2)) < tol
    s2 = torch.rand(N2, device='cuda', dtype=torch.float64)
    assert _nerr(vt.ops.syrk_weighted(X2, s2, precision=precision), vt.ops.syrk_weighted(X2, s2)) < tol


def test_ij_sensitivities_through_the_api(vt):
    """HyperparameterSensitivityLinearApproximation on a GLMObjective with
    precision='tf32x3' / 'tf32': same API, sensitivities within the stated
    tolerance of the oracle (times the conditioning of H); FP64 stays exact."""
    from oracle import models, sensitivity as osens
    n, d = 4000, 64
    X, y, _ = models.synth_logistic(31, n, d)
    w = np.ones(n)
    theta = models.glm_newton(X, y, w)
    ref = osens.linear_sensitivity(models.glm_objective(X, y), theta, w)
    scale = np.abs(ref['sens']).max()
    for precision, tol in (('f64', 1e-10), ('tf32x3', 2e-4), ('tf32', 5e-3)):
        obj = vt.objectives.GLMObjective(X, y, family='logistic', precision=precision)
        sens = vt.HyperparameterSensitivityLinearApproximation(obj, theta, w)
        S = sens.get_dopt_dhyper()
        assert S.shape == (d, n)
        assert np.abs(S - ref['sens']).max() / scale < tol, precision
        assert np.abs(sens.get_hessian_at_opt() - ref['hessian']).max() / np.abs(ref['hessian']).max() < tol


def test_bad_arguments(vt):
    X = _rnd(64, 8)
    with pytest.raises(ValueError):
        vt.ops.ij_apply(torch.eye(8, device='cuda', dtype=torch.float64), X, _rnd(64), precision='fp16')
    with pytest.raises(ValueError):
        vt.objectives.GLMObjective(X, torch.zeros(64, device='cuda', dtype=torch.float64), precision='bf16')
    with pytest.raises(ValueError):
        vt.ops.tf32_gemm(X, X, 'KC', 'KS')
    with pytest.raises(ValueError):
        vt.ops.tf32_gemm(X, _rnd(64, 12))
