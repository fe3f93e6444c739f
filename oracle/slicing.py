"""CPU model of the INT8 error-free-slicing engine (``vittles_b200/csrc/ogemm.cu``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is not a restatement of reference code - the
reference computes ``cho_solve(h_chol, cross_hess)`` (``solver_lib.py:29``, ``sensitivity_lib.py:226``) and
``X^T diag(s) X`` through autograd (``sensitivity_lib.py:381-383``) in float64 - but the arithmetic model of the
engine that replaces those two contractions on the INT8 tensor cores, in exact integer arithmetic, so that its error
bound can be checked without a GPU and the CUDA kernels can be checked digit for digit.

    x_rk = sigma_r * sum_{s=1..S} d_s[r,k] 2^{-7 s} + sigma_r 2^{-7 S} * (0 <= remainder < 1, sign of x)
    sum_k a_mk b_nk ~= sigma_m tau_n sum_{s+t <= S+1} 2^{-7 (s+t)} (A_s B_t^T)[m,n]
"""
import numpy as np


def slice_rows(x, nslices):
    """(digits (S, rows, cols) int8, scale (rows,) float64): per-row power-of-two scale 2^e > max_k |x_rk| and the
    7-bit fields of the fixed-point value trunc(x 2^(7S) / scale), sign-magnitude (truncation toward zero) - the
    computation of ``ozaki_slice_kernel`` / ``digit_of``."""
    x = np.asarray(x, dtype=np.float64)
    m = np.max(np.abs(x), axis=1)
    e = np.where(m > 0, np.frexp(np.where(m > 0, m, 1.0))[1], 0)          # m = f 2^e, f in [0.5, 1)
    scale = np.ldexp(1.0, e)
    q = np.trunc(np.ldexp(x, (7 * nslices - e)[:, None])).astype(np.int64)  # exact: |q| < 2^(7S) <= 2^56
    a = np.abs(q)
    digits = np.empty((nslices,) + x.shape, dtype=np.int8)
    for s in range(nslices):
        f = (a >> (7 * (nslices - 1 - s))) & 127
        digits[s] = np.where(q < 0, -f, f).astype(np.int8)
    return digits, scale


def reconstruct(digits, scale):
    """scale_r * sum_s d_s 2^{-7 (s+1)} in float64 (exact for S <= 7; one rounding for S = 8)."""
    S = digits.shape[0]
    acc = np.zeros(digits.shape[1:], dtype=np.float64)
    for s in range(S - 1, -1, -1):
        acc += digits[s].astype(np.float64) * 2.0 ** (-7 * (s + 1))
    return acc * scale[:, None]


def sliced_gemm(a, b, nslices):
    """A @ B.T the way the engine evaluates it: exact INT64 digit products, products with the same s + t share an
    accumulator, accumulators 0..2 and 3..S-1 are combined in two INT64 words and one FP64 FMA per element
    (``ogemm_kernel`` epilogue), then the row / column scales."""
    da, sa = slice_rows(a, nslices)
    db, sb = slice_rows(b, nslices)
    groups = [np.zeros((a.shape[0], b.shape[0]), dtype=np.int64) for _ in range(nslices)]
    for s in range(nslices):
        for t in range(nslices - s):
            groups[s + t] += da[s].astype(np.int64) @ db[t].astype(np.int64).T
    for g in groups:
        assert np.max(np.abs(g)) < 2 ** 31, 'INT32 accumulator bound violated (K too large)'
    hi = np.zeros_like(groups[0])
    lo = np.zeros_like(groups[0])
    for g in range(nslices):
        if g < 3:
            hi = hi * 128 + groups[g]
        else:
            lo = lo * 128 + groups[g]
    val = lo.astype(np.float64) * 2.0 ** (-7 * (nslices + 1)) + hi.astype(np.float64) * 2.0 ** -28
    return sa[:, None] * sb[None, :] * val


def error_bound(k, nslices):
    """|sliced - exact| <= bound * sigma_m * tau_n: each operand is truncated by < 2^{-7S} of its scale
    (2 K 2^{-7S} for the two first-order terms) and the dropped digit pairs s + t > S + 1 contribute less than
    S K 2^{-7 (S+2)} 127^2 < S K 2^{-7S}."""
    return (2.0 + nslices) * k * 2.0 ** (-7 * nslices)
