// FP64-grade GEMM on the INT8 tensor cores (tcgen05.mma.kind::i8) by error-free
// slicing - the "Ozaki scheme":
//
//   x_rk = sigma_r * sum_{s=1..S} d_s[r][k] 2^{-7s} + O(sigma_r 2^{-7S}),   d_s in [-127, 127] (int8),
//   sigma_r = 2^ceil(log2 max_k |x_rk|)
//
// Every digit product and every K-sum is exact in INT32, so
//
//   sum_k a_mk b_nk = sigma_m tau_n * sum_{s+t <= S+1} 2^{-7(s+t)} (A_s B_t^T)_mn  +  O(S K 2^{-7S} sigma_m tau_n)
//
// needs S(S+1)/2 INT8 GEMMs whose results are combined in INT64 and FP64 in the
// epilogue.  With S = 7 (49 bits per operand; 28 products) the error is ~1e-11 of
// sigma_m tau_n, far inside the rtol 1e-8 parity bar of the FP64 path, at 4.6x the
// throughput bound of the FP64 DMMA pipe (INT8 dense peak 4.5 POP/s vs 37 TFLOP/s);
// S = 8 gives ~1e-13.  Products whose digits sum to the same power of two share
// one TMEM accumulator (S accumulators of 64 columns: at most 512 columns).
//
// This is an additional engine for the H^{-1} G^T apply (precision 'f64_ozaki');
// the FP64 DMMA engine stays the default.
#pragma once
#include "common.cuh"

namespace vt {

constexpr int OZAKI_MAX_SLICES = 8;
constexpr int OZAKI_MAX_K = 16384;      // K * 127^2 * S < 2^31

// digits of every row of X (rows x cols, FP64): out[s][r][k] (int8, row pitch ldo bytes, a multiple of 16; pad bytes
// are zero), scale_out[r] = sigma_r (* fold[r] if fold != nullptr).
int ozaki_slice(const double* X, long ldx, long rows, int cols, int8_t* out, long ldo, long slice_stride, int nslices,
                double* scale_out, const double* fold, cudaStream_t stream);

// C (M x N, FP64) = alpha * rowscale[m] * colscale[n] * sum_{s+t<=S+1} 2^{-7(s+t)} (A_s B_t^T)(m,n)
int ogemm_launch(int M, int N, int K, const int8_t* A, long lda, long a_slice_stride, const int8_t* B, long ldb,
                 long b_slice_stride, int nslices, double alpha, const double* rowscale, const double* colscale, double* C,
                 long ldc, cudaStream_t stream);

// S (D x N) = -Hinv diag(resid) X^T with FP64-grade accuracy on the INT8 tensor cores.
size_t ij_apply_ozaki_workspace_bytes(long N, int D, int nslices);
int ij_apply_ozaki(const double* Hinv, long ldh, const double* X, long ldx, long N, int D, const double* resid,
                   double* S, long lds, int nslices, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// H (D x D) = X^T diag(s) X (s >= 0) with FP64-grade accuracy on the INT8 tensor cores: per chunk of
// observations the digits of sqrt(s_n) x_ni are written transposed with one power-of-two scale per
// feature, the lower tiles of the chunk's Gram matrix run as split-K parts of at most 16384
// observations (INT32 bound) and are accumulated in FP64.
size_t syrk_ozaki_workspace_bytes(long N, int D, int nslices);
int syrk_ozaki(const double* X, long ldx, long N, int D, const double* s, double* H, long ldh, int nslices,
               void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace vt
