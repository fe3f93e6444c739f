// TF32 tensor-core GEMM engine on tcgen05 / TMEM / TMA (sm_100a): the optional
// reduced-precision path of the north_star ("FP64 DMMA, with an optional
// FP32/TF32 path").  The FP64 engine (dgemm.cu) stays the default everywhere;
// nothing here is used unless the caller asks for precision 'tf32' / 'tf32x3'.
//
//   C(m,n) [f64] = alpha * rowscale[m] * colscale[n] * sum_k A(m,k) B(n,k) (+ C)
//
// A and B are FP32 arrays whose values are already rounded to TF32
// (tf32_convert below).  With the low-order parts A_lo, B_lo present the engine
// evaluates the three-term split  A_hi B_hi + A_hi B_lo + A_lo B_hi  ("tf32x3"),
// which recovers FP32-grade accuracy from TF32 products.
//
// Operand layouts:  KC  element (r,k) at p[r*ld + k]  (k contiguous, "K-major")
//                   KS  element (r,k) at p[k*ld + r]  (k strided,   "MN-major")
// Both are fed by 2-D TMA boxes with the 128-byte swizzle into a 4-stage ring;
// one elected thread issues tcgen05.mma (128 x BN x 8 per instruction) into a
// double-buffered TMEM accumulator; four epilogue warps drain TMEM with
// tcgen05.ld, convert to FP64, scale and store (or accumulate) while the next
// tile's MMAs run.
#pragma once
#include "common.cuh"

namespace vt {

struct TGemmParams {
  int M, N;
  long K;
  const float* A_hi; const float* A_lo; long lda; int amode;   // A_lo / B_lo: nullptr -> plain TF32
  const float* B_hi; const float* B_lo; long ldb; int bmode;
  double* C; long ldc;
  double alpha;
  const double* colscale;
  const double* rowscale;
  int parts;               // split-K factor (>= 1; 0 = automatic); > 1 needs a workspace
  double* workspace; size_t workspace_bytes;
};

int tgemm_launch(const TGemmParams& p, cudaStream_t stream);
size_t tgemm_workspace_bytes(int M, int N, long K, int split);

// FP64 -> TF32 (round to nearest, ties away), stored as FP32: hi = tf32(scale_r * x),
// lo = tf32(scale_r * x - hi) (optional).  `ldo` must be a multiple of 4; columns
// cols..ldo-1 are zero-filled.
int tf32_convert(const double* X, long ldx, long rows, int cols, const double* rowscale, int sqrt_scale, float* hi,
                 float* lo, long ldo, cudaStream_t stream);

// S (D x N) = -Hinv * diag(resid) X^T in TF32 / TF32x3, observations converted chunk by chunk.
size_t ij_apply_tf32_workspace_bytes(long N, int D, int split);
int ij_apply_tf32(const double* Hinv, long ldh, const double* X, long ldx, long N, int D, const double* resid,
                  double* S, long lds, int split, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// H (D x D) = X^T diag(s) X in TF32 / TF32x3 (rows scaled by sqrt(s) during the conversion).
size_t syrk_tf32_workspace_bytes(long N, int D, int split);
int syrk_tf32(const double* X, long ldx, long N, int D, const double* s, double* H, long ldh, int split,
              void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace vt
