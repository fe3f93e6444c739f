// Batched small-block kernels for block-arrow Hessians (blockchol.cu).
#pragma once
#include "common.cuh"

namespace vt {

constexpr int BLOCK_MAXM = 32;

// In-place lower Cholesky of G row-major M x M blocks (upper triangle zeroed).
// *info (device int): 0, or 1 + index of a block with a non-positive pivot.
int block_potrf(double* blocks, long G, int M, int* info, cudaStream_t stream);
// C_g <- L_g^{-1} C_g for G blocks of shape (M, Dg).
int block_trsm(const double* Lb, double* C, long G, int M, int Dg, int transpose, cudaStream_t stream);
// b_g <- L_g^{-1} b_g (mode 0) or L_g^{-T} b_g (mode 1), b of shape (G, M).
int block_solve(const double* Lb, double* b, long G, int M, int mode, cudaStream_t stream);
// y = beta * y + alpha * Z x  for a row-major Z (R x Dg) with R very long.
int tall_gemv(const double* Z, long R, int Dg, const double* x, double alpha, double* y, double beta,
              cudaStream_t stream);
// out = alpha * Z^T u + beta * y0 (deterministic two-stage reduction).
size_t tall_colsum_workspace_bytes(int Dg);
int tall_colsum(const double* Z, long R, int Dg, const double* u, double alpha, const double* y0, double beta,
                double* out, double* workspace, size_t workspace_bytes, cudaStream_t stream);
// Closed-form local blocks / cross blocks / responsibilities of the GMM-VB objective.
int gmm_blocks(const double* X, long N, int d, int K, const double* m, const double* rho, const double* log_pi,
               double* blocks, double* cross, double* rmat, double* grad_rho, double* obj_terms,
               cudaStream_t stream);

}  // namespace vt
