// Counter-based synthetic data (synth.cu).
#pragma once
#include "common.cuh"

namespace vt {
int synth_design(double* X, long ldx, long row0, long nrows, int ncols, unsigned long long seed, double scale,
                 cudaStream_t stream);
int synth_uniform(double* u, long row0, long nrows, unsigned long long seed, cudaStream_t stream);
int synth_bernoulli(double* y, const double* z, long row0, long nrows, unsigned long long seed, cudaStream_t stream);
}  // namespace vt
