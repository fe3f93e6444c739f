// Counter-based synthetic data generators (device twin of oracle/models.py:
// synth_design / synth_uniform).  Integer-only up to one final multiply, so a
// row range generated on any GPU equals the numpy oracle bit for bit.
#include "synth.cuh"

namespace vt {

namespace {
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void synth_design_kernel(double* X, long ldx, long row0, long nrows, int ncols, unsigned long long key,
                                    double scale) {
  const long total = nrows * (long)ncols;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / ncols;
    const int c = (int)(e - r * ncols);
    const unsigned long long ctr = (unsigned long long)(row0 + r) * (unsigned long long)ncols + (unsigned long long)c;
    const unsigned long long h = mix64(ctr + key);
    const long s = (long)(h & 0xFFFF) + (long)((h >> 16) & 0xFFFF) + (long)((h >> 32) & 0xFFFF) + (long)(h >> 48);
    X[r * ldx + c] = (double)(s - 131070) * scale;
  }
}

__global__ void synth_uniform_kernel(double* u, long row0, long nrows, unsigned long long key) {
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (long)gridDim.x * blockDim.x) {
    const unsigned long long h = mix64((unsigned long long)(row0 + r) + key);
    u[r] = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  }
}

// y = 1[u < sigmoid(z)]
__global__ void synth_bernoulli_kernel(double* y, const double* z, long row0, long nrows, unsigned long long key) {
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (long)gridDim.x * blockDim.x) {
    const unsigned long long h = mix64((unsigned long long)(row0 + r) + key);
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    const double p = 1.0 / (1.0 + exp(-z[r]));
    y[r] = u < p ? 1.0 : 0.0;
  }
}
}  // namespace

int synth_design(double* X, long ldx, long row0, long nrows, int ncols, unsigned long long seed, double scale,
                 cudaStream_t stream) {
  VT_REQUIRE(X && nrows >= 0 && ncols >= 1 && ldx >= ncols, "synth_design: bad arguments");
  if (nrows == 0) return VT_OK;
  const long total = nrows * (long)ncols;
  const int grid = (int)((total + 255) / 256 < 148L * 32 ? (total + 255) / 256 : 148L * 32);
  synth_design_kernel<<<grid, 256, 0, stream>>>(X, ldx, row0, nrows, ncols, mix64(seed), scale);
  VT_LAUNCH_CHECK();
  return VT_OK;
}

int synth_uniform(double* u, long row0, long nrows, unsigned long long seed, cudaStream_t stream) {
  VT_REQUIRE(u && nrows >= 0, "synth_uniform: bad arguments");
  if (nrows == 0) return VT_OK;
  const int grid = (int)((nrows + 255) / 256 < 148L * 8 ? (nrows + 255) / 256 : 148L * 8);
  synth_uniform_kernel<<<grid, 256, 0, stream>>>(u, row0, nrows, mix64(seed ^ 0xA5A5A5A5A5A5A5A5ull));
  VT_LAUNCH_CHECK();
  return VT_OK;
}

int synth_bernoulli(double* y, const double* z, long row0, long nrows, unsigned long long seed, cudaStream_t stream) {
  VT_REQUIRE(y && z && nrows >= 0, "synth_bernoulli: bad arguments");
  if (nrows == 0) return VT_OK;
  const int grid = (int)((nrows + 255) / 256 < 148L * 8 ? (nrows + 255) / 256 : 148L * 8);
  synth_bernoulli_kernel<<<grid, 256, 0, stream>>>(y, z, row0, nrows, mix64(seed ^ 0xA5A5A5A5A5A5A5A5ull));
  VT_LAUNCH_CHECK();
  return VT_OK;
}

}  // namespace vt
