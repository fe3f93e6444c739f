// Dependent-chain latencies of the FP64 ops on the Cholesky pivot path (one warp, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, double seed) {
  const int lane = threadIdx.x;
  double x = seed + lane * 1e-3, y = 1.0 + seed;
  long long t0, t1;
  const int N = 512;
  // DFMA dependent chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); clk[0] = (t1 - t0);
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); clk[1] = (t1 - t0);
  // shuffle (double = 2 x SHFL.32) chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
  t1 = clock64(); clk[2] = (t1 - t0);
  // rsqrt chain
  x = fabs(x) + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
  t1 = clock64(); clk[3] = (t1 - t0);
  // 1/x chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.0;
  t1 = clock64(); clk[4] = (t1 - t0);
  // sqrt chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.0;
  t1 = clock64(); clk[5] = (t1 - t0);
  // independent DFMA throughput (16 accumulators)
  double a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = x + j;
  t0 = clock64();
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = fma(a[j], y, 1e-9);
  }
  t1 = clock64(); clk[6] = (t1 - t0);
#pragma unroll
  for (int j = 0; j < 16; ++j) x += a[j];
  // float rsqrt + 2 Newton steps in double
  x = fabs(x) + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    double r = (double)rsqrtf((float)x);
    r = r * fma(-0.5 * x * r, r, 1.5);
    r = r * fma(-0.5 * x * r, r, 1.5);
    x = r + 1.0;
  }
  t1 = clock64(); clk[7] = (t1 - t0);
  // shared-memory store -> load round trip
  __shared__ double sh[64];
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { sh[lane] = x; __syncwarp(); x = sh[(lane + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); clk[8] = (t1 - t0);
  out[lane] = x;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&clk, 16 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, clk, 0.5);
  long long h[16];
  cudaMemcpy(h, clk, 16 * 8, cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA dep", "DMUL dep", "SHFL.f64 dep", "rsqrt(double)+add dep", "1/x+add dep", "sqrt+add dep",
                         "DFMA indep (per op)", "rsqrtf+2 Newton+add dep", "STS->LDS+add round trip"};
  for (int i = 0; i < 9; ++i) printf("%-28s %.1f clk/op\n", names[i], h[i] / 512.0);
  return 0;
}
