// FP64 peak microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json has no FP64 entry; SURVEY.md §8(d) asks for the DFMA and
// DMMA issue-rate peaks measured at sustained clocks.  This tool times
// register-resident loops of
//   * DFMA                           (vector FP64 pipe)
//   * mma.sync m8n8k4  f64           (DMMA.8x8x4)
//   * mma.sync m16n8k4 / k8 / k16    (sm_90+ shapes; lowered by ptxas)
// for ~`secs` seconds each and prints one JSON line per variant with TFLOP/s
// and flop/clk/SM.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int reps) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// NT independent accumulator tiles per warp.
template <int NT>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b, int reps) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < NT; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int NT, int KK>
__global__ void __launch_bounds__(256) k_dmma16(double* out, double av, double bv, int reps) {
  double c[NT][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = av + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = bv + i;
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        if (KK == 4) dmma1684(c[i], a, b[0]);
        if (KK == 8) dmma1688(c[i], a, b);
        if (KK == 16) dmma16816(c[i], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static void run(const char* name, F launch, double flops_per_thread_per_rep, int blocks, int threads, double secs) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(1);  // warm-up
  CK(cudaDeviceSynchronize());
  // calibrate
  CK(cudaEventRecord(e0)); launch(4); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  int reps = (int)(secs * 1e3 / (ms / 4.0)); if (reps < 4) reps = 4;
  CK(cudaEventRecord(e0)); launch(reps); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double flops = flops_per_thread_per_rep * (double)reps * blocks * threads;
  int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double tf = flops / (ms * 1e-3) / 1e12;
  printf("{\"variant\": \"%s\", \"tflops\": %.3f, \"ms\": %.2f, \"blocks\": %d, \"threads\": %d, "
         "\"flop_per_clk_per_sm_at_max_clock\": %.2f, \"max_clock_mhz\": %d, \"sms\": %d}\n",
         name, tf, ms, blocks, threads, flops / (ms * 1e-3) / (clk_khz * 1e3) / sms, clk_khz / 1000, sms);
  fflush(stdout);
}

int main(int argc, char** argv) {
  double secs = argc > 1 ? atof(argv[1]) : 1.0;
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc(&out, 8));
  const int T = 256;
  for (int occ = 1; occ <= 4; occ *= 2) {
    int B = sms * occ;
    char nm[64];
    snprintf(nm, 64, "dfma_acc8_occ%d", occ);
    run(nm, [&](int r) { k_dfma<8><<<B, T>>>(out, 1.0000001, 1e-9, r); }, 2.0 * 8 * ITERS, B, T, secs);
    snprintf(nm, 64, "dmma884_t8_occ%d", occ);
    run(nm, [&](int r) { k_dmma884<8><<<B, T>>>(out, 1.0000001, 1e-9, r); }, 512.0 / 32 * 8 * ITERS, B, T, secs);
    snprintf(nm, 64, "dmma884_t16_occ%d", occ);
    run(nm, [&](int r) { k_dmma884<16><<<B, T>>>(out, 1.0000001, 1e-9, r); }, 512.0 / 32 * 16 * ITERS, B, T, secs);
    snprintf(nm, 64, "dmma1684_t8_occ%d", occ);
    run(nm, [&](int r) { k_dmma16<8, 4><<<B, T>>>(out, 1.0000001, 1e-9, r); }, 1024.0 / 32 * 8 * ITERS, B, T, secs);
    snprintf(nm, 64, "dmma1688_t8_occ%d", occ);
    run(nm, [&](int r) { k_dmma16<8, 8><<<B, T>>>(out, 1.0000001, 1e-9, r); }, 2048.0 / 32 * 8 * ITERS, B, T, secs);
    snprintf(nm, 64, "dmma16816_t8_occ%d", occ);
    run(nm, [&](int r) { k_dmma16<8, 16><<<B, T>>>(out, 1.0000001, 1e-9, r); }, 4096.0 / 32 * 8 * ITERS, B, T, secs);
  }
  // 128-thread CTAs (1 warp per SMSP) to see single-warp issue limits
  run("dmma884_t16_1warp_per_smsp", [&](int r) { k_dmma884<16><<<sms, 128>>>(out, 1.0000001, 1e-9, r); },
      512.0 / 32 * 16 * ITERS, sms, 128, secs);
  run("dfma_acc8_1warp_per_smsp", [&](int r) { k_dfma<8><<<sms, 128>>>(out, 1.0000001, 1e-9, r); },
      2.0 * 8 * ITERS, sms, 128, secs);
  return 0;
}
