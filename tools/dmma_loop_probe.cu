// Where does the DMMA GEMM main loop lose its 20%?  Incremental variants of the
// inner loop of vittles_b200/csrc/dgemm.cu, timed on all SMs:
//   V1  64 accumulators, 8 A x 4 B register fragments, no memory traffic
//   V2  + fragments re-loaded from shared memory every k4 step (12 LDS.64), double buffered
//   V3  + one __syncthreads per 4 k4 steps
//   V4  + cp.async refill (8 x 16 B per thread per 4 k4 steps) from an L2-resident buffer
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop_probe dmma_loop_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int MT = 8, NT = 4, LDKC = 20, TILE = 128 * LDKC, STAGES = 4;

template <int V, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(double* out, const double* gsrc, int iters, const double* gbig, size_t nrow_tiles) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  for (int i = tid; i < 2 * STAGES * TILE; i += WARPS * 32) sm[i] = 1e-3 * (i % 7);
  __syncthreads();
  const int wm0 = (warp % 2) * 64, wn0 = ((warp / 2) % 4) * 32;
  const int a_off = (wm0 + g) * LDKC + tig, b_off = STAGES * TILE + (wn0 + g) * LDKC + tig;
  double acc[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double fa[2][MT], fb[2][NT];
#pragma unroll
  for (int i = 0; i < MT; ++i) fa[0][i] = fa[1][i] = 1.0 + i + tid * 1e-3;
#pragma unroll
  for (int j = 0; j < NT; ++j) fb[0][j] = fb[1][j] = 1e-9 * (j + 1);
  int stage = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int cur = kk & 1, nxt = cur ^ 1;
      if (V >= 2) {
        if (V >= 3 && kk == 3) {
          if (V >= 4) asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2));
          __syncthreads();
        }
        const int stg = (kk == 3) ? (stage + 1) % STAGES : stage;
        const int k2 = (kk == 3) ? 0 : kk + 1;
        const double* As = sm + stg * TILE + a_off + k2 * 4;
        const double* Bs = sm + stg * TILE + b_off + k2 * 4;
#pragma unroll
        for (int i = 0; i < MT; ++i) fa[nxt][i] = As[i * 8 * LDKC];
#pragma unroll
        for (int j = 0; j < NT; ++j) fb[nxt][j] = Bs[j * 8 * LDKC];
      }
      if (V >= 4 && kk == 0) {
        const int st2 = (stage + STAGES - 1) % STAGES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = tid + i * 256;   // 2048 chunks of 16 B = A and B tiles
          const int row = (c >> 3) & 127, ch = c & 7, op = c >> 10;
          const double* src;
          if (V >= 5) {
            const size_t tile = (size_t)blockIdx.x + (size_t)(it >> 6) * gridDim.x;       // new 128-row tile every 64 iterations
            const size_t k0 = (size_t)(it & 63) * 16;
            src = (op == 0) ? gbig + ((tile & 7) * 128 + row) * 1024 + k0 + ch * 2                       // "Hinv": 8 MB, L2 resident
                            : gbig + (1 << 20) + ((tile % nrow_tiles) * 128 + row) * 1024 + k0 + ch * 2;  // "X": streamed
          } else {
            src = gsrc + ((size_t)blockIdx.x * 4096 + (size_t)((it * 8 + i) & 7) * 512 + (c & 511) * 2);
          }
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(sm + op * STAGES * TILE + st2 * TILE + row * LDKC + ch * 2)), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[cur][i], fb[cur][j]);
    }
    stage = (stage + 1) % STAGES;
  }
  asm volatile("cp.async.wait_group 0;");
  double s = 0;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 12345.678) out[0] = s;
}

template <int V, int WARPS>
void run(const char* name, double* out, const double* gsrc, int sms, const double* gbig = nullptr, size_t nrow_tiles = 1) {
  const int smem = 2 * STAGES * TILE * 8;
  CK(cudaFuncSetAttribute(k<V, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000;
  k<V, WARPS><<<sms, WARPS * 32, smem>>>(out, gsrc, 1000, gbig, nrow_tiles);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k<V, WARPS><<<sms, WARPS * 32, smem>>>(out, gsrc, iters, gbig, nrow_tiles);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double flops = 512.0 * 32 * 4 * (double)iters * WARPS * sms;
  printf("%-40s warps=%d  %.2f ms  %.2f TFLOP/s\n", name, WARPS, ms, flops / (ms * 1e-3) / 1e12);
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc(&out, 8));
  double* gsrc; CK(cudaMalloc(&gsrc, (size_t)sms * 4096 * 8 + 65536)); CK(cudaMemset(gsrc, 0, (size_t)sms * 4096 * 8 + 65536));
  run<1, 8>("V1 regs only", out, gsrc, sms);
  run<2, 8>("V2 + LDS double-buffered", out, gsrc, sms);
  run<3, 8>("V3 + barrier / 4 steps", out, gsrc, sms);
  run<4, 8>("V4 + cp.async refill", out, gsrc, sms);
  const size_t nrow_tiles = 16384;   // 16384 tiles x 128 rows x 8 KB = 16 GB of "X"
  double* gbig; CK(cudaMalloc(&gbig, ((size_t)(1 << 20) + nrow_tiles * 128 * 1024) * 8)); CK(cudaMemset(gbig, 0, ((size_t)(1 << 20) + nrow_tiles * 128 * 1024) * 8));
  run<5, 8>("V5 + realistic apply-like addresses", out, gsrc, sms, gbig, nrow_tiles);

  return 0;
}
