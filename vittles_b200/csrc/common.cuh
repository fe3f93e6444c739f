// Shared device helpers for the vittles_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include "../../include/vittles_b200.h"
#include <cstdint>
#include <cstdio>

namespace vt {

// ---------------------------------------------------------------- errors ----
// Every C-ABI entry point returns an int status; the message of the last
// failure is kept per thread and read back with vt_last_error().
// Status codes are the VT_OK / VT_ERR_* macros of include/vittles_b200.h.

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define VT_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return ::vt::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define VT_LAUNCH_CHECK()                                                     \
  do {                                                                        \
    ::vt::count_launch();                                                     \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) return ::vt::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define VT_REQUIRE(cond, ...)                                                 \
  do {                                                                        \
    if (!(cond)) { ::vt::set_error(__VA_ARGS__); return VT_ERR_INVALID; } \
  } while (0)

int num_sms();
void count_launch();      // every kernel launch of the library bumps a process-wide counter
long launch_count();

// ------------------------------------------------------------- cp.async ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte async copy, zero-filling bytes beyond src_bytes (0, 8 or 16).
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes));
}
// 8-byte async copy, zero-filling when src_bytes == 0.
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------ FP64 tensor core ----
// D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 (the only FP64 MMA
// shape on sm_100a; the m16n8k* PTX shapes lower to multiples of it).
// Lane l: g = l>>2, t = l&3.  a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// --------------------------------------------------------- warp helpers ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming (read-once) 16-byte global load that does not allocate in L1.
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

}  // namespace vt
