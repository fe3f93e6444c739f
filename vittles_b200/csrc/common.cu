// Error bookkeeping and device queries shared by the C-ABI entry points.
#include "common.cuh"
#include <cstdarg>
#include <cstring>
#include <atomic>

namespace vt {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_last_error; }

static std::atomic<long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error '%s' in %s (%s:%d)", cudaGetErrorString(e), what, file, line);
  return VT_ERR_CUDA;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || cached <= 0)
      cached = 148;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace vt
