// api.cu — error state, device facts
#include <atomic>

#include "common.cuh"

static thread_local char g_err[512] = "";

void pstl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pstl_last_error(void) { return g_err; }

static std::atomic<unsigned long long> g_launches{0};
void pstl_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long pstl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int pstl_version(void) { return 100; }

extern "C" int pstl_device_info(int* sm_count, int* cc) {
  int dev = 0;
  PSTL_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  PSTL_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc) *cc = p.major * 10 + p.minor;
  return PSTL_OK;
}
