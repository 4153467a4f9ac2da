// abi.cu — error reporting, version, device check, launch accounting for libb200unet.
#include "common.cuh"
#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void b200_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void b200_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* b200_last_error(void) { return g_err; }
extern "C" int b200_version(void) { return 100; }
extern "C" int64_t b200_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int b200_check_device(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
  if (p.major != 10) B200_FAIL(B200_ERR_ARCH, "device %d is sm_%d%d; libb200unet is built for sm_100a only", dev, p.major, p.minor);
  return B200_OK;
}
