// Library-wide plumbing: thread-local error string, launch counter, ABI version.
#include <atomic>
#include <cstdarg>

#include "rvb_common.cuh"

namespace rvb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return RVB_ERR_LAUNCH;
  }
  return RVB_OK;
}

int device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : dev % kMaxDevices;
}

std::mutex& attr_mutex() {
  static std::mutex mu;
  return mu;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace rvb

extern "C" int rvb_abi_version(void) { return RVB_ABI_VERSION; }
extern "C" const char* rvb_last_error(void) { return rvb::g_err; }
extern "C" int64_t rvb_launch_count(void) { return (int64_t)rvb::g_launches.load(std::memory_order_relaxed); }
