// Error reporting and device queries shared by all C-ABI entry points.
#include <stdarg.h>

#include <atomic>
#include <stdio.h>

#include "pn_common.cuh"

namespace pn {
namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};  // kernels launched through the C ABI (autograd's backward runs on another thread)
std::atomic<int> g_sm_reserve{0};      // SMs the persistent kernels leave free (pn_reserve_sms)
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_status(const char* what) {
  const cudaError_t e = cudaGetLastError();
  ++g_launches;
  if (e == cudaSuccess) return 0;
  set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
  return 1;
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      n <= 0)
    n = 148;  // B200
  const int keep = n - g_sm_reserve.load();
  return keep > 0 ? keep : 1;
}
}  // namespace pn

extern "C" const char* pn_last_error(void) { return pn::g_err; }
extern "C" int pn_version(void) { return 100; }
extern "C" long long pn_launch_count(void) { return pn::g_launches.load(); }
extern "C" int pn_reserve_sms(int n) {
  if (n < 0) n = 0;
  return pn::g_sm_reserve.exchange(n);
}
