// Error state, version and device query for the es3 C-ABI (include/es3.h).
#include <stdarg.h>
#include "common.cuh"

namespace es3 {
static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_status(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return 2;
}
}  // namespace es3

extern "C" const char* es3_last_error(void) { return es3::g_err; }
extern "C" int es3_version(void) { return 100; }

// Fills sm_count / compute capability of `device`; fails unless it is a CC 10.x (Blackwell) part:
// the library has no other code path.
extern "C" int es3_init(int device, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp p;
  ES3_CHECK_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  ES3_REQUIRE(p.major == 10, "es3 kernels are built for sm_100a only; device %d is CC %d.%d", device, p.major, p.minor);
  return 0;
}
