// Library-level entry points of libconvnp_b200.so: error string, version, device probe.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void cnp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

CNP_API const char* cnp_last_error(void) { return g_err; }

CNP_API int cnp_version(void) { return 100; }

// 0 when the current device can run this library (compute capability 10.x), else -1 / cudaError.
CNP_API int cnp_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cnp_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  CNP_REQUIRE(major == 10, "libconvnp_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
  return 0;
}
