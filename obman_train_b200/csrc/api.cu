// Library-level entry points of libobman_b200.so: version and error reporting.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace obman {
static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
}  // namespace obman

extern "C" int obman_version(void) { return 100; }

extern "C" const char* obman_get_last_error(void) { return obman::g_last_error; }

// Number of kernels in the loaded image that were compiled for sm_100a (sanity for the build check).
extern "C" int obman_device_ok(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    obman::set_error("obman_device_ok: no CUDA device");
    (void)cudaGetLastError();
    return OBMAN_ERR_CUDA;
  }
  if (p.major != 10) {
    obman::set_error("obman_device_ok: device is sm_%d%d, this library is built for sm_100a only", p.major, p.minor);
    return OBMAN_ERR_UNSUPPORTED;
  }
  return OBMAN_OK;
}
