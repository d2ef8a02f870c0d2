// Error plumbing and device checks of the C ABI (include/atdn_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.h"

namespace atdn {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int require_sm100() {
  static int cached_dev = -1;
  static int cached_res = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error((int)e, "cudaGetDevice: %s", cudaGetErrorString(e));
  if (dev == cached_dev) return cached_res;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  cached_res = (major == 10 && minor == 0)
                   ? 0
                   : set_error(ATDN_ERR_ARCH, "device %d is sm_%d%d; libatdn_b200 contains sm_100a code only (no fallback)",
                               dev, major, minor);
  return cached_res;
}

}  // namespace atdn

extern "C" const char* atdn_last_error(void) { return atdn::g_err; }
extern "C" int atdn_version(void) { return 100; }
extern "C" int atdn_check_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return atdn::set_error((int)e, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  return atdn::require_sm100();
}
