// libt2h.so: version / error / device-info entry points.
#include <stdarg.h>
#include <stdio.h>

#include "t2h_internal.h"

namespace t2h {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int cached[64] = {};  // per device of this process (0 = not queried yet)
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& c = cached[dev & 63];
  if (c <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    c = n;
  }
  return c;
}

}  // namespace t2h

extern "C" {

int t2h_version(void) { return T2H_VERSION; }

const char* t2h_last_error(void) { return t2h::g_err; }

int t2h_device_info(int* cc_major, int* cc_minor, int* num_sms) {
  int dev = 0;
  T2H_CUDA(cudaGetDevice(&dev));
  int ma = 0, mi = 0, n = 0;
  T2H_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  T2H_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  T2H_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  if (num_sms) *num_sms = n;
  return T2H_OK;
}

}  // extern "C"
