// Error reporting and device discovery for the C ABI (include/b200moby.h).
#include <cstdarg>
#include <cstdio>
#include "host_util.h"

static thread_local char g_err[512] = "";

b200moby_status b2m_fail(b200moby_status code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool b2m_have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return false; }
  return n > 0;
}

extern "C" {
const char* b200moby_last_error(void) { return g_err; }
int b200moby_abi_version(void) { return B200MOBY_ABI_VERSION; }
int b200moby_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; d++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major >= 10) ok++;
  }
  return ok;
}
}
