// Library-level entry points: version, error string, device gate.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>
#include "common.cuh"

namespace neko {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NEKO_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return NEKO_ECUDA;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("NEKO_PDL");
    cached = (e && atoi(e) == 0) ? 0 : 1;
  }
  return cached == 1;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace neko

extern "C" {

int neko_version(void) { return 100; }

const char* neko_last_error(void) { return neko::g_err; }

int neko_device_check(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return neko::check_cuda(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return neko::check_cuda(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    neko::set_error("neko_b200 needs an sm_100 device (compute capability 10.x), found %d.x", major);
    return NEKO_EDEVICE;
  }
  return NEKO_OK;
}

int neko_sm_count(void) { return neko::sm_count(); }

}  // extern "C"
