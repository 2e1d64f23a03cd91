// api.cu -- library-level entry points: version, error text, device probe.
#include <stdarg.h>
#include <string.h>

#include "pb_common.cuh"

namespace pb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cached_sm_count() {
  static int sm = 0;
  if (sm > 0) return sm;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    set_error("no usable CUDA device: %s", cudaGetErrorString(e));
    sm = 0;
    return -1;
  }
  return sm;
}

}  // namespace pb200

extern "C" int pb200_version(void) { return PB200_VERSION; }

extern "C" const char* pb200_last_error(void) { return pb200::g_err; }

extern "C" int pb200_sm_count(void) { return pb200::cached_sm_count(); }

extern "C" int pb200_device_check(void) {
  int dev = 0, major = 0;
  PB_CUDA(cudaGetDevice(&dev));
  PB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    pb200::set_error("libposeb200 is built for sm_100a only; current device has compute capability %d.x", major);
    return PB200_ERR_UNSUPPORTED;
  }
  return PB200_OK;
}
