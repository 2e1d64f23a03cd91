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

int current_device_ordinal() {
  int dev = -1;
  const cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    set_error("no usable CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range");
    return -1;
  }
  return dev;
}

int cached_sm_count() {
  static int sm[kMaxDevices] = {0};   // one slot per device ordinal: a process may drive several GPUs
  const int dev = current_device_ordinal();
  if (dev < 0) return -1;
  if (sm[dev] > 0) return sm[dev];
  int n = 0;
  const cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess || n <= 0) {
    set_error("no usable CUDA device: %s", cudaGetErrorString(e));
    return -1;
  }
  sm[dev] = n;
  return n;
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
