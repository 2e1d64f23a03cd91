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

#if PB200_DEBUG_CHECKS
int debug_read_decode(int* acc16, int reset);
int debug_read_geometry(int* acc16, int reset);
int debug_read_rpsm(int* acc16, int reset);
#endif

}  // namespace pb200

#if PB200_DEBUG_CHECKS
namespace pb200 {
__global__ void debug_selftest_kernel() { PB_DCHECK(threadIdx.x != 0, 15); }   // fails once, on purpose
int debug_read_api(int* acc16, int reset);
}  // namespace pb200
PB_DEFINE_DEBUG_READER(api)
#endif

extern "C" int pb200_debug_enabled(void) { return PB200_DEBUG_CHECKS; }

/* Debug build only: run one check that fails on purpose (counter 15 goes up by one), to show that the
 * counters are alive.  Returns PB200_ERR_UNSUPPORTED in a release build. */
extern "C" int pb200_debug_selftest(void) {
#if PB200_DEBUG_CHECKS
  pb200::debug_selftest_kernel<<<1, 32>>>();
  PB_LAUNCH_CHECK("debug_selftest_kernel");
  return PB200_OK;
#else
  pb200::set_error("not a debug build");
  return PB200_ERR_UNSUPPORTED;
#endif
}

extern "C" int pb200_debug_violations(int32_t* out16, int reset) {
  PB_REQUIRE(out16 != nullptr, "out16 is null");
  for (int i = 0; i < 16; ++i) out16[i] = 0;
#if PB200_DEBUG_CHECKS
  PB_CUDA(cudaDeviceSynchronize());
  if (pb200::debug_read_decode(out16, reset) != 0 || pb200::debug_read_geometry(out16, reset) != 0 ||
      pb200::debug_read_rpsm(out16, reset) != 0 || pb200::debug_read_api(out16, reset) != 0) {
    pb200::set_error("reading the debug counters failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PB200_ERR_CUDA;
  }
#else
  (void)reset;
#endif
  return PB200_OK;
}

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
