// pb_common.cuh -- error reporting and launch helpers shared by the .cu files.
#ifndef PB200_COMMON_CUH_
#define PB200_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/poseb200.h"
#include "lift_math.cuh"

namespace pb200 {

void set_error(const char* fmt, ...);

#define PB_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      pb200::set_error(__VA_ARGS__);   \
      return PB200_ERR_ARG;            \
    }                                  \
  } while (0)

#define PB_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      pb200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                       __FILE__, __LINE__);                                        \
      return PB200_ERR_CUDA;                                                       \
    }                                                                              \
  } while (0)

#define PB_LAUNCH_CHECK(name)                                                      \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      pb200::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));  \
      return PB200_ERR_CUDA;                                                       \
    }                                                                              \
  } while (0)

// Per-device caches: attributes set with cudaFuncSetAttribute and occupancy answers belong to
// the CURRENT device, and one process may drive several (tests, notebooks, torch.cuda.set_device).
constexpr int kMaxDevices = 64;
int current_device_ordinal();   // -1 (and the error text set) when there is none
int cached_sm_count();          // SM count of the current device, -1 on failure

// One value per device ordinal, zero-initialised; `slot()` is null when there is no device.
template <typename T>
struct PerDevice {
  T v[kMaxDevices];
  T* slot() {
    const int d = current_device_ordinal();
    return d < 0 ? nullptr : &v[d];
  }
};

// ---- debug-assert build (-DPB200_DEBUG_CHECKS=1): compute-sanitizer is not available on the GPU pool, so the
// invariants the hand-offs rely on are checked by the kernels themselves in a second build of the library.  A
// failed check bumps one of 16 counters (pb200_debug_violations reads them); the release build compiles the
// checks away.  tests/test_gpu_debug_build.py runs the hot path through that build.
#ifndef PB200_DEBUG_CHECKS
#define PB200_DEBUG_CHECKS 0
#endif
enum {
  kDbgDecodeMapRange = 0,     // a decode warp was handed a map index outside [0, total)
  kDbgDecodeMapCount = 1,     // maps decoded in one launch != N*J
  kDbgRpsmCandAddr = 2,       // an in-grid candidate address outside the source vector
  kDbgRpsmArg = 3,            // a back-tracked bin outside [0, nbins)
  kDbgRpsmUnit = 4,           // a warp task index outside [0, nunits)
  kDbgRpsmStage = 5,          // the stage was read while it held another (frame, joint, group)
  kDbgRpsmList = 6,           // the offset lists overflow their shared-memory area
  kDbgRansacItem = 7,         // a (joint, pair) item index outside the warp's list
};
#if PB200_DEBUG_CHECKS
// one copy per translation unit (the library is built without relocatable device code); every .cu that uses
// PB_DCHECK defines a reader with PB_DEFINE_DEBUG_READER and api.cu adds the copies up
static __device__ int g_debug_violations[16];
#define PB_DCHECK(cond, code)                                       \
  do {                                                              \
    if (!(cond)) atomicAdd(&pb200::g_debug_violations[code], 1);    \
  } while (0)
#define PB_DEFINE_DEBUG_READER(name)                                                                   \
  namespace pb200 {                                                                                    \
  int debug_read_##name(int* acc16, int reset) {                                                       \
    int local[16];                                                                                     \
    if (cudaMemcpyFromSymbol(local, g_debug_violations, sizeof(local)) != cudaSuccess) return -1;      \
    for (int i = 0; i < 16; ++i) acc16[i] += local[i];                                                 \
    if (reset) {                                                                                       \
      int zeros[16] = {0};                                                                             \
      if (cudaMemcpyToSymbol(g_debug_violations, zeros, sizeof(zeros)) != cudaSuccess) return -1;      \
    }                                                                                                  \
    return 0;                                                                                          \
  }                                                                                                    \
  }
#else
#define PB_DCHECK(cond, code) do { } while (0)
#define PB_DEFINE_DEBUG_READER(name)
#endif

struct HmViews {
  const float* ptr[PB200_MAX_VIEWS];
  int n;  // 1 (single [N,J,H,W] tensor) or V (per-view tensors [N/V,J,H,W])
};

__device__ __forceinline__ const float* map_base(const HmViews& hv, int row, int j, int J, int HW) {
  // row is view-minor: row = frame * V + view
  if (hv.n == 1) return hv.ptr[0] + ((size_t)row * J + j) * HW;
  const int view = row % hv.n, frame = row / hv.n;
  return hv.ptr[view] + ((size_t)frame * J + j) * HW;
}

}  // namespace pb200
#endif
