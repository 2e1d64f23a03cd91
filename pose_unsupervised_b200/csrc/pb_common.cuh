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
