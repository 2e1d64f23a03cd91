// geometry.cu -- K0 (projection), K2/K3 (triangulate / reproject / RANSAC from 2D
// locations), epipolar residual and MPJPE partial sums.
//
// These kernels work on a few hundred bytes per frame (the pseudo-label pass of
// run/test/test_pseudo_label.py starts from 2D locations, not heatmaps), so they are
// latency / FP64-issue bound, not HBM bound: one thread per (frame, joint), float64
// throughout, 128-thread blocks so that every SM holds many independent joints.
#include "lift.cuh"

namespace pb200 {

template <typename T>
struct XYLoader {
  const T* base;  // &xy[frame*V, j, 0]
  int row_stride; // J*2
  __device__ __forceinline__ void operator()(int v, double& x, double& y) const {
    const T* p = base + (size_t)v * row_stride;
    x = (double)p[0];
    y = (double)p[1];
  }
};

enum { kModeTriangulate = 0, kModeReproject = 1 };

struct GeoParams {
  const double* campack;
  const int32_t* cam_index;
  const void* xy;
  const uint8_t* vis;
  int B, V, J;
  int no_dist;
  double reproj_thre;
  int num_inliers;
  double* out_X;
  double* out_proj;
  uint8_t* out_vis;
  double* out_err;
  // lift_after_decode only: visibility = maxval > conf_thre, float32 error output
  const float* maxval;
  int use_conf;
  float conf_thre;
  float* out_err32;
  // optional epipolar residuals of the same coordinates, [B, V(V-1), J]
  const double* fmat;
  const int32_t* subj;
  double* out_resid;
};

template <typename T, int kMode>
__global__ void __launch_bounds__(128) geometry_kernel(GeoParams p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)p.B * p.J) return;
  const int f = (int)(t / p.J), j = (int)(t % p.J);
  const int V = p.V, J = p.J;
  const size_t row0 = (size_t)f * V;
  const int32_t* cam_row = p.cam_index + row0;
  XYLoader<T> xy{reinterpret_cast<const T*>(p.xy) + (row0 * J + j) * 2, J * 2};
  uint32_t mask = 0u;
  for (int v = 0; v < V; ++v) {
    bool on = p.vis == nullptr || p.vis[(row0 + v) * J + j];
    if (p.use_conf) on = p.maxval[(row0 + v) * J + j] > p.conf_thre;
    if (on) mask |= 1u << v;
  }
  const bool nd = p.no_dist != 0;

  double X[3];
  const int nv = triangulate_joint(p.campack, cam_row, V, nd, mask, xy, X);
  if (p.out_X) {
    double* o = p.out_X + ((size_t)f * J + j) * 3;
    o[0] = X[0]; o[1] = X[1]; o[2] = X[2];
  }
  if (kMode == kModeReproject) {
    for (int v = 0; v < V; ++v) {
      double pu = 0.0, pv = 0.0, e = 0.0;
      if (nv >= 2) e = reproject_view(p.campack, cam_row, v, nd, X, xy, pu, pv);
      const size_t o = (row0 + v) * J + j;
      if (p.out_proj) { p.out_proj[2 * o] = pu; p.out_proj[2 * o + 1] = pv; }
      if (p.out_vis) p.out_vis[o] = nv >= 2 ? 1 : 0;
      if (p.out_err) p.out_err[o] = e;
      if (p.out_err32) p.out_err32[o] = (float)e;
    }
    if (p.out_resid)
      epipolar_joint(p.fmat + (size_t)p.subj[f] * V * V * 9, V, xy,
                     p.out_resid + (size_t)f * V * (V - 1) * J + j, (size_t)J);
  }
}

// ---- RANSAC with warp-level compaction of (joint, view pair) items -----------------------------
// In geometry_kernel<., kModeRansac> every lane walks the pairs of its own joint, so a warp pays for
// the joint with the most visible pairs (6 of 6 at V = 4) even when most joints have one or none -- the
// normal case after the confidence threshold of run/test/test_pseudo_label.py:194.  Here the 32 joints
// of a warp first list their (owner lane, view a, view b) items in shared memory (exclusive scan by
// shuffles), then ALL lanes solve items 32 at a time, and finally every lane picks the winner among
// its own items in itertools.combinations order with the reference's tie rules
// (lib/multiviews/triangulate.py:140-165).  Same arithmetic per pair, same result, no idle lanes.
struct PairItem {
  uint8_t owner, a, b, pad;
};
struct PairResult {
  double mean_err;
  uint32_t in_mask;
  int32_t count;
};

template <typename T>
__global__ void __launch_bounds__(128) ransac_compact_kernel(GeoParams p, int max_pairs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int cap = 32 * max_pairs;
  const int V = p.V, J = p.J;
  // per warp: results and items of its (joint, pair) work list, and the two DLT rows of every
  // (joint, view) -- undistortion and M = K[R|-RT] are evaluated once per observation, not once per pair
  PairResult* res = reinterpret_cast<PairResult*>(smem_raw) + (size_t)warp * cap;
  PairItem* items = reinterpret_cast<PairItem*>(smem_raw + (size_t)nwarps * cap * sizeof(PairResult)) + (size_t)warp * cap;
  double* rows = reinterpret_cast<double*>(smem_raw + (size_t)nwarps * cap * (sizeof(PairResult) + sizeof(PairItem))) +
                 (size_t)warp * 32 * V * 8;
  const long long total = (long long)p.B * J;
  const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane;  // first joint of this warp
  const long long g = g0 + lane;
  const bool valid = g < total;
  const bool nd = p.no_dist != 0;
  uint32_t mask = 0u;
  if (valid) {
    const int f = (int)(g / J), j = (int)(g % J);
    const size_t row0 = (size_t)f * V;
    XYLoader<T> xy{reinterpret_cast<const T*>(p.xy) + (row0 * J + j) * 2, J * 2};
    for (int v = 0; v < V; ++v) {
      if (!(p.vis == nullptr || p.vis[(row0 + v) * J + j])) continue;
      mask |= 1u << v;
      Cam c;
      load_cam(p.campack + (size_t)p.cam_index[row0 + v] * PB200_CAM_STRIDE, c);
      double M[12], u, w, ox, oy, r8[8];
      proj_matrix(c, M);
      xy(v, ox, oy);
      undistort_px(c, nd, ox, oy, u, w);
      dlt_rows(M, u, w, r8);
      double* dst = rows + ((size_t)lane * V + v) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = r8[k];
    }
  }
  const int nvis = __popc(mask);
  const int npairs = nvis * (nvis - 1) / 2;
  // exclusive scan of npairs over the warp
  int incl = npairs;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const int base = incl - npairs;
  const int n_items = __shfl_sync(0xffffffffu, incl, 31);
  {
    int k = base;
    for (int a = 0; a < V; ++a) {
      if (!((mask >> a) & 1u)) continue;
      for (int b = a + 1; b < V; ++b) {
        if (!((mask >> b) & 1u)) continue;
        PB_DCHECK(k < cap, kDbgRansacItem);
        items[k++] = PairItem{(uint8_t)lane, (uint8_t)a, (uint8_t)b, 0};   // combinations order
      }
    }
  }
  __syncwarp();
  for (int it = lane; it < n_items; it += 32) {
    const PairItem q = items[it];
    const long long go = g0 + q.owner;
    const int f = (int)(go / J), j = (int)(go % J);
    const size_t row0 = (size_t)f * V;
    const int32_t* cam_row = p.cam_index + row0;
    XYLoader<T> xy{reinterpret_cast<const T*>(p.xy) + (row0 * J + j) * 2, J * 2};
    // same accumulation chain as triangulate_joint on views {a, b}: identical bits
    Sym4 gm;
    sym4_zero(gm);
    dlt_add_rows(gm, rows + ((size_t)q.owner * V + q.a) * 8);
    dlt_add_rows(gm, rows + ((size_t)q.owner * V + q.b) * 8);
    double X[3];
    dlt_solve(gm, X);
    uint32_t in_mask = 0u;
    int count = 0;
    double err_sum = 0.0;
    for (int v = 0; v < V; ++v) {
      double pu, pv;
      const double e = reproject_view(p.campack, cam_row, v, nd, X, xy, pu, pv);
      if (e < p.reproj_thre) { in_mask |= 1u << v; ++count; err_sum += e; }
    }
    PairResult r;
    r.count = count;
    r.in_mask = in_mask;
    r.mean_err = count > 0 ? err_sum / (double)count : 0.0;
    res[it] = r;
  }
  __syncwarp();
  if (!valid) return;
  uint32_t best_mask = 0u;
  int best_count = 0;
  double best_err = 10000.0;
  for (int k = base; k < base + npairs; ++k) {
    const PairResult r = res[k];
    if (r.count < p.num_inliers) continue;
    if (r.count > best_count || (r.count == best_count && r.mean_err < best_err)) {
      best_mask = r.in_mask;
      best_count = r.count;
      best_err = r.mean_err;
    }
  }
  const int f = (int)(g / J), j = (int)(g % J);
  for (int v = 0; v < V; ++v) p.out_vis[((size_t)f * V + v) * J + j] = (best_mask >> v) & 1u;
}

__global__ void project_kernel(const double* __restrict__ campack, int cam_id,
                               const double* __restrict__ pts, int n, int model,
                               double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cam c;
  load_cam(campack + (size_t)cam_id * PB200_CAM_STRIDE, c);
  const double X[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  double u, v;
  if (model == 0 || model == 3) project_h36m(c, X, u, v, model == 3);
  else project_plumb_bob(c, X, model == 1, u, v);
  out[2 * i] = u;
  out[2 * i + 1] = v;
}

// Block-wide sum of one double per thread (fixed tree order), result in thread 0.
template <int kThreads>
__device__ __forceinline__ double block_sum(double v, double* smem) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kThreads / 32; ++w) s += smem[w];
  __syncthreads();
  return s;
}

// thread per (frame, pair, joint); pairs in itertools.permutations(range(V), 2) order
template <typename T, typename TW>
__global__ void __launch_bounds__(256)
epipolar_kernel(const double* __restrict__ fmat, const int32_t* __restrict__ subj,
                const T* __restrict__ xy, const TW* __restrict__ w, int B, int V, int J,
                double* __restrict__ out_resid, double* __restrict__ out_sum) {
  __shared__ double red[8];
  const int P = V * (V - 1);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double r = 0.0;
  if (t < (long long)B * P * J) {
    const int j = (int)(t % J);
    const int pr = (int)((t / J) % P);
    const int f = (int)(t / ((long long)J * P));
    const int a = pr / (V - 1);
    int b = pr % (V - 1);
    if (b >= a) ++b;
    const double* F = fmat + (((size_t)subj[f] * V + a) * V + b) * 9;
    const size_t ra = ((size_t)f * V + a) * J + j, rb = ((size_t)f * V + b) * J + j;
    const double xa = (double)xy[2 * ra], ya = (double)xy[2 * ra + 1];
    const double xb = (double)xy[2 * rb], yb = (double)xy[2 * rb + 1];
    // [x_b, y_b, 1] @ F  (BLAS accumulation order), then sum(. * [x_a, y_a, 1])
    const double t0 = fma(yb, F[3], xb * F[0]) + F[6];
    const double t1 = fma(yb, F[4], xb * F[1]) + F[7];
    const double t2 = fma(yb, F[5], xb * F[2]) + F[8];
    r = fabs((t0 * xa + t1 * ya) + t2);
    if (w != nullptr) r = r * ((double)w[rb] * (double)w[ra]);
    out_resid[t] = r;
  }
  if (out_sum != nullptr) {
    const double s = block_sum<256>(r, red);
    if (threadIdx.x == 0) atomicAdd(out_sum, s);
  }
}

__global__ void __launch_bounds__(256)
mpjpe_kernel(const double* __restrict__ pred, const double* __restrict__ gt, long long n,
             double* __restrict__ out4) {
  __shared__ double red[8];
  double s = 0.0, s2 = 0.0, mx = 0.0, cnt = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double dx = pred[3 * i] - gt[3 * i], dy = pred[3 * i + 1] - gt[3 * i + 1],
                 dz = pred[3 * i + 2] - gt[3 * i + 2];
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    s += d; s2 += d * d; mx = fmax(mx, d); cnt += 1.0;
  }
  const double bs = block_sum<256>(s, red);
  const double bs2 = block_sum<256>(s2, red);
  const double bc = block_sum<256>(cnt, red);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
  if ((threadIdx.x & 31) == 0)  // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(out4 + 2), (unsigned long long)__double_as_longlong(mx));
  if (threadIdx.x == 0) {
    atomicAdd(out4, bs);
    atomicAdd(out4 + 1, bs2);
    atomicAdd(out4 + 3, bc);
  }
}

static int check_geo(const double* campack, const int32_t* cam_index, const void* xy, int xy_dtype,
                     int B, int V, int J) {
  PB_REQUIRE(campack && cam_index && xy, "null input pointer");
  PB_REQUIRE(B >= 0 && J >= 1, "bad shape B=%d J=%d", B, J);
  PB_REQUIRE(V >= 2 && V <= PB200_MAX_VIEWS, "V=%d outside [2,%d]", V, PB200_MAX_VIEWS);
  PB_REQUIRE((xy_dtype | 1) == 1, "xy_dtype must be PB200_F32/PB200_F64");
  return PB200_OK;
}

template <int kMode>
static int launch_geo(const GeoParams& p, int xy_dtype, void* stream) {
  const long long n = (long long)p.B * p.J;
  if (n == 0) return PB200_OK;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (xy_dtype == PB200_F32)
    geometry_kernel<float, kMode><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
  else
    geometry_kernel<double, kMode><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
  PB_LAUNCH_CHECK("geometry_kernel");
  return PB200_OK;
}

// The lifting half of pb200_lift_fused (lift_fused.cu): triangulate and
// reproject the float32 coordinates the decode kernel just wrote.
int launch_lift_after_decode(const double* campack, const int32_t* cam_index, const float* xy,
                             const float* maxval, int use_conf, float conf_thre, int B, int V, int J,
                             int no_dist, double* out_X, float* out_err32, double* out_proj,
                             const double* fmat, const int32_t* subj, double* out_resid, void* stream) {
  GeoParams p{campack, cam_index, xy, nullptr, B, V, J, no_dist, 0.0, 0, out_X, out_proj, nullptr, nullptr,
              maxval, use_conf, conf_thre, out_err32, fmat, subj, out_resid};
  return launch_geo<kModeReproject>(p, PB200_F32, stream);
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_project(const double* campack, int cam_id, const double* pts, int n, int model,
                             double* out, void* stream) {
  PB_REQUIRE(campack && pts && out, "null pointer");
  PB_REQUIRE(n >= 0 && cam_id >= 0, "bad n=%d cam_id=%d", n, cam_id);
  PB_REQUIRE(model >= 0 && model <= 3, "model must be 0 (h36m), 1 (plumb-bob), 2 (pin-hole) or 3 (h36m, fx/fy)");
  if (n == 0) return PB200_OK;
  project_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(campack, cam_id, pts, n, model, out);
  PB_LAUNCH_CHECK("project_kernel");
  return PB200_OK;
}

extern "C" int pb200_triangulate(const double* campack, const int32_t* cam_index, const void* xy,
                                 int xy_dtype, const uint8_t* vis, int B, int V, int J,
                                 int no_distortion, double* out_X, void* stream) {
  int rc = check_geo(campack, cam_index, xy, xy_dtype, B, V, J);
  if (rc != PB200_OK) return rc;
  PB_REQUIRE(out_X, "out_X is null");
  GeoParams p{campack, cam_index, xy, vis, B, V, J, no_distortion, 0.0, 0, out_X, nullptr, nullptr, nullptr,
              nullptr, 0, 0.f, nullptr, nullptr, nullptr, nullptr};
  return launch_geo<kModeTriangulate>(p, xy_dtype, stream);
}

extern "C" int pb200_reproject(const double* campack, const int32_t* cam_index, const void* xy,
                               int xy_dtype, const uint8_t* vis, int B, int V, int J,
                               int no_distortion, double* out_proj, uint8_t* out_vis, double* out_X,
                               double* out_err, void* stream) {
  int rc = check_geo(campack, cam_index, xy, xy_dtype, B, V, J);
  if (rc != PB200_OK) return rc;
  PB_REQUIRE(out_proj && out_vis, "out_proj / out_vis is null");
  GeoParams p{campack, cam_index, xy, vis, B, V, J, no_distortion, 0.0, 0, out_X, out_proj, out_vis, out_err,
              nullptr, 0, 0.f, nullptr, nullptr, nullptr, nullptr};
  return launch_geo<kModeReproject>(p, xy_dtype, stream);
}

extern "C" int pb200_ransac(const double* campack, const int32_t* cam_index, const void* xy,
                            int xy_dtype, const uint8_t* vis, int B, int V, int J, int no_distortion,
                            double reproj_thre, int num_inliers, uint8_t* out_vis, void* stream) {
  int rc = check_geo(campack, cam_index, xy, xy_dtype, B, V, J);
  if (rc != PB200_OK) return rc;
  PB_REQUIRE(out_vis, "out_vis is null");
  PB_REQUIRE(num_inliers >= 1, "num_inliers must be >= 1 (the reference divides by it)");
  GeoParams p{campack, cam_index, xy, vis, B, V, J, no_distortion, reproj_thre, num_inliers,
              nullptr, nullptr, out_vis, nullptr, nullptr, 0, 0.f, nullptr, nullptr, nullptr, nullptr};
  const long long n = (long long)B * J;
  if (n == 0) return PB200_OK;
  const int max_pairs = V * (V - 1) / 2;
  const size_t smem = (size_t)4 * 32 * max_pairs * (sizeof(pb200::PairResult) + sizeof(pb200::PairItem)) +
                      (size_t)4 * 32 * V * 8 * sizeof(double);
  const unsigned blocks = (unsigned)((n + 127) / 128);
  // opt in to the shared-memory size once per device and kernel (V = 8 needs 110 KiB), not on every call
  static PerDevice<size_t> attr_f32, attr_f64;
  size_t* have = (xy_dtype == PB200_F32 ? attr_f32 : attr_f64).slot();
  if (have == nullptr) return PB200_ERR_CUDA;
  if (*have < smem) {
    if (xy_dtype == PB200_F32)
      PB_CUDA(cudaFuncSetAttribute(pb200::ransac_compact_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      PB_CUDA(cudaFuncSetAttribute(pb200::ransac_compact_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *have = smem;
  }
  if (xy_dtype == PB200_F32)
    pb200::ransac_compact_kernel<float><<<blocks, 128, smem, (cudaStream_t)stream>>>(p, max_pairs);
  else
    pb200::ransac_compact_kernel<double><<<blocks, 128, smem, (cudaStream_t)stream>>>(p, max_pairs);
  PB_LAUNCH_CHECK("ransac_compact_kernel");
  return PB200_OK;
}

extern "C" int pb200_epipolar(const double* fmat, const int32_t* subj_index, const void* xy,
                              int xy_dtype, const void* weight, int w_dtype, int B, int V, int J,
                              double* out_resid, double* out_sum, void* stream) {
  PB_REQUIRE(fmat && subj_index && xy && out_resid, "null pointer");
  PB_REQUIRE(B >= 0 && J >= 1 && V >= 2 && V <= PB200_MAX_VIEWS, "bad shape B=%d V=%d J=%d", B, V, J);
  PB_REQUIRE((xy_dtype | 1) == 1 && (w_dtype | 1) == 1, "dtype tags must be PB200_F32/PB200_F64");
  const long long n = (long long)B * V * (V - 1) * J;
  if (n == 0) return PB200_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
#define PB_EPI(T, TW) \
  epipolar_kernel<T, TW><<<blocks, 256, 0, s>>>(fmat, subj_index, (const T*)xy, (const TW*)weight, B, V, J, out_resid, out_sum)
  if (xy_dtype == PB200_F32) { if (w_dtype == PB200_F32) PB_EPI(float, float); else PB_EPI(float, double); }
  else { if (w_dtype == PB200_F32) PB_EPI(double, float); else PB_EPI(double, double); }
#undef PB_EPI
  PB_LAUNCH_CHECK("epipolar_kernel");
  return PB200_OK;
}

extern "C" int pb200_mpjpe_stats(const double* pred, const double* gt, int B, int J, double* out4,
                                 void* stream) {
  PB_REQUIRE(pred && gt && out4, "null pointer");
  PB_REQUIRE(B >= 0 && J >= 1, "bad shape");
  const long long n = (long long)B * J;
  if (n == 0) return PB200_OK;
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm * 8;
  if (blocks > cap) blocks = cap;
  mpjpe_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, n, out4);
  PB_LAUNCH_CHECK("mpjpe_kernel");
  return PB200_OK;
}

// ---- world <-> camera frame (lib/multiviews/cameras.py:57-82) -------------------------
namespace pb200 {
__global__ void frame_change_kernel(const double* __restrict__ R, const double* __restrict__ T,
                                    const double* __restrict__ pts, int n, int to_world,
                                    double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  if (!to_world) {  // R (x - T)
    const double dx = x - T[0], dy = y - T[1], dz = z - T[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) out[3 * i + r] = fma(R[3 * r + 2], dz, fma(R[3 * r + 1], dy, R[3 * r] * dx));
  } else {          // R^T x + T
#pragma unroll
    for (int r = 0; r < 3; ++r) out[3 * i + r] = fma(R[6 + r], z, fma(R[3 + r], y, R[r] * x)) + T[r];
  }
}
}  // namespace pb200

extern "C" int pb200_frame_change(const double* R, const double* T, const double* pts, int n,
                                  int to_world, double* out, void* stream) {
  PB_REQUIRE(R && T && pts && out, "null pointer");
  PB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PB200_OK;
  pb200::frame_change_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(R, T, pts, n, to_world, out);
  PB_LAUNCH_CHECK("frame_change_kernel");
  return PB200_OK;
}

// ---- fundamental matrices from calibrated cameras -----------------------------------------------
// The reference estimates F per (subject, view pair) from data with cv2.findFundamentalMat (LMedS,
// run/test/generate_fundamental_matirx.py:45-57) so that x_b^T F x_a ~ 0.  With calibrated cameras F
// is exact:  F = K_b^-T [t]x R K_a^-1,  R = R_b R_a^T,  t = R_b (C_a - C_b), scaled to unit Frobenius
// norm.  One thread per ordered pair.
namespace pb200 {
__global__ void fundamental_kernel(const double* __restrict__ campack, const int32_t* __restrict__ cam_a,
                                   const int32_t* __restrict__ cam_b, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cam a, b;
  load_cam(campack + (size_t)cam_a[i] * PB200_CAM_STRIDE, a);
  load_cam(campack + (size_t)cam_b[i] * PB200_CAM_STRIDE, b);
  double R[9], t[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      R[3 * r + c] = b.R[3 * r] * a.R[3 * c] + b.R[3 * r + 1] * a.R[3 * c + 1] + b.R[3 * r + 2] * a.R[3 * c + 2];
  const double d[3] = {a.T[0] - b.T[0], a.T[1] - b.T[1], a.T[2] - b.T[2]};
  for (int r = 0; r < 3; ++r) t[r] = b.R[3 * r] * d[0] + b.R[3 * r + 1] * d[1] + b.R[3 * r + 2] * d[2];
  // E = [t]x R
  double E[9];
  for (int c = 0; c < 3; ++c) {
    E[c] = -t[2] * R[3 + c] + t[1] * R[6 + c];
    E[3 + c] = t[2] * R[c] - t[0] * R[6 + c];
    E[6 + c] = -t[1] * R[c] + t[0] * R[3 + c];
  }
  // K^-1 = [[1/fx, 0, -cx/fx], [0, 1/fy, -cy/fy], [0, 0, 1]]
  const double kb[9] = {1.0 / b.fx, 0, -b.cx / b.fx, 0, 1.0 / b.fy, -b.cy / b.fy, 0, 0, 1};
  const double ka[9] = {1.0 / a.fx, 0, -a.cx / a.fx, 0, 1.0 / a.fy, -a.cy / a.fy, 0, 0, 1};
  double T1[9], F[9];
  for (int r = 0; r < 3; ++r)  // K_b^-T E
    for (int c = 0; c < 3; ++c)
      T1[3 * r + c] = kb[r] * E[c] + kb[3 + r] * E[3 + c] + kb[6 + r] * E[6 + c];
  double nrm = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      F[3 * r + c] = T1[3 * r] * ka[c] + T1[3 * r + 1] * ka[3 + c] + T1[3 * r + 2] * ka[6 + c];
      nrm += F[3 * r + c] * F[3 * r + c];
    }
  nrm = sqrt(nrm);
  for (int k = 0; k < 9; ++k) out[9 * (size_t)i + k] = F[k] / nrm;
}

// flag[f] = 1 iff some limb of pose f deviates from its expected length by more than thres * expected
// (run/pose3d/estimate.py:84-96, the trigger of the "combination" mode)
__global__ void limb_break_kernel(const double* __restrict__ poses, const int32_t* __restrict__ edges,
                                  const double* __restrict__ limb, int limb_per_frame, int B, int J, int E,
                                  double thres, uint8_t* __restrict__ flag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  const double* p = poses + (size_t)f * J * 3;
  const double* L = limb + (limb_per_frame ? (size_t)f * E : 0);
  bool bad = false;
  for (int e = 0; e < E; ++e) {
    const double* a = p + 3 * edges[2 * e];
    const double* c = p + 3 * edges[2 * e + 1];
    const double dx = a[0] - c[0], dy = a[1] - c[1], dz = a[2] - c[2];
    const double len = sqrt(dx * dx + dy * dy + dz * dz);
    if (fabs(L[e] - len) > thres * L[e]) bad = true;
  }
  flag[f] = bad ? 1 : 0;
}
}  // namespace pb200

extern "C" int pb200_fundamental(const double* campack, const int32_t* cam_a, const int32_t* cam_b, int n,
                                 double* out_F, void* stream) {
  PB_REQUIRE(campack && cam_a && cam_b && out_F, "null pointer");
  PB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PB200_OK;
  pb200::fundamental_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(campack, cam_a, cam_b, n, out_F);
  PB_LAUNCH_CHECK("fundamental_kernel");
  return PB200_OK;
}

extern "C" int pb200_limb_break(const double* poses, const int32_t* edges, const double* limb,
                                int limb_per_frame, int B, int J, int E, double thres, uint8_t* out_flag,
                                void* stream) {
  PB_REQUIRE(poses && edges && limb && out_flag, "null pointer");
  PB_REQUIRE(B >= 0 && J >= 2 && E >= 1, "bad shape");
  if (B == 0) return PB200_OK;
  pb200::limb_break_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(poses, edges, limb, limb_per_frame,
                                                                         B, J, E, thres, out_flag);
  PB_LAUNCH_CHECK("limb_break_kernel");
  return PB200_OK;
}

PB_DEFINE_DEBUG_READER(geometry)
