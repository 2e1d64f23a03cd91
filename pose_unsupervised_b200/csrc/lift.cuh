// lift.cuh -- per-(frame, joint) lifting: DLT triangulation, reprojection, RANSAC view
// selection.  One thread owns one joint of one frame; shared by geometry.cu and
// lift_fused.cu.
//
// Reference: lib/multiviews/triangulate.py:43-213 (pymvg find3d / find2d restated in
// csrc/lift_math.cuh).
#ifndef PB200_LIFT_CUH_
#define PB200_LIFT_CUH_

#include "pb_common.cuh"

namespace pb200 {

// Triangulate from the views whose bit is set in vis_mask.  xy(v) -> observation of
// view v as (x, y) doubles.  cam_row: packed-camera ids of the frame's V rows.
// Returns the number of views used; X is zero when fewer than two
// (triangulate.py:95-96).
template <typename XYFn>
__device__ __forceinline__ int triangulate_joint(const double* __restrict__ campack,
                                                 const int32_t* __restrict__ cam_row, int V,
                                                 bool no_dist, uint32_t vis_mask, XYFn xy,
                                                 double X[3]) {
  Sym4 g;
  sym4_zero(g);
  int nv = 0;
  for (int v = 0; v < V; ++v) {
    if (!((vis_mask >> v) & 1u)) continue;
    Cam c;
    load_cam(campack + (size_t)cam_row[v] * PB200_CAM_STRIDE, c);
    double M[12], u, w, ox, oy;
    proj_matrix(c, M);
    xy(v, ox, oy);
    undistort_px(c, no_dist, ox, oy, u, w);
    dlt_add_view(g, M, u, w);
    ++nv;
  }
  X[0] = X[1] = X[2] = 0.0;
  if (nv >= 2) dlt_solve(g, X);
  return nv;
}

// pymvg find2d of X into view v and the pixel distance to the observation.
template <typename XYFn>
__device__ __forceinline__ double reproject_view(const double* __restrict__ campack,
                                                 const int32_t* __restrict__ cam_row, int v,
                                                 bool no_dist, const double X[3], XYFn xy,
                                                 double& pu, double& pv) {
  Cam c;
  load_cam(campack + (size_t)cam_row[v] * PB200_CAM_STRIDE, c);
  project_plumb_bob(c, X, !no_dist, pu, pv);
  double ox, oy;
  xy(v, ox, oy);
  const double dx = pu - ox, dy = pv - oy;
  return sqrt(dx * dx + dy * dy);
}

// RANSAC over view pairs (triangulate.py:102-166).  Returns the bit mask of inlier
// views of the best pair (0 if no pair reaches num_inliers).
template <typename XYFn>
__device__ __forceinline__ uint32_t ransac_joint(const double* __restrict__ campack,
                                                 const int32_t* __restrict__ cam_row, int V,
                                                 bool no_dist, uint32_t vis_mask, XYFn xy,
                                                 double reproj_thre, int num_inliers) {
  if (__popc(vis_mask) < 2) return 0u;
  uint32_t best_mask = 0u;
  int best_count = 0;
  double best_err = 10000.0;
  // itertools.combinations over the visible views, in view order
  for (int a = 0; a < V; ++a) {
    if (!((vis_mask >> a) & 1u)) continue;
    for (int b = a + 1; b < V; ++b) {
      if (!((vis_mask >> b) & 1u)) continue;
      double X[3];
      triangulate_joint(campack, cam_row, V, no_dist, (1u << a) | (1u << b), xy, X);
      uint32_t in_mask = 0u;
      int count = 0;
      double err_sum = 0.0;
      for (int j = 0; j < V; ++j) {
        double pu, pv;
        const double e = reproject_view(campack, cam_row, j, no_dist, X, xy, pu, pv);
        if (e < reproj_thre) {
          in_mask |= 1u << j;
          ++count;
          err_sum += e;
        }
      }
      if (count < num_inliers) continue;
      const double mean_err = err_sum / (double)count;
      if (count > best_count || (count == best_count && mean_err < best_err)) {
        best_mask = in_mask;
        best_count = count;
        best_err = mean_err;
      }
    }
  }
  return best_mask;
}

}  // namespace pb200
#endif
