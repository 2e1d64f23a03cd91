// lift.cuh -- per-(frame, joint) lifting: DLT triangulation and reprojection (the RANSAC view
// selection built on them is in geometry.cu).  One thread owns one joint of one frame; shared by geometry.cu and
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

// Algebraic epipolar residuals |x_b^T F_(a,b) x_a| of one joint for the V(V-1) ordered view pairs
// (itertools.permutations order), lib/core/loss.py:121-127, run/test/test_fund_mtx.py:61-67.
// fmat_subj: the [V][V][9] block of the frame's subject; out[pr * pair_stride] receives pair pr.
template <typename XYFn>
__device__ __forceinline__ void epipolar_joint(const double* __restrict__ fmat_subj, int V, XYFn xy,
                                               double* __restrict__ out, size_t pair_stride) {
  int pr = 0;
  for (int a = 0; a < V; ++a) {
    double xa, ya;
    xy(a, xa, ya);
    for (int b = 0; b < V; ++b) {
      if (b == a) continue;
      double xb, yb;
      xy(b, xb, yb);
      const double* F = fmat_subj + ((size_t)a * V + b) * 9;
      const double t0 = fma(yb, F[3], xb * F[0]) + F[6];
      const double t1 = fma(yb, F[4], xb * F[1]) + F[7];
      const double t2 = fma(yb, F[5], xb * F[2]) + F[8];
      out[(size_t)pr * pair_stride] = fabs((t0 * xa + t1 * ya) + t2);
      ++pr;
    }
  }
}

}  // namespace pb200
#endif
