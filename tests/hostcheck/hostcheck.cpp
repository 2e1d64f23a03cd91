// hostcheck.cpp -- TEST-ONLY host build of csrc/lift_math.cuh.
//
// The container that develops this repo has no GPU.  This file compiles the
// __host__ __device__ arithmetic of the CUDA kernels with g++ so that
// tests/test_hostcheck.py can compare it with the oracle on the CPU before any
// GPU time is spent.  It is built into tests/hostcheck/_hostcheck.so by the test
// itself, is never imported by the pose_unsupervised_b200 package, and is not a
// fallback: the product fails loudly without a CUDA device.
#include "../../pose_unsupervised_b200/csrc/lift_math.cuh"

using namespace pb200;

extern "C" {

void hc_crop_affine(const void* center, int c_f64, const void* scale, int s_f64, const double* rot_sincos,
                    double shift_x, double shift_y, int shift_f64, int n, int out_w, int out_h, int inv,
                    double* out) {
  for (int i = 0; i < n; ++i) {
    CropSpec q = crop_spec_plain();
    if (rot_sincos) { q.sn = rot_sincos[2 * i]; q.cs = rot_sincos[2 * i + 1]; }
    q.shift[0] = shift_x; q.shift[1] = shift_y; q.shift_f64 = shift_f64 != 0;
    crop_affine_row(center, c_f64, scale, s_f64, i, q, out_w, out_h, inv, out + 6 * i);
  }
}

void hc_project(const double* campack, const double* pts, int n, int model, double* out) {
  Cam c;
  load_cam(campack, c);
  for (int i = 0; i < n; ++i) {
    if (model == 0) project_h36m(c, pts + 3 * i, out[2 * i], out[2 * i + 1]);
    else project_plumb_bob(c, pts + 3 * i, model == 1, out[2 * i], out[2 * i + 1]);
  }
}

void hc_undistort(const double* campack, int no_dist, const double* uv, int n, double* out) {
  Cam c;
  load_cam(campack, c);
  for (int i = 0; i < n; ++i) undistort_px(c, no_dist != 0, uv[2 * i], uv[2 * i + 1], out[2 * i], out[2 * i + 1]);
}

// one joint: V observations (camera pack per view), visibility mask
int hc_triangulate(const double* campacks, const double* xy, const unsigned char* vis, int V,
                   int no_dist, double* X) {
  Sym4 g;
  sym4_zero(g);
  int nv = 0;
  for (int v = 0; v < V; ++v) {
    if (vis && !vis[v]) continue;
    Cam c;
    load_cam(campacks + PB200_CAM_STRIDE_ * v, c);
    double M[12], u, w;
    proj_matrix(c, M);
    undistort_px(c, no_dist != 0, xy[2 * v], xy[2 * v + 1], u, w);
    dlt_add_view(g, M, u, w);
    ++nv;
  }
  X[0] = X[1] = X[2] = 0.0;
  if (nv < 2) return nv;
  dlt_solve(g, X);
  return nv;
}

void hc_unary(const float* hm, int w, int h, const double* campack, const double* aff,
              const double* pts, int n, double img_w, double img_h, double* out) {
  Cam c;
  load_cam(campack, c);
  for (int i = 0; i < n; ++i) {
    double hx, hy;
    grid_to_heatmap(c, aff, pts + 3 * i, w, h, img_w, img_h, hx, hy);
    out[i] = bilinear_zero_outside([&](int t) { return hm[t]; }, w, h, hx, hy);
  }
}

void hc_grid(double size, int n, const double* centre, double* out) {
  // bin b <-> (iy = b / n^2, ix = (b / n) % n, iz = b % n)
  for (int b = 0; b < n * n * n; ++b) {
    out[3 * b] = grid_coord(size, n, (b / n) % n, centre[0]);
    out[3 * b + 1] = grid_coord(size, n, b / (n * n), centre[1]);
    out[3 * b + 2] = grid_coord(size, n, b % n, centre[2]);
  }
}
}
