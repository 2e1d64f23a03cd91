// lift_math.cuh -- per-point geometry of the lifting path in float64.
//
// Plain C++ with no CUDA intrinsics, marked __host__ __device__ so that
// tests/hostcheck can compile the very same arithmetic with g++ and compare it
// with the oracle on a machine without a GPU.  The library itself only ever runs
// these functions inside CUDA kernels (there is no host code path that calls
// them).  Compile with -fmad=false (nvcc) / -ffp-contract=off (g++): where the
// reference's numpy arithmetic is unfused the kernels must be too; explicit
// fma() is used where fusing is wanted.
#ifndef PB200_LIFT_MATH_CUH_
#define PB200_LIFT_MATH_CUH_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

#define PB200_CAM_STRIDE_ 24

namespace pb200 {

// ---------------------------------------------------------------------------
// Packed camera (layout documented in include/poseb200.h)
// ---------------------------------------------------------------------------
struct Cam {
  double R[9];
  double T[3];
  double fx, fy, cx, cy;
  double k[3];
  double p[2];
};

PB_HD void load_cam(const double* __restrict__ pack, Cam& c) {
#pragma unroll
  for (int i = 0; i < 9; ++i) c.R[i] = pack[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) c.T[i] = pack[9 + i];
  c.fx = pack[12]; c.fy = pack[13]; c.cx = pack[14]; c.cy = pack[15];
  c.k[0] = pack[16]; c.k[1] = pack[17]; c.k[2] = pack[18];
  c.p[0] = pack[19]; c.p[1] = pack[20];
}

// world -> camera frame:  R (X - T)      (lib/multiviews/cameras.py:41,67)
PB_HD void world_to_cam(const Cam& c, const double X[3], double xc[3]) {
  const double dx = X[0] - c.T[0], dy = X[1] - c.T[1], dz = X[2] - c.T[2];
#pragma unroll
  for (int r = 0; r < 3; ++r) xc[r] = fma(c.R[3 * r + 2], dz, fma(c.R[3 * r + 1], dy, c.R[3 * r] * dx));
}

// H36M projection with averaged focal length (lib/multiviews/cameras.py:25-54).
// Operation order follows the numpy expression so results track the reference
// to the last bits (the 3x3 product is the one place numpy goes through BLAS).
// separate_f: `f` = [fx, fy] (unfold_camera_param(camera, avg_f=False), cameras.py:17-18) instead
// of the averaged focal length project_pose uses.
PB_HD void project_h36m(const Cam& c, const double X[3], double& u_px, double& v_px,
                        bool separate_f = false) {
  double xc[3];
  world_to_cam(c, X, xc);
  const double u = xc[0] / xc[2], v = xc[1] / xc[2];
  const double r2 = u * u + v * v;
  const double poly = (c.k[0] * r2 + c.k[1] * (r2 * r2)) + c.k[2] * (r2 * r2 * r2);
  const double gain = (1.0 + poly) + (c.p[0] * v + c.p[1] * u);
  const double f = 0.5 * (c.fx + c.fy);
  u_px = (separate_f ? c.fx : f) * (u * gain + c.p[1] * r2) + c.cx;
  v_px = (separate_f ? c.fy : f) * (v * gain + c.p[0] * r2) + c.cy;
}

// pymvg find2d: pin-hole, OpenCV plumb-bob distortion, separate fx / fy
// (lib/multiviews/triangulate.py:147,210; D = [k0, k1, p0, p1, k2], :34).
PB_HD void project_plumb_bob(const Cam& c, const double X[3], bool distorted,
                             double& u_px, double& v_px) {
  double xc[3];
  world_to_cam(c, X, xc);
  double x = xc[0] / xc[2], y = xc[1] / xc[2];
  if (distorted) {
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2.0 * x * y;
    const double barrel = 1.0 + c.k[0] * r2 + c.k[1] * r4 + c.k[2] * r6;
    const double xd = x * barrel + c.p[0] * a1 + c.p[1] * (r2 + 2.0 * (x * x));
    const double yd = y * barrel + c.p[0] * (r2 + 2.0 * (y * y)) + c.p[1] * a1;
    x = xd; y = yd;
  }
  u_px = x * c.fx + c.cx;
  v_px = y * c.fy + c.cy;
}

// pymvg CameraModel.undistort == 5 fixed-point iterations of cv2.undistortPoints
// with P = K.  no_distortion: D = 0, which still normalises and
// de-normalises the point, exactly like pymvg does.
PB_HD void undistort_px(const Cam& c, bool no_distortion, double u, double v,
                        double& uo, double& vo) {
  const double xd = (u - c.cx) / c.fx, yd = (v - c.cy) / c.fy;
  double x = xd, y = yd;
  if (!no_distortion) {
    const double k1 = c.k[0], k2 = c.k[1], k3 = c.k[2], p1 = c.p[0], p2 = c.p[1];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const double r2 = x * x + y * y;
      const double icd = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
      const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
      const double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
      x = (xd - dx) * icd;
      y = (yd - dy) * icd;
    }
  }
  uo = x * c.fx + c.cx;
  vo = y * c.fy + c.cy;
}

// M = K [R | -R T], the 3x4 matrix of lib/multiviews/triangulate.py:29-36.
PB_HD void proj_matrix(const Cam& c, double M[12]) {
  double t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    t[r] = -((c.R[3 * r] * c.T[0] + c.R[3 * r + 1] * c.T[1]) + c.R[3 * r + 2] * c.T[2]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double r0 = j < 3 ? c.R[j] : t[0];
    const double r1 = j < 3 ? c.R[3 + j] : t[1];
    const double r2 = j < 3 ? c.R[6 + j] : t[2];
    M[j] = c.fx * r0 + c.cx * r2;
    M[4 + j] = c.fy * r1 + c.cy * r2;
    M[8 + j] = r2;
  }
}

// ---------------------------------------------------------------------------
// Linear triangulation: accumulate G = A^T A of the DLT rows
//   x*M[2]-M[0],  y*M[2]-M[1]      (pymvg find3d, triangulate.py:53)
// and take the eigenvector of the smallest eigenvalue by cyclic Jacobi.
// G is kept as the 10 upper-triangle entries.
// ---------------------------------------------------------------------------
struct Sym4 {
  double a00, a01, a02, a03, a11, a12, a13, a22, a23, a33;
};

PB_HD void sym4_zero(Sym4& g) {
  g.a00 = g.a01 = g.a02 = g.a03 = g.a11 = g.a12 = g.a13 = g.a22 = g.a23 = g.a33 = 0.0;
}

PB_HD void sym4_add_row(Sym4& g, const double r[4]) {
  g.a00 = fma(r[0], r[0], g.a00); g.a01 = fma(r[0], r[1], g.a01);
  g.a02 = fma(r[0], r[2], g.a02); g.a03 = fma(r[0], r[3], g.a03);
  g.a11 = fma(r[1], r[1], g.a11); g.a12 = fma(r[1], r[2], g.a12);
  g.a13 = fma(r[1], r[3], g.a13); g.a22 = fma(r[2], r[2], g.a22);
  g.a23 = fma(r[2], r[3], g.a23); g.a33 = fma(r[3], r[3], g.a33);
}

// the two DLT rows of one observation: rows[0..3] = x*M[2]-M[0], rows[4..7] = y*M[2]-M[1]
PB_HD void dlt_rows(const double M[12], double x, double y, double rows[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) rows[j] = x * M[8 + j] - M[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) rows[4 + j] = y * M[8 + j] - M[4 + j];
}

PB_HD void dlt_add_rows(Sym4& g, const double rows[8]) {
  sym4_add_row(g, rows);
  sym4_add_row(g, rows + 4);
}

PB_HD void dlt_add_view(Sym4& g, const double M[12], double x, double y) {
  double rows[8];
  dlt_rows(M, x, y, rows);
  dlt_add_rows(g, rows);
}

// One Jacobi rotation in the (P,Q) plane of the symmetric 4x4 `a` with eigenvector
// accumulation in `v` (columns).  P<Q are compile-time so everything stays in
// registers.  Returns true if a rotation was applied.
template <int P, int Q>
PB_HD bool jacobi_rotate(double (&a)[4][4], double (&v)[4][4]) {
  const double apq = a[P][Q];
  const double app = a[P][P], aqq = a[Q][Q];
  // relative criterion (keeps the small eigenpairs of the graded Gram matrix accurate):
  // |apq| > 1e-17 sqrt(|app aqq|), compared in squares to spare the square root
  if (!(apq * apq > 1.0e-34 * fabs(app * aqq))) return false;
  // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) with theta = (aqq - app) / (2 apq), written with ONE
  // division and one square root:  t = sgn(d h) |h| / (|d| + sqrt(d^2 + h^2)),  d = aqq - app, h = 2 apq
  const double d = aqq - app, h = 2.0 * apq;
  double t = fabs(h) / (fabs(d) + sqrt(fma(d, d, h * h)));
  if ((d < 0.0) != (h < 0.0)) t = -t;
#if defined(__CUDA_ARCH__)
  const double c = rsqrt(fma(t, t, 1.0));
#else
  const double c = 1.0 / sqrt(fma(t, t, 1.0));
#endif
  const double s = t * c;
  a[P][P] = app - t * apq;
  a[Q][Q] = aqq + t * apq;
  a[P][Q] = 0.0;
  a[Q][P] = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r != P && r != Q) {
      const double arp = a[r][P], arq = a[r][Q];
      const double np_ = c * arp - s * arq;
      const double nq_ = s * arp + c * arq;
      a[r][P] = np_; a[P][r] = np_;
      a[r][Q] = nq_; a[Q][r] = nq_;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double vrp = v[r][P], vrq = v[r][Q];
    v[r][P] = c * vrp - s * vrq;
    v[r][Q] = s * vrp + c * vrq;
  }
  return true;
}

// Eigenvector (unit norm, sign arbitrary) of the smallest eigenvalue of G.
PB_HD void sym4_smallest_eigvec(const Sym4& g, double out[4]) {
  double a[4][4] = {{g.a00, g.a01, g.a02, g.a03},
                    {g.a01, g.a11, g.a12, g.a13},
                    {g.a02, g.a12, g.a22, g.a23},
                    {g.a03, g.a13, g.a23, g.a33}};
  double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    bool any = false;
    any |= jacobi_rotate<0, 1>(a, v);
    any |= jacobi_rotate<2, 3>(a, v);
    any |= jacobi_rotate<0, 2>(a, v);
    any |= jacobi_rotate<1, 3>(a, v);
    any |= jacobi_rotate<0, 3>(a, v);
    any |= jacobi_rotate<1, 2>(a, v);
    if (!any) break;
  }
  int m = 0;
  double lo = a[0][0];
  if (a[1][1] < lo) { lo = a[1][1]; m = 1; }
  if (a[2][2] < lo) { lo = a[2][2]; m = 2; }
  if (a[3][3] < lo) { lo = a[3][3]; m = 3; }
#pragma unroll
  for (int r = 0; r < 4; ++r)
    out[r] = m == 0 ? v[r][0] : (m == 1 ? v[r][1] : (m == 2 ? v[r][2] : v[r][3]));
}

PB_HD void dlt_solve(const Sym4& g, double X[3]) {
  double e[4];
  sym4_smallest_eigvec(g, e);
  X[0] = e[0] / e[3];
  X[1] = e[1] / e[3];
  X[2] = e[2] / e[3];
}

// ---------------------------------------------------------------------------
// Crop affine (lib/utils/transforms.py:76-109): float32 point triples, then
// cv2.getAffineTransform's 6x6 elimination with partial pivoting in float64.
// ---------------------------------------------------------------------------
// What get_affine_transform is given besides centre and scale.  numpy's promotion rules decide in
// which precision each sum of :94-95 is rounded before it is stored into the float32 `src` array,
// so the dtypes travel with the values.
struct CropSpec {
  double sn, cs;       // np.sin / np.cos of np.pi * rot / 180, evaluated on the host (float64)
  double shift[2];     // `shift` (default float32 zeros)
  bool shift_f64;
};

PB_HD CropSpec crop_spec_plain() {
  CropSpec q;
  q.sn = 0.0; q.cs = 1.0; q.shift[0] = q.shift[1] = 0.0; q.shift_f64 = false;
  return q;
}

// scale_px = scale * 200.0 in the dtype of `scale` (:84), returned as doubles
PB_HD void crop_points(double cx, double cy, bool c_f64, double sw, double sh, bool s_f64,
                       const CropSpec& q, float S[3][2], int out_w, int out_h, float D[3][2]) {
  // t = scale_px * shift: float32 product unless either operand is float64
  double tx, ty;
  const bool t_f64 = s_f64 || q.shift_f64;
  if (t_f64) { tx = sw * q.shift[0]; ty = sh * q.shift[1]; }
  else { tx = (double)((float)sw * (float)q.shift[0]); ty = (double)((float)sh * (float)q.shift[1]); }
  // src[0] = center + t (:94): one float32 add when both are float32, else float64
  if (c_f64 || t_f64) { S[0][0] = (float)(cx + tx); S[0][1] = (float)(cy + ty); }
  else { S[0][0] = (float)cx + (float)tx; S[0][1] = (float)cy + (float)ty; }
  // src_dir = get_dir([0, src_w * -0.5], rot_rad) (:88,128-135): products with the float64 sin/cos
  const double half = sw * -0.5;
  const double dir0 = 0.0 * q.cs - half * q.sn;
  const double dir1 = 0.0 * q.sn + half * q.cs;
  // src[1] = center + src_dir + t (:95): `center + list of float64` is float64 whatever center is
  S[1][0] = (float)((cx + dir0) + tx);
  S[1][1] = (float)((cy + dir1) + ty);
  // get_3rd_point (:123-125): float32 arithmetic
  const float dsx = S[0][0] - S[1][0], dsy = S[0][1] - S[1][1];
  S[2][0] = S[1][0] + (-dsy);
  S[2][1] = S[1][1] + dsx;
  const double dw = (double)out_w, dh = (double)out_h;
  D[0][0] = (float)(dw * 0.5);
  D[0][1] = (float)(dh * 0.5);
  D[1][0] = (float)(dw * 0.5 + 0.0);
  D[1][1] = (float)(dh * 0.5 + (double)((float)(dw * -0.5)));
  const float ddx = D[0][0] - D[1][0], ddy = D[0][1] - D[1][1];
  D[2][0] = D[1][0] + (-ddy);
  D[2][1] = D[1][1] + ddx;
}

// Solve  to_k = M [frm_k, 1]  (k = 0..2) the way OpenCV does; m[6] row-major 2x3.
PB_HD void affine_from_triples(const float frm[3][2], const float to[3][2], double m[6]) {
  double a[6][6];
  double b[6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) a[i][j] = 0.0;
  for (int i = 0; i < 3; ++i) {
    const double x = (double)frm[i][0], y = (double)frm[i][1];
    a[2 * i][0] = x; a[2 * i][1] = y; a[2 * i][2] = 1.0;
    a[2 * i + 1][3] = x; a[2 * i + 1][4] = y; a[2 * i + 1][5] = 1.0;
    b[2 * i] = (double)to[i][0];
    b[2 * i + 1] = (double)to[i][1];
  }
  for (int i = 0; i < 6; ++i) {
    int k = i;
    for (int j = i + 1; j < 6; ++j)
      if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
    if (k != i) {
      for (int j = i; j < 6; ++j) { const double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
      const double t = b[i]; b[i] = b[k]; b[k] = t;
    }
    const double d = -1.0 / a[i][i];
    for (int j = i + 1; j < 6; ++j) {
      const double alpha = a[j][i] * d;
      for (int c = i + 1; c < 6; ++c) a[j][c] = a[j][c] + alpha * a[i][c];
      b[j] = b[j] + alpha * b[i];
    }
  }
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
    for (int c = i + 1; c < 6; ++c) s = s - a[i][c] * b[c];
    b[i] = s / a[i][i];
  }
  for (int i = 0; i < 6; ++i) m[i] = b[i];
}

// (scale * 200.0)[k] in the dtype of `scale` (lib/utils/transforms.py:84-85)
PB_HD double crop_scale_px(const void* scale, int is_f64, int row, int k) {
  if (is_f64) return ((const double*)scale)[2 * row + k] * 200.0;
  return (double)(((const float*)scale)[2 * row + k] * 200.0f);
}

PB_HD void crop_affine_row(const void* center, int c_f64, const void* scale, int s_f64, int row,
                           const CropSpec& q, int out_w, int out_h, int inv, double m[6]) {
  double cx, cy;
  if (c_f64) { cx = ((const double*)center)[2 * row]; cy = ((const double*)center)[2 * row + 1]; }
  else { cx = (double)((const float*)center)[2 * row]; cy = (double)((const float*)center)[2 * row + 1]; }
  float S[3][2], D[3][2];
  crop_points(cx, cy, c_f64 != 0, crop_scale_px(scale, s_f64, row, 0), crop_scale_px(scale, s_f64, row, 1),
              s_f64 != 0, q, S, out_w, out_h, D);
  if (inv) affine_from_triples(D, S, m);
  else affine_from_triples(S, D, m);
}

// ---------------------------------------------------------------------------
// RPSM per-sample arithmetic
// ---------------------------------------------------------------------------
// numpy.linspace(-size/2, size/2, n)[k] + centre   (lib/multiviews/pictorial.py:108-113)
PB_HD double grid_coord(double size, int n, int k, double centre) {
  const double start = -size / 2, stop = size / 2;
  double g;
  if (n == 1) g = start;
  else if (k == n - 1) g = stop;
  else g = (double)k * ((stop - start) / (double)(n - 1)) + start;
  return g + centre;
}

// world point -> heatmap pixel coordinates (pictorial.py:168-174):
// project_pose, crop affine (image -> crop), then * [w,h] / img_size.
PB_HD void grid_to_heatmap(const Cam& c, const double aff[6], const double X[3], int w, int h,
                           double img_w, double img_h, double& hx, double& hy) {
  double u, v;
  project_h36m(c, X, u, v);
  const double ax = fma(v, aff[1], u * aff[0]) + aff[2];
  const double ay = fma(v, aff[4], u * aff[3]) + aff[5];
  hx = ax * (double)w / img_w;
  hy = ay * (double)h / img_h;
}

// scipy RegularGridInterpolator(linear, bounds_error=False, fill_value=0) on float32
// values, term order and weights as scipy's generic path (pictorial.py:176-187), in two halves: the
// part that depends only on the sample position (shared by all joints of a view) and the part that
// reads the map.
//   pos >= 0 : y0 * w + x0 of the top-left tap;  kBilinearOutside : the sample is 0;
//   kBilinearNaN : a NaN coordinate gives NaN (fx then carries x + y)
constexpr int kBilinearOutside = -1;
constexpr int kBilinearNaN = -2;

PB_HD void bilinear_prepare(int w, int h, double x, double y, int& pos, double& fx, double& fy) {
  fx = 0.0; fy = 0.0;
  if (x != x || y != y) { pos = kBilinearNaN; fx = x + y; return; }
  if (x < 0.0 || x > (double)(w - 1) || y < 0.0 || y > (double)(h - 1)) { pos = kBilinearOutside; return; }
  int x0 = (int)floor(x), y0 = (int)floor(y);
  if (x0 > w - 2) x0 = w - 2;
  if (y0 > h - 2) y0 = h - 2;
  if (x0 < 0) x0 = 0;
  if (y0 < 0) y0 = 0;
  fx = x - (double)x0;
  fy = y - (double)y0;
  pos = y0 * w + x0;
}

// load(k): the float32 map value at flat index k
template <typename LoadF>
PB_HD double bilinear_apply(LoadF load, int w, int pos, double fx, double fy) {
  if (pos < 0) return pos == kBilinearNaN ? fx : 0.0;
  const double gx = 1.0 - fx, gy = 1.0 - fy;
  double val = 0.0;
  val = val + (double)load(pos) * (gx * gy);
  val = val + (double)load(pos + w) * (gx * fy);
  val = val + (double)load(pos + 1) * (fx * gy);
  val = val + (double)load(pos + w + 1) * (fx * fy);
  return val;
}

template <typename LoadF>
PB_HD double bilinear_zero_outside(LoadF load, int w, int h, double x, double y) {
  int pos;
  double fx, fy;
  bilinear_prepare(w, h, x, y, pos, fx, fy);
  return bilinear_apply(load, w, pos, fx, fy);
}

}  // namespace pb200
#endif  // PB200_LIFT_MATH_CUH_
