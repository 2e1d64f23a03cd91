// lift_fused.cu -- pb200_lift_fused: decode -> triangulate -> reprojection error (+ epipolar
// residuals), the headline pass of BASELINE.json config 2.
//
// The heatmap tensor is streamed exactly once by decode_tma_kernel (csrc/decode.cu); the
// per-(frame, joint) lift (csrc/geometry.cu::geometry_kernel<float,1>) then reads the 16*V*J bytes
// of decoded coordinates per frame back from L2 and writes the 3D point, the reprojection error
// of every view and -- with a fundamental table -- the V(V-1) algebraic epipolar residuals in
// the same thread.  Both launches go to the caller's stream back to back (and into the caller's
// CUDA graph, if one is being captured).
//
// Round 1 also carried two single-kernel variants in which the warp that decoded a frame's last
// map lifted it in place.  They were slower on B200 (0.79 / 1.13 ms against 0.68 ms: lifting on
// 17 of 32 lanes with long float64 chains takes the warp's TMA ring out of the stream) and their
// fence-free hand-off had no release/acquire pairing, so they were removed rather than repaired;
// the measurements stay in profiles/r01_lift_fused_*.ncu.txt.
#include "decode.cuh"

namespace pb200 {
int fill_views(const float* const* hm_views_host, int n_ptr, int N, HmViews& hv);
int launch_decode(const HmViews& hv, int N, int J, int H, int W, const double* affine, int post_process,
                  float* out_xy, float* out_maxval, int32_t* out_idx, void* stream);
int launch_lift_after_decode(const double* campack, const int32_t* cam_index, const float* xy,
                             const float* maxval, int use_conf, float conf_thre, int B, int V, int J,
                             int no_dist, double* out_X, float* out_err32, double* out_proj,
                             const double* fmat, const int32_t* subj, double* out_resid, void* stream);
}  // namespace pb200

using namespace pb200;

extern "C" int pb200_lift_fused(const float* const* hm_views_host, int n_ptr, int B, int V, int J,
                                int H, int W, const double* affine, int post_process,
                                const double* campack, const int32_t* cam_index, int no_distortion,
                                int use_conf, float conf_thre, float* out_xy, float* out_maxval,
                                int32_t* out_idx, double* out_X, float* out_err, double* out_proj,
                                const double* fmat, const int32_t* subj_index, double* out_resid,
                                void* stream) {
  PB_REQUIRE(B >= 0 && J >= 1 && H >= 1 && W >= 1, "bad shape B=%d J=%d H=%d W=%d", B, J, H, W);
  if (B == 0) return PB200_OK;
  PB_REQUIRE(V >= 2 && V <= PB200_MAX_VIEWS, "V=%d outside [2,%d]", V, PB200_MAX_VIEWS);
  PB_REQUIRE((long long)H * W < (1LL << 24), "map of %dx%d exceeds the float32-exact index range", H, W);
  PB_REQUIRE((long long)B * V * J < (1LL << 31) - (1 << 20), "B*V*J too large for one launch; split the batch");
  PB_REQUIRE(affine && campack && cam_index, "null input pointer");
  PB_REQUIRE(out_xy && out_maxval && out_X && out_err, "null output pointer");
  PB_REQUIRE(n_ptr == 1 || n_ptr == V, "n_ptr must be 1 or V");
  PB_REQUIRE(out_resid == nullptr || (fmat && subj_index), "out_resid needs fmat and subj_index");
  HmViews hv;
  int rc = fill_views(hm_views_host, n_ptr, B * V, hv);
  if (rc != PB200_OK) return rc;
  rc = launch_decode(hv, B * V, J, H, W, affine, post_process, out_xy, out_maxval, out_idx, stream);
  if (rc != PB200_OK) return rc;
  return launch_lift_after_decode(campack, cam_index, out_xy, out_maxval, use_conf, conf_thre, B, V, J,
                                  no_distortion, out_X, out_err, out_proj, fmat, subj_index, out_resid,
                                  stream);
}

// The second half on its own: triangulate + reproject (+ epipolar residuals) the float32 coordinates a
// previous pb200_decode wrote, with joints_vis = maxval > conf_thre.  Callers that need to put work
// between the two launches (parallel.PoseExchange forks the previous step's all-gather there, so that
// it runs under this latency-bound kernel instead of competing with the HBM-bound decode) use
// pb200_decode + pb200_lift_decoded; the results are those of pb200_lift_fused.
extern "C" int pb200_lift_decoded(const double* campack, const int32_t* cam_index, const float* xy,
                                  const float* maxval, int use_conf, float conf_thre, int B, int V, int J,
                                  int no_distortion, double* out_X, float* out_err, double* out_proj,
                                  const double* fmat, const int32_t* subj_index, double* out_resid,
                                  void* stream) {
  PB_REQUIRE(B >= 0 && J >= 1, "bad shape B=%d J=%d", B, J);
  if (B == 0) return PB200_OK;
  PB_REQUIRE(V >= 2 && V <= PB200_MAX_VIEWS, "V=%d outside [2,%d]", V, PB200_MAX_VIEWS);
  PB_REQUIRE(campack && cam_index && xy && maxval, "null input pointer");
  PB_REQUIRE(out_X && out_err, "null output pointer");
  PB_REQUIRE(out_resid == nullptr || (fmat && subj_index), "out_resid needs fmat and subj_index");
  return launch_lift_after_decode(campack, cam_index, xy, maxval, use_conf, conf_thre, B, V, J, no_distortion,
                                  out_X, out_err, out_proj, fmat, subj_index, out_resid, stream);
}
