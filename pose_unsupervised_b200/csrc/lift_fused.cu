// lift_fused.cu -- decode -> triangulate -> reprojection error in ONE pass over HBM.
//
// BASELINE.json config 2 (B frames x V views x J joints of HxW float32 heatmaps).
// The only large operand is the heatmap tensor; it is streamed once by
// warp-per-map decoders (csrc/decode.cuh).  Lifting a frame needs the V*J decoded
// coordinates of that frame, which live in 8*V*J bytes of L2-resident output, so
// instead of a second launch the warp that completes the LAST map of a frame lifts
// that frame in place: lanes = joints, float64 DLT + Jacobi per lane
// (csrc/lift.cuh), reprojection error per view.
//
// Scheduling: persistent blocks (a whole number per SM), warps pull map indices
// from a global counter, so a warp that is busy lifting simply pulls fewer maps and
// the HBM stream never waits for it.  A per-frame arrival counter (release: decoded
// outputs -> __threadfence -> atomicAdd; acquire: atomicAdd -> __threadfence ->
// ld.cg) finds the completing warp.  All counters are left at zero on exit, so the
// workspace needs zeroing only once.
#include "decode.cuh"
#include "lift.cuh"

namespace pb200 {

constexpr int kFusedWarps = 8;

struct FusedParams {
  HmViews hv;
  int B, V, J, H, W;
  int vec_ok;
  const double* affine;
  int post_process;
  const double* campack;
  const int32_t* cam_index;
  int no_dist;
  int use_conf;
  float conf_thre;
  float* out_xy;
  float* out_maxval;
  int32_t* out_idx;
  double* out_X;
  float* out_err;
  double* out_proj;
  int32_t* ws;  // [0] next map, [1] finished blocks, [2 + f] maps decoded of frame f
};

struct FusedXY {
  const float2* xy;  // &out_xy[frame*V, j]
  int row_stride;    // J
  __device__ __forceinline__ void operator()(int v, double& x, double& y) const {
    const float2 q = __ldcg(xy + (size_t)v * row_stride);
    x = (double)q.x;
    y = (double)q.y;
  }
};

__device__ __noinline__ void lift_frame_joint(const FusedParams& p, int f, int j) {
  const int V = p.V, J = p.J;
  const size_t row0 = (size_t)f * V;
  const int32_t* cam_row = p.cam_index + row0;
  FusedXY xy{reinterpret_cast<const float2*>(p.out_xy) + row0 * J + j, J};
  uint32_t mask = 0u;
  for (int v = 0; v < V; ++v) {
    bool on = true;
    if (p.use_conf) on = __ldcg(p.out_maxval + (row0 + v) * J + j) > p.conf_thre;
    if (on) mask |= 1u << v;
  }
  const bool nd = p.no_dist != 0;
  double X[3];
  const int nv = triangulate_joint(p.campack, cam_row, V, nd, mask, xy, X);
  double* ox = p.out_X + ((size_t)f * J + j) * 3;
  ox[0] = X[0]; ox[1] = X[1]; ox[2] = X[2];
  for (int v = 0; v < V; ++v) {
    double pu = 0.0, pv = 0.0, e = 0.0;
    if (nv >= 2) e = reproject_view(p.campack, cam_row, v, nd, X, xy, pu, pv);
    const size_t o = (row0 + v) * J + j;
    p.out_err[o] = (float)e;
    if (p.out_proj) { p.out_proj[2 * o] = pu; p.out_proj[2 * o + 1] = pv; }
  }
}

__global__ void __launch_bounds__(kFusedWarps * 32, 4) lift_fused_kernel(const FusedParams p) {
  const int lane = threadIdx.x & 31;
  const int J = p.J, V = p.V, HW = p.H * p.W;
  const int per_frame = V * J;
  const int total = p.B * per_frame;
  for (;;) {
    int m = 0;
    if (lane == 0) m = atomicAdd(p.ws, 1);
    m = __shfl_sync(0xffffffffu, m, 0);
    if (m >= total) break;
    const int row = m / J, j = m - row * J;
    const float* base = map_base(p.hv, row, j, J, HW);
    const DecodeOut o = decode_map(base, p.H, p.W, p.vec_ok != 0, p.affine + 6 * (size_t)row,
                                   p.post_process != 0, lane);
    int last = 0;
    const int f = row / V;
    if (lane == 0) {
      reinterpret_cast<float2*>(p.out_xy)[m] = make_float2(o.x, o.y);
      p.out_maxval[m] = o.maxval;
      if (p.out_idx) p.out_idx[m] = o.idx;
      __threadfence();  // release the decoded outputs
      last = atomicAdd(p.ws + 2 + f, 1) == per_frame - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();  // acquire the other warps' outputs
      for (int jj = lane; jj < J; jj += 32) lift_frame_joint(p, f, jj);
      if (lane == 0) p.ws[2 + f] = 0;  // leave the workspace clean
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(p.ws + 1, 1) == (int)gridDim.x - 1) {  // last block out resets the counters
      p.ws[0] = 0;
      p.ws[1] = 0;
    }
  }
}

int fill_views(const float* const* hm_views_host, int n_ptr, int N, HmViews& hv);
bool views_vec_ok(const HmViews& hv, int HW);

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_lift_workspace_ints(int B) {
  return (size_t)(B < 0 ? 0 : B) + 4;
}

extern "C" int pb200_lift_fused(const float* const* hm_views_host, int n_ptr, int B, int V, int J,
                                int H, int W, const double* affine, int post_process,
                                const double* campack, const int32_t* cam_index, int no_distortion,
                                int use_conf, float conf_thre, float* out_xy, float* out_maxval,
                                int32_t* out_idx, double* out_X, float* out_err, double* out_proj,
                                int32_t* workspace, void* stream) {
  PB_REQUIRE(B >= 0 && J >= 1 && H >= 1 && W >= 1, "bad shape B=%d J=%d H=%d W=%d", B, J, H, W);
  PB_REQUIRE(V >= 2 && V <= PB200_MAX_VIEWS, "V=%d outside [2,%d]", V, PB200_MAX_VIEWS);
  PB_REQUIRE((long long)H * W < (1LL << 24), "map of %dx%d exceeds the float32-exact index range", H, W);
  PB_REQUIRE((long long)B * V * J < (1LL << 31) - 65536, "B*V*J too large for one launch; split the batch");
  PB_REQUIRE(affine && campack && cam_index, "null input pointer");
  PB_REQUIRE(out_xy && out_maxval && out_X && out_err && workspace, "null output / workspace pointer");
  PB_REQUIRE(n_ptr == 1 || n_ptr == V, "n_ptr must be 1 or V");
  FusedParams p;
  int rc = fill_views(hm_views_host, n_ptr, B * V, p.hv);
  if (rc != PB200_OK) return rc;
  if (B == 0) return PB200_OK;
  p.B = B; p.V = V; p.J = J; p.H = H; p.W = W;
  p.vec_ok = views_vec_ok(p.hv, H * W) ? 1 : 0;
  p.affine = affine; p.post_process = post_process;
  p.campack = campack; p.cam_index = cam_index; p.no_dist = no_distortion;
  p.use_conf = use_conf; p.conf_thre = conf_thre;
  p.out_xy = out_xy; p.out_maxval = out_maxval; p.out_idx = out_idx;
  p.out_X = out_X; p.out_err = out_err; p.out_proj = out_proj;
  p.ws = workspace;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    int n = 0;
    PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lift_fused_kernel, kFusedWarps * 32, 0));
    blocks_per_sm = n > 0 ? n : 1;
  }
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  long long blocks = (long long)sm * blocks_per_sm;
  const long long need = ((long long)B * V * J + kFusedWarps - 1) / kFusedWarps;
  if (blocks > need) blocks = need;
  lift_fused_kernel<<<(unsigned)blocks, kFusedWarps * 32, 0, (cudaStream_t)stream>>>(p);
  PB_LAUNCH_CHECK("lift_fused_kernel");
  return PB200_OK;
}
