// lift_fused.cu -- decode -> triangulate -> reprojection error in ONE pass over HBM.
//
// BASELINE.json config 2 (B frames x V views x J joints of HxW float32 heatmaps).
// The only large operand is the heatmap tensor; it is streamed once by
// warp-per-map decoders (csrc/decode.cuh).  Lifting a frame needs the V*J decoded
// coordinates of that frame (16*V*J bytes, L2 resident), so instead of a second
// launch the warp that completes the LAST map of a frame lifts that frame in
// place: lanes = joints, float64 DLT + Jacobi per lane (csrc/lift.cuh),
// reprojection error per view.
//
// Scheduling: persistent blocks (a whole number per SM); warps pull map indices
// from a global counter (the next claim is issued before the current map is
// scanned, so its round trip is hidden), so a warp that is busy lifting simply
// pulls fewer maps and the HBM stream never waits for it.
//
// Hand-off without memory fences: a decoder publishes its result as ONE 16-byte
// record {x, y, maxval, tag=1} (a single st.global.v4) and then bumps the frame's
// arrival counter with a relaxed atomic.  The warp that sees the last arrival reads
// the V*J records with volatile loads; a record whose tag is still 0 is simply
// re-read -- its store was issued before the counter was bumped, so it is already
// in flight and the wait does not depend on any other warp being scheduled.  The
// lifter zeroes the tags again, the last block out resets the claim counter: the
// workspace is all-zero after every launch.  (The first version used
// __threadfence() + atomicAdd per map; ncu showed 35 % of all stall cycles on
// MEMBAR, see profiles/.)
//
// Two streaming front ends, same arithmetic:
//   variant 0  LDG.128 (ld.global.nc.L1::no_allocate), 8 loads per lane per iteration
//   variant 1  per-warp ring of 4 KiB cp.async.bulk (TMA) chunks in shared memory with
//              mbarrier completion; the copy engine keeps kStages chunks per warp in
//              flight independently of the warp's registers and of its epilogue.
#include "decode.cuh"
#include "lift.cuh"

namespace pb200 {

#ifndef PB_FUSED_WARPS
#define PB_FUSED_WARPS 8  // x kStages x 4 KiB = 96 KiB of ring per block, 2 blocks per SM
#endif
#ifndef PB_FUSED_MIN_BLOCKS
#define PB_FUSED_MIN_BLOCKS 2
#endif
#ifndef PB_CLAIM_BATCH
#define PB_CLAIM_BATCH 2   // maps taken per atomic on the global claim counter
#endif
#ifndef PB_SKIP_LIFT
#define PB_SKIP_LIFT 0     // timing experiments only: 1 = decode + hand-off, no lifting
#endif
constexpr int kFusedWarps = PB_FUSED_WARPS;

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
  const double* fmat;     // optional epipolar table [S][V][V][9], subject slot per frame, residuals out
  const int32_t* subj;
  double* out_resid;
  int32_t* ws;     // [0] next map, [1] finished blocks, [4 + f] maps decoded of frame f
  float4* records; // [B*V*J] {x, y, maxval, tag}
};

struct RecordXY {
  const float4* rec;  // &records[frame*V*J + j]
  int row_stride;     // J
  __device__ __forceinline__ float4 get(int v) const {
    const float4* p = rec + (size_t)v * row_stride;
    float4 r = __ldcv(p);
    while (__float_as_int(r.w) == 0) r = __ldcv(p);  // store already in flight, see header
    return r;
  }
  __device__ __forceinline__ void operator()(int v, double& x, double& y) const {
    const float4 r = get(v);
    x = (double)r.x;
    y = (double)r.y;
  }
};

__device__ __noinline__ void lift_frame_joint(const FusedParams& p, int f, int j) {
  const int V = p.V, J = p.J;
  const size_t row0 = (size_t)f * V;
  const int32_t* cam_row = p.cam_index + row0;
  RecordXY xy{p.records + row0 * J + j, J};
  uint32_t mask = 0u;
  for (int v = 0; v < V; ++v) {
    const float4 r = xy.get(v);  // also makes sure every record of this joint has landed
    if (!p.use_conf || r.z > p.conf_thre) mask |= 1u << v;
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
  if (p.out_resid)
    epipolar_joint(p.fmat + (size_t)p.subj[f] * V * V * 9, V, xy,
                   p.out_resid + (size_t)f * V * (V - 1) * J + j, (size_t)J);
  for (int v = 0; v < V; ++v)  // leave the workspace clean for the next launch
    p.records[(row0 + v) * J + j] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// lane 0 publishes one decoded map; returns (to every lane) whether it was the frame's last
__device__ __forceinline__ bool publish(const FusedParams& p, int m, int f, const DecodeOut& o, int lane) {
  int last = 0;
  if (lane == 0) {
    reinterpret_cast<float2*>(p.out_xy)[m] = make_float2(o.x, o.y);
    p.out_maxval[m] = o.maxval;
    if (p.out_idx) p.out_idx[m] = o.idx;
    p.records[m] = make_float4(o.x, o.y, o.maxval, __int_as_float(1));
    last = atomicAdd(p.ws + 4 + f, 1) == p.V * p.J - 1;
  }
  return __shfl_sync(0xffffffffu, last, 0) != 0;
}

__device__ __forceinline__ void lift_if_last(const FusedParams& p, bool last, int f, int lane) {
  if (!last) return;
#if PB_SKIP_LIFT
  for (int jj = lane; jj < p.J; jj += 32)
    for (int v = 0; v < p.V; ++v) p.records[((size_t)f * p.V + v) * p.J + jj] = make_float4(0.f, 0.f, 0.f, 0.f);
#else
  for (int jj = lane; jj < p.J; jj += 32) lift_frame_joint(p, f, jj);
#endif
  if (lane == 0) p.ws[4 + f] = 0;
}

__device__ __forceinline__ void block_exit(const FusedParams& p) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(p.ws + 1, 1) == (int)gridDim.x - 1) {  // last block out resets the counters
      p.ws[0] = 0;
      p.ws[1] = 0;
    }
  }
}

__device__ __forceinline__ int claim(const FusedParams& p, int lane) {
  int m = 0;
  if (lane == 0) m = atomicAdd(p.ws, 1);
  return m;  // valid in lane 0 only until broadcast
}

// ---- variant 0: LDG front end ----------------------------------------------------------
__global__ void __launch_bounds__(kFusedWarps * 32, 4) lift_fused_kernel(const FusedParams p) {
  const int lane = threadIdx.x & 31;
  const int J = p.J, V = p.V, HW = p.H * p.W;
  const int total = p.B * V * J;
  int m = __shfl_sync(0xffffffffu, claim(p, lane), 0);
  while (m < total) {
    const int next_raw = claim(p, lane);  // in flight while this map is scanned
    const int row = m / J, j = m - row * J;
    const float* base = map_base(p.hv, row, j, J, HW);
    const DecodeOut o = decode_map(base, p.H, p.W, p.vec_ok != 0, p.affine + 6 * (size_t)row,
                                   p.post_process != 0, lane);
    const int f = row / V;
    lift_if_last(p, publish(p, m, f, o, lane), f, lane);
    m = __shfl_sync(0xffffffffu, next_raw, 0);
  }
  block_exit(p);
}

// ---- variant 1: TMA bulk-copy ring front end (csrc/decode.cuh::stream_maps_tma) -------------
__global__ void __launch_bounds__(kFusedWarps * 32, PB_FUSED_MIN_BLOCKS) lift_fused_tma_kernel(const FusedParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int J = p.J, V = p.V;
  const int total = p.B * V * J;
  int batch_next = 0, batch_left = 0;  // warp-uniform: maps left of the batch claimed last
  stream_maps_tma(
      smem_raw, kFusedWarps, p.hv, J, p.H, p.W, total, p.affine, p.post_process != 0,
      [&]() {
        if (batch_left == 0) {
          int m0 = 0;
          if (lane == 0) m0 = atomicAdd(p.ws, PB_CLAIM_BATCH);
          batch_next = __shfl_sync(0xffffffffu, m0, 0);
          batch_left = PB_CLAIM_BATCH;
        }
        --batch_left;
        const int m = batch_next++;
        return m < total ? m : total;
      },
      [&](int m, const DecodeOut& o) {  // publish: stores + the arrival atomic are issued, not awaited
        int old = 0;
        if (lane == 0) {
          reinterpret_cast<float2*>(p.out_xy)[m] = make_float2(o.x, o.y);
          p.out_maxval[m] = o.maxval;
          if (p.out_idx) p.out_idx[m] = o.idx;
          p.records[m] = make_float4(o.x, o.y, o.maxval, __int_as_float(1));
          old = atomicAdd(p.ws + 4 + (m / J) / V, 1);
        }
        return old;
      },
      [&](int m, int old) {             // complete: the warp that saw the last arrival lifts the frame
        const bool last = __shfl_sync(0xffffffffu, old, 0) == V * J - 1;
        lift_if_last(p, last, (m / J) / V, lane);
      });
  block_exit(p);
}

int fill_views(const float* const* hm_views_host, int n_ptr, int N, HmViews& hv);
bool views_vec_ok(const HmViews& hv, int HW);
int launch_decode(const HmViews& hv, int N, int J, int H, int W, const double* affine, int post_process,
                  float* out_xy, float* out_maxval, int32_t* out_idx, void* stream);
int launch_lift_after_decode(const double* campack, const int32_t* cam_index, const float* xy,
                             const float* maxval, int use_conf, float conf_thre, int B, int V, int J,
                             int no_dist, double* out_X, float* out_err32, double* out_proj,
                             const double* fmat, const int32_t* subj, double* out_resid, void* stream);

// 2 = two kernels (decode, then one thread per (frame, joint)); measured faster than either fused
// kernel on B200 because lifting inside the streaming warps steals their time (profiles/)
static int g_lift_variant = 2;

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_set_tuning(int key, int value) {
  PB_REQUIRE(key == PB200_TUNE_LIFT_VARIANT, "unknown tuning key %d", key);
  PB_REQUIRE(value >= 0 && value <= 2, "lift variant must be 0 (fused, LDG), 1 (fused, TMA ring) or 2 (two kernels)");
  g_lift_variant = value;
  return PB200_OK;
}

extern "C" size_t pb200_lift_workspace_bytes(int B, int V, int J) {
  if (B < 0 || V < 0 || J < 0) return 0;
  const size_t ints = (((size_t)B + 4 + 3) / 4) * 4;  // keeps the records 16-byte aligned
  return ints * 4 + (size_t)B * V * J * 16;
}

extern "C" int pb200_lift_fused(const float* const* hm_views_host, int n_ptr, int B, int V, int J,
                                int H, int W, const double* affine, int post_process,
                                const double* campack, const int32_t* cam_index, int no_distortion,
                                int use_conf, float conf_thre, float* out_xy, float* out_maxval,
                                int32_t* out_idx, double* out_X, float* out_err, double* out_proj,
                                const double* fmat, const int32_t* subj_index, double* out_resid,
                                void* workspace, void* stream) {
  PB_REQUIRE(B >= 0 && J >= 1 && H >= 1 && W >= 1, "bad shape B=%d J=%d H=%d W=%d", B, J, H, W);
  if (B == 0) return PB200_OK;
  PB_REQUIRE(V >= 2 && V <= PB200_MAX_VIEWS, "V=%d outside [2,%d]", V, PB200_MAX_VIEWS);
  PB_REQUIRE((long long)H * W < (1LL << 24), "map of %dx%d exceeds the float32-exact index range", H, W);
  PB_REQUIRE((long long)B * V * J < (1LL << 31) - (1 << 20), "B*V*J too large for one launch; split the batch");
  PB_REQUIRE(affine && campack && cam_index, "null input pointer");
  PB_REQUIRE(out_xy && out_maxval && out_X && out_err && workspace, "null output / workspace pointer");
  PB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
  PB_REQUIRE(n_ptr == 1 || n_ptr == V, "n_ptr must be 1 or V");
  PB_REQUIRE(out_resid == nullptr || (fmat && subj_index), "out_resid needs fmat and subj_index");
  FusedParams p;
  int rc = fill_views(hm_views_host, n_ptr, B * V, p.hv);
  if (rc != PB200_OK) return rc;
  p.B = B; p.V = V; p.J = J; p.H = H; p.W = W;
  p.vec_ok = views_vec_ok(p.hv, H * W) ? 1 : 0;
  p.affine = affine; p.post_process = post_process;
  p.campack = campack; p.cam_index = cam_index; p.no_dist = no_distortion;
  p.use_conf = use_conf; p.conf_thre = conf_thre;
  p.out_xy = out_xy; p.out_maxval = out_maxval; p.out_idx = out_idx;
  p.out_X = out_X; p.out_err = out_err; p.out_proj = out_proj;
  p.fmat = fmat; p.subj = subj_index; p.out_resid = out_resid;
  p.ws = reinterpret_cast<int32_t*>(workspace);
  const size_t ints = (((size_t)B + 4 + 3) / 4) * 4;
  p.records = reinterpret_cast<float4*>(p.ws + ints);
  if (g_lift_variant == 2) {
    rc = launch_decode(p.hv, B * V, J, H, W, affine, post_process, out_xy, out_maxval, out_idx, stream);
    if (rc != PB200_OK) return rc;
    return launch_lift_after_decode(campack, cam_index, out_xy, out_maxval, use_conf, conf_thre, B, V, J,
                                    no_distortion, out_X, out_err, out_proj, fmat, subj_index, out_resid,
                                    stream);
  }
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  const long long need = ((long long)B * V * J + kFusedWarps - 1) / kFusedWarps;
  const bool tma = g_lift_variant == 1 && p.vec_ok;  // bulk copies need 16-byte aligned maps
  if (tma) {
    const size_t smem = tma_ring_smem_bytes(kFusedWarps);
    static int blocks_per_sm_tma = 0;
    if (blocks_per_sm_tma == 0) {
      PB_CUDA(cudaFuncSetAttribute(lift_fused_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int n = 0;
      PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lift_fused_tma_kernel, kFusedWarps * 32, smem));
      blocks_per_sm_tma = n > 0 ? n : 1;
    }
    long long blocks = (long long)sm * blocks_per_sm_tma;
    if (blocks > need) blocks = need;
    lift_fused_tma_kernel<<<(unsigned)blocks, kFusedWarps * 32, smem, (cudaStream_t)stream>>>(p);
    PB_LAUNCH_CHECK("lift_fused_tma_kernel");
  } else {
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
      int n = 0;
      PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lift_fused_kernel, kFusedWarps * 32, 0));
      blocks_per_sm = n > 0 ? n : 1;
    }
    long long blocks = (long long)sm * blocks_per_sm;
    if (blocks > need) blocks = need;
    lift_fused_kernel<<<(unsigned)blocks, kFusedWarps * 32, 0, (cudaStream_t)stream>>>(p);
    PB_LAUNCH_CHECK("lift_fused_kernel");
  }
  return PB200_OK;
}
