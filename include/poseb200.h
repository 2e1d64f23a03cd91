/*
 * poseb200.h -- C ABI of libposeb200.so: the multiview 2D->3D lifting hot path of
 * LouisNUST/pose-unsupervised as hand-written sm_100a CUDA.
 *
 * The reference has no FFI layer: its boundary is a set of module-level Python
 * functions.  Each entry point below names the reference
 * function it replaces; pose_unsupervised_b200/ binds them with ctypes and keeps
 * the reference's Python signatures (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no allocation, no host synchronisation, no host<->device copy inside: work is
 *     enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default);
 *   - rows are view-minor as in the reference: row = frame * V + view
 *     (lib/multiviews/triangulate.py:83,93; lib/core/function.py:639);
 *   - return value: PB200_OK or a negative PB200_ERR_*; pb200_last_error() gives
 *     the text (thread-local);
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     returns PB200_ERR_CUDA.
 */
#ifndef POSEB200_H_
#define POSEB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_VERSION 100            /* 0.1.0 */

#define PB200_OK 0
#define PB200_ERR_ARG (-1)           /* bad argument (null, size, unsupported combination) */
#define PB200_ERR_CUDA (-2)          /* CUDA runtime error, see pb200_last_error() */
#define PB200_ERR_UNSUPPORTED (-3)   /* shape outside the compiled limits */

#define PB200_MAX_VIEWS 8            /* BASELINE.json sweeps 2/4/8 views */
#define PB200_MAX_JOINTS 64
#define PB200_CAM_STRIDE 24          /* doubles per packed camera */
#define PB200_RPSM_MAX_JOINTS 32

/* Packed camera (PB200_CAM_STRIDE doubles), from the dict of
 * lib/multiviews/cameras.py:12-22:
 *   [0..8] R row-major   [9..11] T (camera centre, world)   [12] fx [13] fy
 *   [14] cx [15] cy      [16..18] k (radial)   [19..20] p (tangential)   [21..23] 0 */

/* dtype tags for small per-row side inputs */
#define PB200_F32 0
#define PB200_F64 1

int pb200_version(void);
const char* pb200_last_error(void);
/* PB200_OK iff a CUDA device of compute capability 10.x is current. */
int pb200_device_check(void);
/* Number of SMs of the current device (grid sizing of the persistent kernels). */
int pb200_sm_count(void);

/* ---- crop affine ----------------------------------------------------------------
 * Replaces utils.transforms.get_affine_transform(center, scale, rot, output_size, shift,
 * inv) (lib/utils/transforms.py:76-109), one 2x3 float64 matrix per row, following
 * numpy's dtype promotion of the float32 point triples and cv2.getAffineTransform's
 * elimination order bit for bit.
 *   center [n,2], scale [n,2] : dtype given by center_dtype / scale_dtype
 *   rot_sincos [n,2] float64  : (np.sin, np.cos) of np.pi * rot / 180 per row, evaluated by
 *                               the caller on the host; NULL = no rotation (the lifting path)
 *   shift_x, shift_y          : the `shift` argument (default 0, 0) and its dtype tag
 *   out    [n,6] float64 row-major 2x3
 */
int pb200_crop_affine(const void* center, int center_dtype, const void* scale, int scale_dtype,
                      const double* rot_sincos, double shift_x, double shift_y, int shift_dtype,
                      int n, int out_w, int out_h, int inv, double* out, void* stream);

/* ---- debug-assert build ---------------------------------------------------------------
 * A library compiled with -DPB200_DEBUG_CHECKS=1 (python -m pose_unsupervised_b200.build --debug ->
 * libposeb200_debug.so) makes the kernels check the invariants their hand-offs rely on (map indices and
 * counts of the decode schedule, candidate addresses / back-tracked bins / task indices / stage contents of the
 * on-chip RPSM, item indices of RANSAC).  pb200_debug_enabled() is 1 for such a build;
 * pb200_debug_violations synchronises the device and copies the 16 violation counters (all zero in a
 * release build), optionally resetting them.  Stands in for compute-sanitizer, which the GPU pool does
 * not offer. */
int pb200_debug_enabled(void);
int pb200_debug_violations(int32_t* out16, int reset);
/* debug build: one check that fails on purpose (counter 15 += 1); release build: PB200_ERR_UNSUPPORTED */
int pb200_debug_selftest(void);

/* ---- tuning (process-wide) --------------------------------------------------------
 * PB200_TUNE_DECODE_SCHEDULE: how the persistent decode kernel hands maps to its warps.
 *   PB200_DECODE_STATIC  warp w of block b takes maps b*8+w, +grid*8, ...
 *   PB200_DECODE_DYNAMIC (default) the same strided share for the first 7/8 of an even split, the rest claimed
 *                        in small batches from a counter: the end of the kernel evens out (1 % faster alone)
 *                        and a block that starts late because another kernel (an overlapped NCCL collective)
 *                        still holds its SM decodes fewer maps instead of finishing late.  The library keeps
 *                        a 64 KiB pool of claim counters per device for this (allocated on first use; a launch
 *                        that finds none -- e.g. the first ever call happens inside a stream capture -- uses
 *                        the static form).  Results are identical.
 */
#define PB200_TUNE_DECODE_SCHEDULE 2
#define PB200_DECODE_STATIC 0
#define PB200_DECODE_DYNAMIC 1
int pb200_set_tuning(int key, int value);

/* ---- heatmap decode ---------------------------------------------------------------
 * Replaces core.inference.get_max_preds (lib/core/inference.py:19-47) when
 * affine == NULL and core.inference.get_final_preds (:50-75) otherwise.
 *   hm_views_host : host array of n_ptr device pointers.  n_ptr == 1: one tensor
 *                   [N,J,H,W] float32.  n_ptr == V: V tensors [N/V,J,H,W] as
 *                   validate() holds them (lib/core/function.py:560,632); output rows
 *                   are then interleaved view-minor.
 *   affine        : [N,6] float64 from pb200_crop_affine(inv=1, out=(W,H)), or NULL
 *   post_process  : config.TEST.POST_PROCESS (quarter-pixel shift), only with affine
 *   out_xy [N,J,2] float32, out_maxval [N,J] float32, out_idx [N,J] int32 (flat
 *   argmax, first maximum, NaN counts as maximum; may be NULL)
 */
int pb200_decode(const float* const* hm_views_host, int n_ptr, int N, int J, int H, int W,
                 const double* affine, int post_process,
                 float* out_xy, float* out_maxval, int32_t* out_idx, void* stream);

/* Flip-test averaging fused with the decode: replaces lib/core/function.py:567-583 (flip_back_th,
 * TEST.SHIFT_HEATMAP, (view + view_flipped) * 0.5) followed by get_final_preds (:632-640).
 *   hm_views_host / hm_flip_views_host : the plain and the mirrored-input network outputs
 *               (n_ptr tensors each, as in pb200_decode)
 *   joint_src [J] int32 (device) : joint whose flipped map feeds joint j (j itself, or its
 *               left/right partner from dataset.flip_pairs)
 *   out_avg   [N,J,H,W] float32 : the averaged heatmaps, view-minor rows (what validate() stores)
 */
int pb200_decode_flip(const float* const* hm_views_host, const float* const* hm_flip_views_host,
                      int n_ptr, int N, int J, int H, int W, const int32_t* joint_src, int shift_heatmap,
                      const double* affine, int post_process,
                      float* out_avg, float* out_xy, float* out_maxval, int32_t* out_idx, void* stream);

/* Replaces utils.transforms.transform_preds (lib/utils/transforms.py:67-73) on already
 * decoded heatmap coordinates: out[n,j] = [x, y, 1] @ affine[n].T in float64.
 *   coords [N,J,2] (coords_dtype), affine [N,6] float64 -> out [N,J,2] float64
 */
int pb200_transform_preds(const void* coords, int coords_dtype, const double* affine, int N, int J,
                          double* out, void* stream);

/* ---- camera projection ------------------------------------------------------------
 * model 0 replaces multiviews.cameras.project_pose (lib/multiviews/cameras.py:25-54,
 * averaged focal length, H36M tangential form); model 1 is pymvg find2d
 * (distorted plumb-bob, lib/multiviews/triangulate.py:147,210); model 2 is model 1
 * without distortion; model 3 is model 0 with separate fx, fy (project_point_radial with
 * f = [fx, fy], cameras.py:17-18,52).  pts [n,3] float64 world -> out [n,2] float64 pixels.
 */
int pb200_project(const double* campack, int cam_id, const double* pts, int n, int model,
                  double* out, void* stream);

/* world_to_camera_frame (to_world = 0:  R (x - T)) and camera_to_world_frame
 * (to_world = 1:  R^T x + T), lib/multiviews/cameras.py:57-82.
 * R [9], T [3], pts [n,3] -> out [n,3], all float64.
 */
int pb200_frame_change(const double* R, const double* T, const double* pts, int n, int to_world,
                       double* out, void* stream);

/* ---- triangulation / reprojection / RANSAC ---------------------------------------
 * Common inputs:
 *   campack   [ncam, PB200_CAM_STRIDE] float64
 *   cam_index [B*V] int32  : packed-camera id of every row
 *   xy        [B*V, J, 2]  : float32 or float64 (xy_dtype)
 *   vis       [B*V, J] uint8 (non-zero = visible) or NULL (all visible)
 *
 * pb200_triangulate replaces multiviews.triangulate.triangulate_poses
 * (lib/multiviews/triangulate.py:57-99): out_X [B,J,3] float64, zeros where fewer
 * than two views are visible.
 * pb200_reproject replaces reproject_poses (:169-213): out_proj [B*V,J,2] float64,
 * out_vis [B*V,J] uint8; optional out_X [B,J,3], out_err [B*V,J] float64 (pixel
 * distance between out_proj and xy).
 * pb200_ransac replaces ransac (:102-166): out_vis [B*V,J] uint8.
 */
int pb200_triangulate(const double* campack, const int32_t* cam_index, const void* xy, int xy_dtype,
                      const uint8_t* vis, int B, int V, int J, int no_distortion,
                      double* out_X, void* stream);
int pb200_reproject(const double* campack, const int32_t* cam_index, const void* xy, int xy_dtype,
                    const uint8_t* vis, int B, int V, int J, int no_distortion,
                    double* out_proj, uint8_t* out_vis, double* out_X, double* out_err,
                    void* stream);
int pb200_ransac(const double* campack, const int32_t* cam_index, const void* xy, int xy_dtype,
                 const uint8_t* vis, int B, int V, int J, int no_distortion,
                 double reproj_thre, int num_inliers, uint8_t* out_vis, void* stream);

/* ---- epipolar residual ------------------------------------------------------------
 * Replaces the body of FundamentalLoss.__call__ (lib/core/loss.py:101-133) and
 * run/test/test_fund_mtx.py:56-69.
 *   fmat      [S, V, V, 9] float64 : F for (subject, a, b), row-major 3x3
 *   subj_index[B] int32
 *   weight    [B*V, J] (w_dtype) or NULL
 *   out_resid [B, V*(V-1), J] float64 : |[x_b,1] F [x_a,1]| (times w_b*w_a),
 *             pairs in itertools.permutations(range(V), 2) order
 *   out_sum   [1] float64 or NULL : += sum of out_resid (must be zeroed by the caller;
 *             one atomic add per thread block)
 */
int pb200_epipolar(const double* fmat, const int32_t* subj_index, const void* xy, int xy_dtype,
                   const void* weight, int w_dtype, int B, int V, int J,
                   double* out_resid, double* out_sum, void* stream);

/* Exact fundamental matrices from calibrated cameras, F = K_b^-T [t]x R K_a^-1 (unit Frobenius
 * norm), so that x_b^T F x_a = 0 for pin-hole projections: replaces the offline LMedS estimate
 * of run/test/generate_fundamental_matirx.py:45-57.  cam_a / cam_b [n] int32 ids into campack
 * -> out_F [n,9] float64.
 */
int pb200_fundamental(const double* campack, const int32_t* cam_a, const int32_t* cam_b, int n,
                      double* out_F, void* stream);

/* break_limb_length of run/pose3d/estimate.py:84-96: out_flag[f] = 1 iff some limb of pose f
 * deviates from its expected length by more than thres * expected.  poses [B,J,3] float64,
 * edges [E,2] int32, limb [B,E] (limb_per_frame = 1) or [E] float64 -> out_flag [B] uint8.
 */
int pb200_limb_break(const double* poses, const int32_t* edges, const double* limb, int limb_per_frame,
                     int B, int J, int E, double thres, uint8_t* out_flag, void* stream);

/* ---- differentiable epipolar term (training-time consumer, lib/core/function.py:298-310) -------
 * pb200_softargmax_fwd replaces generate_integral_preds_2d_th (lib/utils/transforms.py:149-171):
 *   p = softmax(beta * hm) per map (beta = 100), out_xy [N,J,2] = (sum p*col, sum p*row) float32,
 *   out_stats [N,J,2] = (max logit, sum of exp) kept for the backward pass.
 * pb200_softargmax_bwd: grad_hm [N,J,H,W] = beta * p * ((col - x) * gx + (row - y) * gy).
 * pb200_epipolar_grad: gradient of sum |x_b^T F x_a| * w_b * w_a * gscale (FundamentalLoss,
 *   lib/core/loss.py:101-133) with respect to xy, ACCUMULATED into grad_xy [B*V,J,2] float64
 *   (zero it first).
 */
int pb200_softargmax_fwd(const float* hm, int N, int J, int H, int W, float beta, float* out_xy,
                         float* out_stats, void* stream);
int pb200_softargmax_bwd(const float* hm, const float* stats, const float* xy, const float* grad_xy,
                         int N, int J, int H, int W, float beta, float* grad_hm, void* stream);
int pb200_epipolar_grad(const double* fmat, const int32_t* subj_index, const void* xy, int xy_dtype,
                        const void* weight, int w_dtype, int B, int V, int J, double gscale,
                        double* grad_xy, void* stream);

/* ---- MPJPE partial sums -----------------------------------------------------------
 * run/test/test_triangulate.py:98-101: norm = |pred - gt| over [B,J];
 * out[0]+=sum, out[1]+=sum of squares, out[2]=max(out[2],.), out[3]+=count  (float64;
 * zero out4 before the first call).
 * This is the payload of the multi-GPU all-reduce.
 */
int pb200_mpjpe_stats(const double* pred, const double* gt, int B, int J, double* out4,
                      void* stream);

/* ---- lift: decode -> triangulate -> reprojection error (+ epipolar residuals) --------
 * The headline path of BASELINE.json (config 2): pb200_decode (get_final_preds) followed
 * by pb200_reproject on the decoded coordinates with joints_vis = maxval > conf_thre, issued
 * back to back on `stream`.  The heatmaps are read from HBM exactly once; the lift reads the
 * decoded coordinates back from L2.
 *   conf_thre : a joint is visible in a view iff maxval > conf_thre
 *               (run/test/test_pseudo_label.py:194); pass use_conf = 0 for "all visible"
 *   outputs   : out_xy/out_maxval/out_idx as pb200_decode; out_X [B,J,3] float64;
 *               out_err [B*V,J] float32 reprojection error in pixels (zeros where the joint
 *               was not lifted); out_proj [B*V,J,2] float64 or NULL.
 *   epipolar  : with fmat [S,V,V,9] and subj_index [B] (as pb200_epipolar) and out_resid
 *               [B, V(V-1), J] float64 non-NULL, the algebraic epipolar residuals of the decoded
 *               coordinates are written by the same lift threads (all three NULL to skip).
 */
int pb200_lift_fused(const float* const* hm_views_host, int n_ptr, int B, int V, int J, int H, int W,
                     const double* affine, int post_process,
                     const double* campack, const int32_t* cam_index, int no_distortion,
                     int use_conf, float conf_thre,
                     float* out_xy, float* out_maxval, int32_t* out_idx,
                     double* out_X, float* out_err, double* out_proj,
                     const double* fmat, const int32_t* subj_index, double* out_resid,
                     void* stream);

/* The lifting half of pb200_lift_fused on its own, for callers that put work between the two
 * launches: xy [B*V,J,2] / maxval [B*V,J] float32 as written by pb200_decode; outputs as above. */
int pb200_lift_decoded(const double* campack, const int32_t* cam_index, const float* xy,
                       const float* maxval, int use_conf, float conf_thre, int B, int V, int J,
                       int no_distortion, double* out_X, float* out_err, double* out_proj,
                       const double* fmat, const int32_t* subj_index, double* out_resid, void* stream);

/* ---- RPSM: recursive pictorial structure grid search -------------------------------
 * Replaces multiviews.pictorial.rpsm (lib/multiviews/pictorial.py:214-250) for a
 * batch of frames, one thread block per frame.
 *   hm        [B, V, J, H, W] float32
 *   cam_index [B*V] int32 ; box_affine [B*V, 6] float64 from
 *             pb200_crop_affine(inv=0, out=IMAGE_SIZE)
 *   root      [B,3] float64 grid centre; limb [B,E] float64 per-frame limb lengths,
 *             E = J-1 edges in edge order (see tree below)
 *   tree      : edges [E,2] int32 = (parent, child) in the reference's iteration order
 *               (for node in skeleton: for child in node['children'],
 *               lib/multiviews/pictorial.py:124-128); order [J] int32 = processing order,
 *               children before parents (skeleton_sorted_by_level,
 *               lib/multiviews/body.py:39-57); root_idx = body.root_idx
 *   pair_bits [E, nbins0, nbins0/32] uint32 : the level-0 `pairwise_constraint`
 *             (lib/multiviews/pictorial.py:240) as a bit matrix, bit (j%32) of word
 *             [e][i][j/32] set iff P_e[i,j] != 0 (nbins0 = first_nbins^3)
 *   use_lut   : 1 if pb200_pairwise_lut_check found the matrix to depend on the index offset
 *             (|dy|,|dx|,|dz|) only -- true for matrices generated on the regular grid; the
 *             kernel then keeps row 0 of each edge in shared memory instead of reading rows
 *   max_reach : out_flag[1] of pb200_pairwise_lut_check (largest |bin offset| of an allowed pair),
 *             or -1 if unknown.  With use_lut = 1, max_reach in [0, 5] and first_nbins <= 16 level 0
 *             runs entirely on chip (heatmaps staged in shared memory by cp.async.bulk, energies in
 *             shared memory; each joint's final energy vector is also kept in the workspace for the
 *             back-tracking); any other combination takes the generic kernel.  Both give identical
 *             results.
 *   workspace : bytes from pb200_rpsm_workspace_bytes, 256-byte aligned
 *   out_pose  [B,J,3] float64 ; out_trace [B, depth+1, J] int32 chosen bin per level (or NULL)
 */
size_t pb200_rpsm_workspace_bytes(int B, int J, int first_nbins, int n_sm);
int pb200_rpsm(const float* hm, int B, int V, int J, int H, int W,
               const double* campack, const int32_t* cam_index, const double* box_affine,
               int img_w, int img_h, const double* root, const double* limb,
               const int32_t* edges, const int32_t* order, int root_idx,
               const uint32_t* pair_bits, int use_lut, int max_reach,
               int first_nbins, int recur_nbins, int recur_depth, double grid_size, double tolerance,
               void* workspace, size_t workspace_bytes,
               double* out_pose, int32_t* out_trace, void* stream);

/* Level-0 pairwise bit matrix from average limb lengths: replaces the offline
 * O(n^6) Python generator run/test/generate_pairwise_constraints.py:60-95
 * (P[i,j] = | |g_i - g_j| - L | < 0.4 L on the zero-centred nbins^3 grid).
 *   avg_limb [E] float64 -> pair_bits [E, n^3, n^3/32] uint32
 */
int pb200_pairwise_level0(const double* avg_limb, int E, int nbins, double box_size,
                          uint32_t* pair_bits, void* stream);
/* out_flag: device int32[2], zeroed by the caller.  out_flag[0] |= 1 unless every P_e[i,j] equals
 * P_e[0, (|dy|*n+|dx|)*n+|dz|], i.e. unless the bit matrix is a function of the offset only;
 * out_flag[1] = largest max(|dy|,|dx|,|dz|) over the allowed pairs of row 0 of any edge. */
int pb200_pairwise_lut_check(const uint32_t* pair_bits, int E, int nbins, int32_t* out_flag,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POSEB200_H_ */
