// rpsm.cu -- K4: recursive pictorial structure model (RPSM) 3D grid search, batched.
//
// Reference: lib/multiviews/pictorial.py:19-250 (rpsm, compute_unary_term, infer,
// recursive_infer) and run/test/generate_pairwise_constraints.py:60-95 (level-0
// limb-length predicate).
//
// One persistent thread block per frame slot; a block walks frames slot, slot+G, ...
//   level 0 : shared n0^3 grid (4096 bins).  unary_j[b] = sum over views of a bilinear
//             heatmap sample at the projected bin (float64, views in order);
//             max-product up the tree with the level-0 pairwise bit matrix:
//             E_p[i] = unary_p[i] * prod_c max_j ( P_pc[i,j] ? E_c[j] : 0 ), first
//             maximum kept as back pointer; root argmax; back-tracking.
//   level 1..D : per-joint nR^3 grids centred on the current estimate, cell size
//             divided by nR each level; same max-product with the predicate
//             | |g_p[i]-g_c[j]| - L_pc | <= tolerance evaluated on the fly.
// Not HBM bound: per frame the heatmaps are read once (V*J*H*W*4 bytes) while level 0
// does E * n0^6 masked compare steps out of shared memory.
#include "pb_common.cuh"

namespace pb200 {

constexpr int kRpsmThreads = 512;
constexpr int kRpsmMaxJ = PB200_RPSM_MAX_JOINTS;
constexpr int kRpsmMaxBinsR = 64;  // per-joint bins of a refinement level (nR <= 4)

struct RpsmParams {
  const float* hm;
  int B, V, J, H, W;
  const double* campack;
  const int32_t* cam_index;
  const double* box_affine;
  double img_w, img_h;
  const double* root;
  const double* limb;
  const int32_t* edges;  // [E,2] (parent, child), reference iteration order
  const int32_t* order;  // [J] children before parents
  int root_idx;
  const uint32_t* pair_bits;
  int n0, nR, depth;
  double grid_size, tol;
  double* energy_ws;   // [slots][J][nb0]
  uint16_t* bp_ws;     // [slots][E][nb0]
  double* out_pose;
  int32_t* out_trace;
};

struct RpsmShared {
  Cam cam[PB200_MAX_VIEWS];
  double aff[PB200_MAX_VIEWS][6];
  double pose[kRpsmMaxJ][3];
  double limb[kRpsmMaxJ];
  double eR[kRpsmMaxJ][kRpsmMaxBinsR];
  uint8_t bpR[kRpsmMaxJ][kRpsmMaxBinsR];
  int edge_p[kRpsmMaxJ], edge_c[kRpsmMaxJ], order[kRpsmMaxJ], bin[kRpsmMaxJ];
  double red_val[kRpsmThreads / 32];
  int red_idx[kRpsmThreads / 32];
  int root_bin;
};

// unary of joint j at world point X: views accumulated in order from 0.0
__device__ __forceinline__ double unary_at(const RpsmParams& p, const RpsmShared& s, int f, int j,
                                           const double X[3]) {
  const int HW = p.H * p.W;
  double u = 0.0;
  for (int v = 0; v < p.V; ++v) {
    double hx, hy;
    grid_to_heatmap(s.cam[v], s.aff[v], X, p.W, p.H, p.img_w, p.img_h, hx, hy);
    const float* m = p.hm + (((size_t)f * p.V + v) * p.J + j) * HW;
    const int W = p.W;
    u = u + bilinear_zero_outside([m, W](int y, int x) { return __ldg(m + y * W + x); }, p.W, p.H, hx, hy);
  }
  return u;
}

__device__ __forceinline__ void bin_to_point(double size, int n, int b, const double c[3], double X[3]) {
  // np.meshgrid 'xy' indexing flattened C-order: b <-> (iy = b/n^2, ix = (b/n)%n, iz = b%n)
  X[0] = grid_coord(size, n, (b / n) % n, c[0]);
  X[1] = grid_coord(size, n, b / (n * n), c[1]);
  X[2] = grid_coord(size, n, b % n, c[2]);
}

// warp-wide first-max: (value descending, index ascending)
__device__ __forceinline__ void warp_first_max(double& v, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, i, off);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

__global__ void __launch_bounds__(kRpsmThreads, 2) rpsm_kernel(const RpsmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RpsmShared& s = *reinterpret_cast<RpsmShared*>(smem_raw);
  double* Ec = reinterpret_cast<double*>(smem_raw + ((sizeof(RpsmShared) + 15) / 16) * 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = kRpsmThreads / 32;
  const int J = p.J, E = J - 1, V = p.V;
  const int n0 = p.n0, nb0 = n0 * n0 * n0, words0 = (nb0 + 31) / 32;
  const int nR = p.nR, nbR = nR * nR * nR;
  double* energy = p.energy_ws + (size_t)blockIdx.x * J * nb0;
  uint16_t* bp = p.bp_ws + (size_t)blockIdx.x * E * nb0;

  if (tid < E) { s.edge_p[tid] = p.edges[2 * tid]; s.edge_c[tid] = p.edges[2 * tid + 1]; }
  if (tid < J) s.order[tid] = p.order[tid];

  for (int f = blockIdx.x; f < p.B; f += gridDim.x) {
    __syncthreads();
    if (tid < V) {
      load_cam(p.campack + (size_t)p.cam_index[(size_t)f * V + tid] * PB200_CAM_STRIDE, s.cam[tid]);
      for (int k = 0; k < 6; ++k) s.aff[tid][k] = p.box_affine[((size_t)f * V + tid) * 6 + k];
    }
    if (tid < E) s.limb[tid] = p.limb[(size_t)f * E + tid];
    __syncthreads();
    const double centre[3] = {p.root[3 * (size_t)f], p.root[3 * (size_t)f + 1], p.root[3 * (size_t)f + 2]};

    // ---- level 0: unary on the shared grid --------------------------------------
    for (int b = tid; b < nb0; b += kRpsmThreads) {
      double X[3];
      bin_to_point(p.grid_size, n0, b, centre, X);
      for (int j = 0; j < J; ++j) energy[(size_t)j * nb0 + b] = unary_at(p, s, f, j, X);
    }
    __syncthreads();

    // ---- level 0: max-product, leaves -> root -------------------------------------
    for (int oi = 0; oi < J; ++oi) {
      const int par = s.order[oi];
      for (int e = 0; e < E; ++e) {
        if (s.edge_p[e] != par) continue;
        const double* ec = energy + (size_t)s.edge_c[e] * nb0;
        for (int b = tid; b < nb0; b += kRpsmThreads) Ec[b] = ec[b];
        __syncthreads();
        for (int i = warp; i < nb0; i += nwarps) {
          const uint32_t* row = p.pair_bits + ((size_t)e * nb0 + i) * words0;
          double best = -INFINITY;
          int bidx = 0x7fffffff;
          for (int w0 = 0; w0 < words0; w0 += 32) {
            const uint32_t mine = (w0 + lane < words0) ? __ldg(row + w0 + lane) : 0u;
            const int wend = min(32, words0 - w0);
            for (int w = 0; w < wend; ++w) {
              const uint32_t m = __shfl_sync(0xffffffffu, mine, w);
              const int jj = (w0 + w) * 32 + lane;
              if (jj < nb0) {
                const double val = ((m >> lane) & 1u) ? Ec[jj] : 0.0;
                if (bidx == 0x7fffffff || val > best) { best = val; bidx = jj; }
              }
            }
          }
          warp_first_max(best, bidx);
          if (lane == 0) {
            energy[(size_t)par * nb0 + i] = energy[(size_t)par * nb0 + i] * best;
            bp[(size_t)e * nb0 + i] = (uint16_t)bidx;
          }
        }
        __syncthreads();
      }
    }

    // ---- level 0: root argmax (first maximum) and back-tracking -------------------
    {
      const double* er = energy + (size_t)p.root_idx * nb0;
      double best = -INFINITY;
      int bidx = 0x7fffffff;
      for (int b = tid; b < nb0; b += kRpsmThreads) {
        const double v = er[b];
        if (bidx == 0x7fffffff || v > best) { best = v; bidx = b; }
      }
      warp_first_max(best, bidx);
      if (lane == 0) { s.red_val[warp] = best; s.red_idx[warp] = bidx; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < nwarps; ++w)
          if (s.red_val[w] > best || (s.red_val[w] == best && s.red_idx[w] < bidx)) {
            best = s.red_val[w];
            bidx = s.red_idx[w];
          }
        s.bin[p.root_idx] = bidx;
        for (int oi = J - 1; oi >= 0; --oi) {  // parents before children
          const int par = s.order[oi];
          for (int e = 0; e < E; ++e)
            if (s.edge_p[e] == par) s.bin[s.edge_c[e]] = bp[(size_t)e * nb0 + s.bin[par]];
        }
      }
      __syncthreads();
      if (tid < J) {
        double X[3];
        bin_to_point(p.grid_size, n0, s.bin[tid], centre, X);
        s.pose[tid][0] = X[0]; s.pose[tid][1] = X[1]; s.pose[tid][2] = X[2];
        if (p.out_trace) p.out_trace[((size_t)f * (p.depth + 1)) * J + tid] = s.bin[tid];
      }
      __syncthreads();
    }

    // ---- refinement levels ----------------------------------------------------------
    double cur = p.grid_size / (double)n0;
    for (int lvl = 1; lvl <= p.depth; ++lvl) {
      for (int t = tid; t < J * nbR; t += kRpsmThreads) {
        const int j = t / nbR, b = t - j * nbR;
        double X[3];
        bin_to_point(cur, nR, b, s.pose[j], X);
        s.eR[j][b] = unary_at(p, s, f, j, X);
      }
      __syncthreads();
      if (warp == 0) {
        for (int oi = 0; oi < J; ++oi) {
          const int par = s.order[oi];
          for (int i = lane; i < nbR; i += 32) {
            double Xp[3];
            bin_to_point(cur, nR, i, s.pose[par], Xp);
            double acc = s.eR[par][i];
            for (int e = 0; e < E; ++e) {
              if (s.edge_p[e] != par) continue;
              const int c = s.edge_c[e];
              double best = 0.0;
              int bidx = -1;
              for (int jj = 0; jj < nbR; ++jj) {
                double Xc[3];
                bin_to_point(cur, nR, jj, s.pose[c], Xc);
                const double dx = Xp[0] - Xc[0], dy = Xp[1] - Xc[1], dz = Xp[2] - Xc[2];
                const double d = sqrt((dx * dx + dy * dy) + dz * dz);
                const double val = (fabs(d - s.limb[e]) <= p.tol) ? s.eR[c][jj] : 0.0;
                if (bidx < 0 || val > best) { best = val; bidx = jj; }
              }
              acc = acc * best;
              s.bpR[e][i] = (uint8_t)bidx;
            }
            s.eR[par][i] = acc;
          }
          __syncwarp();
        }
        if (lane == 0) {
          double best = s.eR[p.root_idx][0];
          int bidx = 0;
          for (int b = 1; b < nbR; ++b)
            if (s.eR[p.root_idx][b] > best) { best = s.eR[p.root_idx][b]; bidx = b; }
          s.bin[p.root_idx] = bidx;
          for (int oi = J - 1; oi >= 0; --oi) {
            const int par = s.order[oi];
            for (int e = 0; e < E; ++e)
              if (s.edge_p[e] == par) s.bin[s.edge_c[e]] = s.bpR[e][s.bin[par]];
          }
        }
      }
      __syncthreads();
      if (tid < J) {
        double X[3];
        bin_to_point(cur, nR, s.bin[tid], s.pose[tid], X);
        if (p.out_trace) p.out_trace[((size_t)f * (p.depth + 1) + lvl) * J + tid] = s.bin[tid];
        s.pose[tid][0] = X[0]; s.pose[tid][1] = X[1]; s.pose[tid][2] = X[2];
      }
      __syncthreads();
      cur = cur / (double)nR;
    }
    if (tid < J) {
      double* o = p.out_pose + ((size_t)f * J + tid) * 3;
      o[0] = s.pose[tid][0]; o[1] = s.pose[tid][1]; o[2] = s.pose[tid][2];
    }
  }
}

// P[e][i][j] = | |g_i - g_j| - L_e | < 0.4 L_e on the zero-centred n^3 grid
__global__ void pairwise_level0_kernel(const double* __restrict__ avg_limb, int E, int n,
                                       double box_size, uint32_t* __restrict__ bits) {
  const int nb = n * n * n, words = (nb + 31) / 32;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)E * nb * words) return;
  const int w = (int)(t % words);
  const int i = (int)((t / words) % nb);
  const int e = (int)(t / ((long long)words * nb));
  const double zero[3] = {0.0, 0.0, 0.0};
  double Xi[3];
  bin_to_point(box_size, n, i, zero, Xi);
  const double L = avg_limb[e];
  uint32_t m = 0u;
  for (int k = 0; k < 32; ++k) {
    const int j = w * 32 + k;
    if (j >= nb) break;
    double Xj[3];
    bin_to_point(box_size, n, j, zero, Xj);
    const double dx = Xi[0] - Xj[0], dy = Xi[1] - Xj[1], dz = Xi[2] - Xj[2];
    const double d = sqrt((dx * dx + dy * dy) + dz * dz);
    if (fabs(d - L) < 0.4 * L) m |= 1u << k;
  }
  bits[t] = m;
}

static size_t rpsm_smem_bytes(int nb0) {
  return ((sizeof(RpsmShared) + 15) / 16) * 16 + (size_t)nb0 * sizeof(double);
}

static int rpsm_slots(int B, int n_sm) {
  const int cap = n_sm * 2;
  return B < cap ? (B > 0 ? B : 1) : cap;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_rpsm_workspace_bytes(int B, int J, int first_nbins, int n_sm) {
  if (B <= 0 || J < 2 || first_nbins < 1 || n_sm < 1) return 0;
  const size_t nb0 = (size_t)first_nbins * first_nbins * first_nbins;
  const size_t slots = (size_t)rpsm_slots(B, n_sm);
  const size_t e_bytes = ((slots * J * nb0 * sizeof(double) + 255) / 256) * 256;
  return e_bytes + slots * (size_t)(J - 1) * nb0 * sizeof(uint16_t);
}

extern "C" int pb200_rpsm(const float* hm, int B, int V, int J, int H, int W, const double* campack,
                          const int32_t* cam_index, const double* box_affine, int img_w, int img_h,
                          const double* root, const double* limb, const int32_t* edges,
                          const int32_t* order, int root_idx, const uint32_t* pair_bits,
                          int first_nbins, int recur_nbins, int recur_depth, double grid_size,
                          double tolerance, void* workspace, size_t workspace_bytes,
                          double* out_pose, int32_t* out_trace, void* stream) {
  PB_REQUIRE(hm && campack && cam_index && box_affine && root && limb && edges && order && pair_bits,
             "null input pointer");
  PB_REQUIRE(out_pose && workspace, "null output / workspace pointer");
  PB_REQUIRE(B >= 0 && H >= 2 && W >= 2, "bad shape B=%d H=%d W=%d", B, H, W);
  PB_REQUIRE(V >= 1 && V <= PB200_MAX_VIEWS, "V=%d outside [1,%d]", V, PB200_MAX_VIEWS);
  PB_REQUIRE(J >= 2 && J <= kRpsmMaxJ, "J=%d outside [2,%d]", J, kRpsmMaxJ);
  PB_REQUIRE(root_idx >= 0 && root_idx < J, "root_idx out of range");
  PB_REQUIRE(first_nbins >= 1 && first_nbins <= 40, "first_nbins=%d outside [1,40] (uint16 back pointers)", first_nbins);
  PB_REQUIRE(recur_nbins >= 1 && recur_nbins * recur_nbins * recur_nbins <= kRpsmMaxBinsR,
             "recur_nbins^3 must be <= %d", kRpsmMaxBinsR);
  PB_REQUIRE(recur_depth >= 0, "recur_depth < 0");
  if (B == 0) return PB200_OK;
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  const int nb0 = first_nbins * first_nbins * first_nbins;
  PB_REQUIRE(workspace_bytes >= pb200_rpsm_workspace_bytes(B, J, first_nbins, sm),
             "workspace too small: %zu < %zu", workspace_bytes, pb200_rpsm_workspace_bytes(B, J, first_nbins, sm));
  const size_t smem = rpsm_smem_bytes(nb0);
  PB_REQUIRE(smem <= 227 * 1024, "first_nbins=%d needs %zu bytes of shared memory", first_nbins, smem);
  PB_CUDA(cudaFuncSetAttribute(rpsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int slots = rpsm_slots(B, sm);
  RpsmParams p;
  p.hm = hm; p.B = B; p.V = V; p.J = J; p.H = H; p.W = W;
  p.campack = campack; p.cam_index = cam_index; p.box_affine = box_affine;
  p.img_w = (double)img_w; p.img_h = (double)img_h;
  p.root = root; p.limb = limb; p.edges = edges; p.order = order; p.root_idx = root_idx;
  p.pair_bits = pair_bits; p.n0 = first_nbins; p.nR = recur_nbins; p.depth = recur_depth;
  p.grid_size = grid_size; p.tol = tolerance;
  p.energy_ws = reinterpret_cast<double*>(workspace);
  const size_t e_bytes = (((size_t)slots * J * nb0 * sizeof(double) + 255) / 256) * 256;
  p.bp_ws = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(workspace) + e_bytes);
  p.out_pose = out_pose; p.out_trace = out_trace;
  rpsm_kernel<<<slots, kRpsmThreads, smem, (cudaStream_t)stream>>>(p);
  PB_LAUNCH_CHECK("rpsm_kernel");
  return PB200_OK;
}

extern "C" int pb200_pairwise_level0(const double* avg_limb, int E, int nbins, double box_size,
                                     uint32_t* pair_bits, void* stream) {
  PB_REQUIRE(avg_limb && pair_bits, "null pointer");
  PB_REQUIRE(E >= 1 && nbins >= 1 && nbins <= 40, "bad E=%d nbins=%d", E, nbins);
  const long long nb = (long long)nbins * nbins * nbins, words = (nb + 31) / 32;
  const long long n = (long long)E * nb * words;
  pairwise_level0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      avg_limb, E, nbins, box_size, pair_bits);
  PB_LAUNCH_CHECK("pairwise_level0_kernel");
  return PB200_OK;
}
