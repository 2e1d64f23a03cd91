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
//
// The level-0 maximisation is the reference's hot spot (73 % of its time: it densifies
// a 4096x4096 matrix per edge).  Here it is done EXACTLY but without touching most of
// the matrix: the child bins are sorted once per edge by (energy descending, index
// ascending) in shared memory (bitonic sort), and every parent bin walks that list
// until it meets its first allowed child -- which is by construction the first maximum
// np.argmax would return.  The "0 * E" entries of the reference's product (disallowed
// children contribute 0, which wins when every allowed energy is <= 0) are handled
// explicitly.  The pairwise predicate is read either from the bit matrix row of the
// parent, or -- when the matrix is a function of |index offset| only, which
// pb200_pairwise_lut_check verifies -- from row 0 of the edge kept in shared memory
// (512 bytes per edge for 16^3 bins), so the walk never leaves the SM.
//
// Not HBM bound: per frame the heatmaps are read once (V*J*H*W*4 bytes, ~1.1 MB);
// the work is shared-memory sorting / probing and float64 sampling.
#include "pb_common.cuh"

namespace pb200 {

constexpr int kRpsmThreads = 512;
constexpr int kRpsmMaxJ = PB200_RPSM_MAX_JOINTS;
constexpr int kRpsmMaxBinsR = 64;     // per-joint bins of a refinement level (nR <= 4)
constexpr int kRpsmMaxBins0 = 16384;  // level-0 bins (n0 <= 25)
#ifndef PB_RPSM_ENUM_REACH
#define PB_RPSM_ENUM_REACH 5
#endif
constexpr int kRpsmEnumReach = PB_RPSM_ENUM_REACH;  // shells within +-reach bins are enumerated, larger ones walked
constexpr int kRpsmEnumDensity = 160; // same decision for arbitrary bit matrices: allowed children per row

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
  int use_lut;
  int n0, nR, depth;
  int npad;              // n0^3 rounded up to a power of two (bitonic sort length)
  double grid_size, tol;
  double* energy_ws;     // [slots][J][nb0]
  uint16_t* bp_ws;       // [slots][E][nb0]
  double* out_pose;
  int32_t* out_trace;
};

struct RpsmShared {
  Cam cam[PB200_MAX_VIEWS];
  double aff[PB200_MAX_VIEWS][6];
  double pose[kRpsmMaxJ][3];
  double limb[kRpsmMaxJ];
  int edge_p[kRpsmMaxJ], edge_c[kRpsmMaxJ], order[kRpsmMaxJ], bin[kRpsmMaxJ];
  double red_val[32];
  int red_idx[32];
  int reach;
};

__device__ __forceinline__ double sample_view(const RpsmParams& p, const RpsmShared& s, int f, int v,
                                              int j, const double* X) {
  double hx, hy;
  grid_to_heatmap(s.cam[v], s.aff[v], X, p.W, p.H, p.img_w, p.img_h, hx, hy);
  const float* m = p.hm + (((size_t)f * p.V + v) * p.J + j) * (size_t)(p.H * p.W);
  const int W = p.W;
  return bilinear_zero_outside([m](int t) { return __ldg(m + t); }, p.W, p.H, hx, hy);
}

// unary of joint j at world point X: views accumulated in order from 0.0
__device__ __forceinline__ double unary_at(const RpsmParams& p, const RpsmShared& s, int f, int j,
                                           const double X[3]) {
  double u = 0.0;
  for (int v = 0; v < p.V; ++v) u = u + sample_view(p, s, f, v, j, X);
  return u;
}

__device__ __forceinline__ void bin_coords(int n, int b, int& iy, int& ix, int& iz) {
  // np.meshgrid 'xy' indexing flattened C-order: b <-> (iy = b/n^2, ix = (b/n)%n, iz = b%n)
  iz = b % n;
  const int q = b / n;
  ix = q % n;
  iy = q / n;
}

__device__ __forceinline__ void bin_to_point(double size, int n, int b, const double c[3], double X[3]) {
  int iy, ix, iz;
  bin_coords(n, b, iy, ix, iz);
  X[0] = grid_coord(size, n, ix, c[0]);
  X[1] = grid_coord(size, n, iy, c[1]);
  X[2] = grid_coord(size, n, iz, c[2]);
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

// Sort ord[0..npad) so that position 0 holds the bin np.argmax would pick first:
// energy descending, bin index ascending; pad ids (>= nb0) go last.
__device__ __forceinline__ void sort_children(const double* __restrict__ Ec, uint16_t* ord, int nb0,
                                              int npad) {
  const int tid = threadIdx.x;
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (npad >> 1); t += kRpsmThreads) {
        const int i = 2 * t - (t & (j - 1));
        const int l = i + j;
        const int a = ord[i], b = ord[l];
        const double ea = a < nb0 ? Ec[a] : -INFINITY, eb = b < nb0 ? Ec[b] : -INFINITY;
        const bool a_pad = a >= nb0, b_pad = b >= nb0;
        bool a_first;
        if (a_pad != b_pad) a_first = b_pad;
        else a_first = (ea > eb) || (ea == eb && a < b);
        const bool up = (i & k) == 0;
        if (up != a_first) { ord[i] = (uint16_t)b; ord[l] = (uint16_t)a; }
      }
      __syncthreads();
    }
  }
}

// ---- refinement levels (pictorial.py:193-211, 243-248) -----------------------------------------
// per-joint nR^3 grids centred on the current estimate; one thread per (view, joint, bin) sample,
// ordered sum over views, limb predicate on the fly, tree max-product by warp 0.  s.pose holds
// the level-0 estimate on entry and the final pose on exit.  Block-wide (T threads).
template <int T>
__device__ __forceinline__ void refine_levels(const RpsmParams& p, RpsmShared& s, int f, double* gp,
                                              double* eR, double* sv, uint8_t* bpR) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int J = p.J, E = J - 1, V = p.V;
  const int n0 = p.n0;
  const int nR = p.nR, nbR = nR * nR * nR;
  double cur = p.grid_size / (double)n0;
  for (int lvl = 1; lvl <= p.depth; ++lvl) {
    for (int t = tid; t < J * nbR; t += T) {
      const int j = t / nbR, b = t - j * nbR;
      double X[3];
      bin_to_point(cur, nR, b, s.pose[j], X);
      gp[3 * t] = X[0]; gp[3 * t + 1] = X[1]; gp[3 * t + 2] = X[2];
    }
    __syncthreads();
    // one thread per (view, joint, bin) sample, then the ordered sum over views
    for (int t = tid; t < V * J * nbR; t += T) {
      const int v = t / (J * nbR), r = t - v * (J * nbR);
      sv[t] = sample_view(p, s, f, v, r / nbR, gp + 3 * r);
    }
    __syncthreads();
    for (int t = tid; t < J * nbR; t += T) {
      double u = 0.0;
      for (int v = 0; v < V; ++v) u = u + sv[v * (J * nbR) + t];
      eR[t] = u;
    }
    __syncthreads();
    if (warp == 0) {
      for (int oi = 0; oi < J; ++oi) {
        const int par = s.order[oi];
        if (nbR == 8) {
          // lanes = (parent bin i = lane/4) x (child bins 2q, 2q+1 with q = lane%4)
          const int i = lane >> 2, q = lane & 3;
          const double* gi = gp + 3 * (par * 8 + i);
          double acc = eR[par * 8 + i];
          for (int e = 0; e < E; ++e) {
            if (s.edge_p[e] != par) continue;
            const int c = s.edge_c[e];
            double best = 0.0;
            int bidx = -1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int jj = 2 * q + h;
              const double* gj = gp + 3 * (c * 8 + jj);
              const double dx = gi[0] - gj[0], dy = gi[1] - gj[1], dz = gi[2] - gj[2];
              const double d = sqrt((dx * dx + dy * dy) + dz * dz);
              const double val = (fabs(d - s.limb[e]) <= p.tol) ? eR[c * 8 + jj] : 0.0;
              if (bidx < 0 || val > best) { best = val; bidx = jj; }
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {  // merge the 4 lanes of this parent bin
              const double ov = __shfl_xor_sync(0xffffffffu, best, o);
              const int ob = __shfl_xor_sync(0xffffffffu, bidx, o);
              if (ov > best || (ov == best && ob < bidx)) { best = ov; bidx = ob; }
            }
            acc = acc * best;
            if (q == 0) bpR[e * 8 + i] = (uint8_t)bidx;
          }
          __syncwarp();
          if (q == 0) eR[par * 8 + i] = acc;
        } else {
          for (int i = lane; i < nbR; i += 32) {
            const double* gi = gp + 3 * (par * nbR + i);
            double acc = eR[par * nbR + i];
            for (int e = 0; e < E; ++e) {
              if (s.edge_p[e] != par) continue;
              const int c = s.edge_c[e];
              double best = 0.0;
              int bidx = -1;
              for (int jj = 0; jj < nbR; ++jj) {
                const double* gj = gp + 3 * (c * nbR + jj);
                const double dx = gi[0] - gj[0], dy = gi[1] - gj[1], dz = gi[2] - gj[2];
                const double d = sqrt((dx * dx + dy * dy) + dz * dz);
                const double val = (fabs(d - s.limb[e]) <= p.tol) ? eR[c * nbR + jj] : 0.0;
                if (bidx < 0 || val > best) { best = val; bidx = jj; }
              }
              acc = acc * best;
              bpR[e * nbR + i] = (uint8_t)bidx;
            }
            eR[par * nbR + i] = acc;
          }
        }
        __syncwarp();
      }
      if (lane == 0) {
        const double* er = eR + p.root_idx * nbR;
        double best = er[0];
        int bidx = 0;
        for (int b = 1; b < nbR; ++b)
          if (er[b] > best) { best = er[b]; bidx = b; }
        s.bin[p.root_idx] = bidx;
        for (int oi = J - 1; oi >= 0; --oi) {
          const int par = s.order[oi];
          for (int e = 0; e < E; ++e)
            if (s.edge_p[e] == par) s.bin[s.edge_c[e]] = bpR[e * nbR + s.bin[par]];
        }
      }
    }
    __syncthreads();
    if (tid < J) {
      const int b = s.bin[tid];
      if (p.out_trace) p.out_trace[((size_t)f * (p.depth + 1) + lvl) * J + tid] = b;
      const double* g = gp + 3 * (tid * nbR + b);
      const double X0 = g[0], X1 = g[1], X2 = g[2];
      s.pose[tid][0] = X0; s.pose[tid][1] = X1; s.pose[tid][2] = X2;
    }
    __syncthreads();
    cur = cur / (double)nR;
  }
}

__global__ void __launch_bounds__(kRpsmThreads, 2) rpsm_kernel(const RpsmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RpsmShared& s = *reinterpret_cast<RpsmShared*>(smem_raw);
  const int n0 = p.n0, nb0 = n0 * n0 * n0, words0 = (nb0 + 31) / 32, npad = p.npad;
  size_t off = ((sizeof(RpsmShared) + 15) / 16) * 16;
  double* Ec = reinterpret_cast<double*>(smem_raw + off);
  off += (size_t)nb0 * sizeof(double);
  uint32_t* coord = reinterpret_cast<uint32_t*>(smem_raw + off);  // packed (iy, ix, iz) per bin
  off += (size_t)nb0 * sizeof(uint32_t);
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + off);    // row 0 of the edge's bit matrix
  off += (size_t)words0 * sizeof(uint32_t);
  uint32_t* zmask = reinterpret_cast<uint32_t*>(smem_raw + off);  // [|dy|][|dx|][iz] -> allowed jz bits
  off += (n0 <= 32 ? (size_t)nb0 : 0) * sizeof(uint32_t);
  uint16_t* ord = reinterpret_cast<uint16_t*>(smem_raw + off);
  off += (size_t)npad * sizeof(uint16_t);
  off = ((off + 15) / 16) * 16;
  // refinement-level arrays, sized by the actual J, V and nR^3
  const int nbR_ = p.nR * p.nR * p.nR;
  double* gp = reinterpret_cast<double*>(smem_raw + off);   // [J][nbR][3] grid points
  off += (size_t)p.J * nbR_ * 3 * sizeof(double);
  double* eR = reinterpret_cast<double*>(smem_raw + off);   // [J][nbR] energies
  off += (size_t)p.J * nbR_ * sizeof(double);
  double* sv = reinterpret_cast<double*>(smem_raw + off);   // [V][J*nbR] per-view samples
  off += (size_t)p.V * p.J * nbR_ * sizeof(double);
  uint8_t* bpR = smem_raw + off;                            // [E][nbR] back pointers

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = kRpsmThreads / 32;
  const int J = p.J, E = J - 1, V = p.V;
  const int nR = p.nR, nbR = nR * nR * nR;
  double* energy = p.energy_ws + (size_t)blockIdx.x * J * nb0;
  uint16_t* bp = p.bp_ws + (size_t)blockIdx.x * E * nb0;

  if (tid < E) { s.edge_p[tid] = p.edges[2 * tid]; s.edge_c[tid] = p.edges[2 * tid + 1]; }
  if (tid < J) s.order[tid] = p.order[tid];
  for (int b = tid; b < nb0; b += kRpsmThreads) {
    int iy, ix, iz;
    bin_coords(n0, b, iy, ix, iz);
    coord[b] = (uint32_t)iy | ((uint32_t)ix << 8) | ((uint32_t)iz << 16);
  }

  for (int f = blockIdx.x; f < p.B; f += gridDim.x) {
    __syncthreads();
    if (tid < V) {
      load_cam(p.campack + (size_t)p.cam_index[(size_t)f * V + tid] * PB200_CAM_STRIDE, s.cam[tid]);
      for (int k = 0; k < 6; ++k) s.aff[tid][k] = p.box_affine[((size_t)f * V + tid) * 6 + k];
    }
    if (tid < E) s.limb[tid] = p.limb[(size_t)f * E + tid];
    __syncthreads();
    const double centre[3] = {p.root[3 * (size_t)f], p.root[3 * (size_t)f + 1], p.root[3 * (size_t)f + 2]};

    // ---- level 0: unary on the shared grid (projection once per (bin, view)) ----------
    for (int b = tid; b < nb0; b += kRpsmThreads) {
      double X[3];
      bin_to_point(p.grid_size, n0, b, centre, X);
      double hx[PB200_MAX_VIEWS], hy[PB200_MAX_VIEWS];
      for (int v = 0; v < V; ++v)
        grid_to_heatmap(s.cam[v], s.aff[v], X, p.W, p.H, p.img_w, p.img_h, hx[v], hy[v]);
      const int W = p.W;
      const size_t HW = (size_t)p.H * p.W;
      for (int j = 0; j < J; ++j) {
        double u = 0.0;
        for (int v = 0; v < V; ++v) {
          const float* m = p.hm + (((size_t)f * V + v) * J + j) * HW;
          u = u + bilinear_zero_outside([m](int t) { return __ldg(m + t); }, p.W, p.H,
                                        hx[v], hy[v]);
        }
        energy[(size_t)j * nb0 + b] = u;
      }
    }
    __syncthreads();

    // ---- level 0: max-product, leaves -> root ---------------------------------------
    for (int oi = 0; oi < J; ++oi) {
      const int par = s.order[oi];
      for (int e = 0; e < E; ++e) {
        if (s.edge_p[e] != par) continue;
        const double* ec = energy + (size_t)s.edge_c[e] * nb0;
        for (int b = tid; b < nb0; b += kRpsmThreads) Ec[b] = ec[b];
        for (int b = tid; b < npad; b += kRpsmThreads) ord[b] = (uint16_t)b;
        const uint32_t* bits_e = p.pair_bits + (size_t)e * nb0 * words0;
        for (int w = tid; w < words0; w += kRpsmThreads) lut[w] = bits_e[w];  // row of bin 0
        if (tid == 0) s.reach = 0;
        __syncthreads();
        // reach = largest |index offset| with an allowed pair.  Short limbs have tiny shells
        // (6 bins for a 130 mm limb on the 133 mm grid): walking a sorted list would take
        // ~4096/|shell| steps per parent, enumerating the (2*reach+1)^3 neighbourhood is exact
        // and far cheaper.  Large shells go through the sorted walk.
        bool enumerate = false;
        if (p.use_lut) {
          int r = 0;
          for (int d = tid; d < nb0; d += kRpsmThreads)
            if ((lut[d >> 5] >> (d & 31)) & 1u) {
              const uint32_t cd = coord[d];
              r = max(r, max((int)(cd & 255), max((int)((cd >> 8) & 255), (int)(cd >> 16))));
            }
          if (r > 0) atomicMax(&s.reach, r);
          __syncthreads();
          enumerate = s.reach <= kRpsmEnumReach;
          if (enumerate && n0 <= 32) {
            // allowed child z-bins as a bit mask per (|dy|, |dx|, parent z): the enumeration below
            // then visits allowed children only
            for (int t = tid; t < nb0; t += kRpsmThreads) {
              const uint32_t ct = coord[t];
              const int ady = ct & 255, adx = (ct >> 8) & 255, pz = ct >> 16;
              uint32_t m = 0u;
              for (int jz = 0; jz < n0; ++jz) {
                const int d = (ady * n0 + adx) * n0 + abs(pz - jz);
                m |= ((lut[d >> 5] >> (d & 31)) & 1u) << jz;
              }
              zmask[t] = m;
            }
            __syncthreads();
          }
        } else {
          // arbitrary bit matrix: estimate the row density from every 64th row; sparse rows are
          // enumerated bit by bit, dense ones walked in sorted order
          int cnt = 0;
          const int nrows = (nb0 + 63) / 64;
          for (int t = tid; t < nrows * words0; t += kRpsmThreads)
            cnt += __popc(__ldg(bits_e + (size_t)((t / words0) * 64) * words0 + (t % words0)));
          if (cnt > 0) atomicAdd(&s.reach, cnt);
          __syncthreads();
          enumerate = s.reach <= kRpsmEnumDensity * nrows;
        }
        if (!enumerate) sort_children(Ec, ord, nb0, npad);
        const int reach = s.reach;
        for (int i = tid; i < nb0; i += kRpsmThreads) {
          const uint32_t ci = coord[i];
          const int iy = ci & 255, ix = (ci >> 8) & 255, iz = ci >> 16;
          const uint32_t* row = bits_e + (size_t)i * words0;
          auto allowed = [&](int j) -> bool {
            if (p.use_lut) {
              const uint32_t cj = coord[j];
              const int dy = abs(iy - (int)(cj & 255)), dx = abs(ix - (int)((cj >> 8) & 255)),
                        dz = abs(iz - (int)(cj >> 16));
              const int d = (dy * n0 + dx) * n0 + dz;
              return (lut[d >> 5] >> (d & 31)) & 1u;
            }
            return (__ldg(row + (j >> 5)) >> (j & 31)) & 1u;
          };
          // `found` = first maximum over the allowed children (np.argmax order)
          int found = -1;
          if (enumerate && !p.use_lut) {
            double best = 0.0;
            for (int w = 0; w < words0; ++w) {
              uint32_t m = __ldg(row + w);
              while (m) {
                const int j = w * 32 + (__ffs(m) - 1);   // ascending
                m &= m - 1;
                if (j < nb0) {
                  const double v = Ec[j];
                  if (found < 0 || v > best) { best = v; found = j; }
                }
              }
            }
          } else if (enumerate && n0 <= 32) {
            double best = 0.0;
            for (int jy = max(iy - reach, 0); jy <= min(iy + reach, n0 - 1); ++jy)
              for (int jx = max(ix - reach, 0); jx <= min(ix + reach, n0 - 1); ++jx) {
                uint32_t m = zmask[(abs(iy - jy) * n0 + abs(ix - jx)) * n0 + iz];
                const int base = (jy * n0 + jx) * n0;
                while (m) {
                  const int j = base + (__ffs(m) - 1);   // ascending in this loop order
                  m &= m - 1;
                  const double v = Ec[j];
                  if (found < 0 || v > best) { best = v; found = j; }
                }
              }
          } else if (enumerate) {
            double best = 0.0;
            for (int jy = max(iy - reach, 0); jy <= min(iy + reach, n0 - 1); ++jy)
              for (int jx = max(ix - reach, 0); jx <= min(ix + reach, n0 - 1); ++jx)
                for (int jz = max(iz - reach, 0); jz <= min(iz + reach, n0 - 1); ++jz) {
                  const int d = (abs(iy - jy) * n0 + abs(ix - jx)) * n0 + abs(iz - jz);
                  if ((lut[d >> 5] >> (d & 31)) & 1u) {
                    const int j = (jy * n0 + jx) * n0 + jz;   // ascending in this loop order
                    const double v = Ec[j];
                    if (found < 0 || v > best) { best = v; found = j; }
                  }
                }
          } else {
            for (int k = 0; k < nb0; ++k) {  // (energy desc, index asc) order
              const int j = ord[k];
              if (allowed(j)) { found = j; break; }
            }
          }
          double val;
          int arg;
          const double mA = found >= 0 ? Ec[found] : 0.0;
          if (found >= 0 && mA > 0.0) {
            val = mA;
            arg = found;
          } else {
            // the reference multiplies by the 0/1 matrix: disallowed children are candidates
            // with value 0 at their own index
            int first_dis = -1;
            for (int j = 0; j < nb0; ++j)
              if (!allowed(j)) { first_dis = j; break; }
            if (found < 0) { val = 0.0; arg = 0; }                       // nothing allowed: all zeros
            else if (mA == 0.0) { val = 0.0; arg = (first_dis >= 0 && first_dis < found) ? first_dis : found; }
            else if (first_dis >= 0) { val = 0.0; arg = first_dis; }     // every allowed energy < 0
            else { val = mA; arg = found; }
          }
          energy[(size_t)par * nb0 + i] = energy[(size_t)par * nb0 + i] * val;
          bp[(size_t)e * nb0 + i] = (uint16_t)arg;
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

    refine_levels<kRpsmThreads>(p, s, f, gp, eR, sv, bpR);
    if (tid < J) {
      double* o = p.out_pose + ((size_t)f * J + tid) * 3;
      o[0] = s.pose[tid][0]; o[1] = s.pose[tid][1]; o[2] = s.pose[tid][2];
    }
  }
}

// =================================================================================================
// On-chip level 0 -- the B200-shaped path (pb200_rpsm picks it for offset-only pairwise tables with
// shells within +-kRpsmEnumReach bins and n0 <= 16, i.e. every table built on the regular grid).
//
// One persistent 1024-thread block per SM; per frame
//   * the joint-independent half of every bilinear sample (top-left tap and the two fractions of each
//     (bin, view)) is computed once and parked in an L2-resident scratch of V*nb0*20 bytes per block;
//   * the tree is walked depth-first by a small program built once per block (thread 0): every
//     joint's unary is sampled exactly when it is first needed, from ITS V heatmaps staged in shared
//     memory by the copy engine (cp.async.bulk + mbarrier; the next joint's maps -- or the next
//     frame's first -- are requested as soon as the stage is free, so the copy runs under the
//     max-product of the current joint);
//   * energy vectors (nb0 float64) live in shared memory: with the first child's message folded into the
//     parent's unary on the fly, the two reference skeletons never need more than 4 live vectors
//     (128 KiB); deeper trees spill the extra ones to scratch;
//   * the FORWARD pass computes only the VALUE max_j(P[i,j] ? S[j] : 0) per parent bin, which makes the
//     maximum order-free: a lane owns two z-neighbouring parents and walks the edge's child-offset lists
//     (children of one parent, of both, of the other), so most energies are read from shared memory once
//     and compared twice (oc_maxprod_unit_flat);
//   * parents whose accumulated energy is exactly 0 skip the maximisation (0 * finite = 0);
//   * every child's final vector is also copied to an L2 scratch (E * nb0 float64 per block), and the
//     BACK-TRACKING resolves the reference's first-argmax rule there, for the ONE parent bin per edge that
//     lies on the chosen pose -- edges of one tree depth in parallel, 256 threads each (oc_pick_*);
//   * the 2^3 refinement levels evaluate all limb predicates and heatmap samples one per thread, and one warp
//     runs the 8-bin max-product by tree depth (refine_levels8); their arrays share the vectors' memory.
// Bit-identical to rpsm_kernel (tests/test_gpu_rpsm.py compares both with each other, with the oracle and
// with the reference's golden frames).
// =================================================================================================
#ifndef PB_RPSM_L2_HINTS
#define PB_RPSM_L2_HINTS 1   // evict_first on the heatmap stream, evict_last on the parked sample taps: same speed,
                             // 0.27x the algorithmic heatmap bytes less DRAM traffic (profiles/r02_rpsm_*)
#endif
constexpr int kOcThreads = 1024;
constexpr int kOcMaxPer = 4;                 // level-0 bins per thread: nb0 <= 4096
constexpr int kOcMaxOps = 2 * kRpsmMaxJ;
constexpr int kOcMaxUnits = kOcMaxPer * kOcThreads / 32;   // warp tasks of 32 consecutive parent bins
constexpr int kOcStageBytes = 64 * 1024;
constexpr int kOcSmemBudget = 227 * 1024;
constexpr int kOcRefineSteps = 12;           // refinement schedule: (depth, four edges) steps / tree depths it can hold

enum { kOpLeaf = 0, kOpFirst = 1, kOpAcc = 2 };
struct OcOp {
  uint8_t kind, joint, edge, src, dst, samp, pad0, pad1;
};

struct OcLayout {   // filled by the host (rpsm_onchip_layout)
  int stage_off, stage_views;   // staged views per group; 0 = sample with __ldg (odd-sized / huge maps)
  int dzm_off, dzm_cap;         // uint16 entries: allowed |dz| per (edge, |dy|, dx)
  int list_off, list_cap;       // 8-byte entries: the edges' sorted child-offset lists
  int refine_off;
  int vec_off, vec_stride;      // vec_stride doubles per vector: nb0 (rounded up to even)
  int nsm, nspill;              // vectors in shared memory / in scratch per block
  size_t smem_bytes;
  double* coords_ws;            // [blocks][V][nb0][2]  bilinear fractions
  int32_t* tap_ws;              // [blocks][V][nb0]     top-left tap of the sample
  double* spill_ws;             // [blocks][nspill][vec_stride]
  double* sfin_ws;              // [blocks][E][vec_stride]  every child's final energy vector (back-tracking)
};

struct OcShared {
  RpsmShared base;
  OcOp ops[kOcMaxOps];
  int nops, nsamp, root_buf, prog_err;
  int reach[kRpsmMaxJ];
  int doff[kRpsmMaxJ + 1];      // offset of edge e in the |dz| table
  uint16_t loff[kRpsmMaxJ][8][2];   // offset / (even) length of the two offset sub-lists of (edge, |oy|): children
  uint16_t lcnt[kRpsmMaxJ][8][2];   // of ONE of a lane's two parents (A, B alternating), children of both
  uint32_t use_flat;            // bit e: edge e's offset lists are in shared memory (n0 = 16 and they fit)
  int child_start[kRpsmMaxJ + 1];
  uint8_t child_edge[kRpsmMaxJ];
  uint8_t samp_joint[kRpsmMaxJ];
  uint8_t unit_order[kOcMaxUnits];
  uint8_t pair_order[64];       // the 64 warp tasks of the fast form (64 parents each), interior first
  uint8_t pmask[kRpsmMaxJ * 8]; // refinement: allowed child bins (bit j) per (edge, parent bin)
  uint8_t bt_edge[kRpsmMaxJ];   // edges sorted by the tree depth of their child joint (root's children: 1) ...
  uint8_t bt_start[kRpsmMaxJ + 2];   // ... and where depth d starts in that list: the back-tracking order
  int max_depth;
  // refinement (refine_levels8): what every lane of warp 0 does in each step of the 8-bin max-product and of its
  // back-tracking, precomputed so that the tables' walk is not a chain of dependent shared-memory reads
  uint2 rstep[kOcRefineSteps][32];     // x = edge | child joint << 8 | #children of the child << 16 | active << 24,
                                       // y = the edges of those children (4 x 8 bits)
  uint32_t rback[kOcRefineSteps][32];  // per tree depth: edge | parent joint << 8 | child joint << 16 | active << 24
  uint32_t rroot;                      // the root's child edges (4 x 8 bits)
  int rsteps, rroot_n;                 // steps in rstep (-1: the tree does not fit the tables); children of the root
  double bt_val[kOcThreads / 32], bt_fval[kOcThreads / 32];   // back-tracking: per-warp partial results
  int bt_idx[kOcThreads / 32], bt_fidx[kOcThreads / 32];
  unsigned long long mbar;
  long long stage_tag;          // (frame, sample slot, group) of the bulk copy issued last; -1 = none
  unsigned stage_seq;           // bulk-copy groups issued so far (mbarrier phase = stage_seq - 1)
  int unit_next;                // next warp task of the running max-product
  int nonfinite;
};

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// parked per-frame data: written once, re-read by every joint -- ask L2 to keep it
__device__ __forceinline__ void st_keep_f64x2(double* ptr, double a, double b, uint64_t pol) {
#if PB_RPSM_L2_HINTS
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(ptr), "d"(a), "d"(b), "l"(pol) : "memory");
#else
  __stcg(reinterpret_cast<double2*>(ptr), make_double2(a, b));
#endif
}
__device__ __forceinline__ void st_keep_s32(int32_t* ptr, int32_t v, uint64_t pol) {
#if PB_RPSM_L2_HINTS
  asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(ptr), "r"(v), "l"(pol) : "memory");
#else
  __stcg(ptr, v);
#endif
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void oc_mbar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void oc_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oc_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "OCWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra OCDONE_%=;\n\t"
      "bra OCWAIT_%=;\n\t"
      "OCDONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void oc_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                            uint64_t pol) {
#if PB_RPSM_L2_HINTS
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
#endif
}

// Offsets of the edges' |dz| tables (thread 0, once the shell reach of every edge is known).
__device__ void oc_table_offsets(OcShared& os, int E) {
  int off = 0;
  for (int e = 0; e < E; ++e) {
    os.doff[e] = off;
    off += (os.reach[e] + 1) * (2 * os.reach[e] + 1);   // rows |oy| = 0..r, columns ox = -r..r
  }
  os.doff[E] = off;
}

// Depth-first program (thread 0, once per block).  LEAF: dst = unary(joint).  FIRST: dst = unary(joint) *
// msg(edge, src).  ACC: dst *= msg(edge, src).  Children in edge-array order, which is the order the
// reference multiplies them in (pictorial.py:44-56).  Buffers are numbered so that the low ids (shared
// memory) are reused first.  Also: child lists per joint (CSR) and the edges sorted by tree depth.
__device__ void oc_build_program(OcShared& os, int J, int E, int root_idx, int nbuf) {
  const RpsmShared& s = os.base;
  struct Frame { int node, pos, acc, pending; };
  Frame st[kRpsmMaxJ];
  uint32_t free_mask = nbuf >= 32 ? 0xffffffffu : ((1u << nbuf) - 1u);
  int sp = 0, nops = 0, nsamp = 0, ret = -1, err = 0;
  st[sp++] = Frame{root_idx, 0, -1, -1};
  auto alloc = [&]() -> int {
    if (free_mask == 0u) { err = 1; return 0; }
    const int b = __ffs(free_mask) - 1;
    free_mask &= ~(1u << b);
    return b;
  };
  while (sp > 0 && nops < kOcMaxOps - 1) {
    Frame& fr = st[sp - 1];
    if (ret >= 0) {   // a child has just returned its buffer
      OcOp op;
      op.joint = (uint8_t)fr.node; op.edge = (uint8_t)fr.pending; op.src = (uint8_t)ret;
      op.samp = 0; op.pad0 = op.pad1 = 0;
      if (fr.acc < 0) {
        fr.acc = alloc();
        op.kind = kOpFirst;
        op.samp = (uint8_t)nsamp;
        os.samp_joint[nsamp++] = (uint8_t)fr.node;
      } else {
        op.kind = kOpAcc;
      }
      op.dst = (uint8_t)fr.acc;
      os.ops[nops++] = op;
      free_mask |= 1u << ret;
      ret = -1;
    }
    int e = fr.pos;
    while (e < E && s.edge_p[e] != fr.node) ++e;
    if (e < E) {
      fr.pos = e + 1;
      fr.pending = e;
      if (sp >= kRpsmMaxJ) { err = 1; break; }
      st[sp++] = Frame{s.edge_c[e], 0, -1, -1};
      continue;
    }
    if (fr.acc < 0) {   // leaf
      OcOp op;
      op.kind = kOpLeaf; op.joint = (uint8_t)fr.node; op.edge = 0; op.src = 0; op.pad0 = op.pad1 = 0;
      op.dst = (uint8_t)alloc();
      op.samp = (uint8_t)nsamp;
      os.samp_joint[nsamp++] = (uint8_t)fr.node;
      os.ops[nops++] = op;
      ret = op.dst;
    } else {
      ret = fr.acc;
    }
    --sp;
  }
  if (sp != 0 || nsamp != J) err = 1;   // not a tree spanning all joints
  os.nops = nops;
  os.nsamp = nsamp;
  os.root_buf = ret < 0 ? 0 : ret;
  os.prog_err = err;
  int ce = 0;
  for (int j = 0; j < J; ++j) {
    os.child_start[j] = ce;
    for (int e = 0; e < E; ++e)
      if (s.edge_p[e] == j) os.child_edge[ce++] = (uint8_t)e;
  }
  os.child_start[J] = ce;
  int md = 0;
  uint8_t depth[kRpsmMaxJ], up[kRpsmMaxJ];
  for (int j = 0; j < J; ++j) up[j] = (uint8_t)root_idx;
  for (int e = 0; e < E; ++e) up[s.edge_c[e]] = (uint8_t)s.edge_p[e];
  for (int e = 0; e < E; ++e) {
    int d = 1, j = s.edge_p[e];
    for (int guard = 0; j != root_idx && guard < E; ++guard) {
      j = up[j];
      ++d;
    }
    depth[e] = (uint8_t)d;
    md = d > md ? d : md;
  }
  os.max_depth = md;
  int pos = 0;
  for (int d = 1; d <= md; ++d) {
    os.bt_start[d] = (uint8_t)pos;
    for (int e = 0; e < E; ++e)
      if (depth[e] == d) os.bt_edge[pos++] = (uint8_t)e;
  }
  os.bt_start[md + 1] = (uint8_t)pos;
}

// Schedule of refine_levels8's warp 0 (built by warp 0 once per block, after oc_build_program): step = (tree
// depth, deepest first) x (four edges of that depth), lanes = (edge slot) x (bin).
__device__ void oc_refine_schedule(OcShared& os, int J, int E, int root_idx) {
  const RpsmShared& s = os.base;
  const int lane = threadIdx.x & 31, sl = lane >> 3;
  int st = 0;
  bool fit = os.max_depth <= kOcRefineSteps;
  for (int d = os.max_depth; d >= 1; --d)
    for (int c0 = os.bt_start[d]; c0 < os.bt_start[d + 1]; c0 += 4, ++st) {
      if (st >= kOcRefineSteps) continue;
      const bool on = c0 + sl < os.bt_start[d + 1];
      const int e = on ? os.bt_edge[c0 + sl] : 0, c = s.edge_c[e];
      const int k0 = os.child_start[c], nch = on ? os.child_start[c + 1] - k0 : 0;
      if (nch > 4) fit = false;
      uint32_t ch = 0u;
      for (int k = 0; k < nch && k < 4; ++k) ch |= (uint32_t)os.child_edge[k0 + k] << (8 * k);
      os.rstep[st][lane] = make_uint2((uint32_t)e | ((uint32_t)c << 8) | ((uint32_t)nch << 16) | ((on ? 1u : 0u) << 24), ch);
    }
  for (int d = 1; d <= os.max_depth && d <= kOcRefineSteps; ++d) {
    const int k = os.bt_start[d] + lane;
    const bool on = k < os.bt_start[d + 1];
    const int e = on ? os.bt_edge[k] : 0;
    os.rback[d - 1][lane] = (uint32_t)e | ((uint32_t)s.edge_p[e] << 8) | ((uint32_t)s.edge_c[e] << 16) | ((on ? 1u : 0u) << 24);
    if (os.bt_start[d + 1] - os.bt_start[d] > 32) fit = false;
  }
  const int r0 = os.child_start[root_idx], rn = os.child_start[root_idx + 1] - r0;
  if (rn > 4 || st > kOcRefineSteps) fit = false;
  fit = __all_sync(0xffffffffu, fit);
  if (lane == 0) {
    uint32_t ch = 0u;
    for (int k = 0; k < rn && k < 4; ++k) ch |= (uint32_t)os.child_edge[r0 + k] << (8 * k);
    os.rroot = ch;
    os.rroot_n = rn;
    os.rsteps = fit ? st : -1;
  }
}

// ---- refinement levels for 2^3 grids (the reference's RECUR_NBINS = 2), block-wide ------------------
// Same arithmetic as refine_levels, arranged for 1024 threads: the limb predicates of all E x 8 x 8 (parent
// bin, child bin) pairs do not depend on the energies, so they are evaluated one per thread next to the
// heatmap samples; what is left for warp 0 is a table-driven max-product over 8-bin vectors.
template <int T>
__device__ __forceinline__ void refine_levels8(const RpsmParams& p, OcShared& os, int f, double* eR, double* sv,
                                               double* msgR, uint8_t* bpR) {
  RpsmShared& s = os.base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int J = p.J, E = J - 1, V = p.V;
  double cur = p.grid_size / (double)p.n0;
  for (int lvl = 1; lvl <= p.depth; ++lvl) {
    // (a 2^3 grid point is three selects and adds: every thread makes the ones it needs, no table, no barrier)
    for (int t = tid; t < V * J * 8; t += T) {
      const int v = t / (J * 8), r = t - v * (J * 8);
      double X[3];
      bin_to_point(cur, 2, r & 7, s.pose[r >> 3], X);
      sv[t] = sample_view(p, s, f, v, r >> 3, X);
    }
    for (int t = tid; t < E * 64; t += T) {   // E * 64 is a multiple of 32: whole warps
      const int e = t >> 6, i = (t >> 3) & 7, jj = t & 7;
      double gi[3], gj[3];
      bin_to_point(cur, 2, i, s.pose[s.edge_p[e]], gi);
      bin_to_point(cur, 2, jj, s.pose[s.edge_c[e]], gj);
      const double dx = gi[0] - gj[0], dy = gi[1] - gj[1], dz = gi[2] - gj[2];
      const double d = sqrt((dx * dx + dy * dy) + dz * dz);
      const unsigned bal = __ballot_sync(0xffffffffu, fabs(d - s.limb[e]) <= p.tol);
      if (jj == 0) os.pmask[e * 8 + i] = (uint8_t)((bal >> (lane & 24)) & 0xffu);
    }
    __syncthreads();
    if (warp == 0) {
      for (int t = lane; t < J * 8; t += 32) {
        double u = 0.0;
        for (int v = 0; v < V; ++v) u = u + sv[v * (J * 8) + t];
        eR[t] = u;
      }
      __syncwarp();
      // Max-product by tree depth, deepest edges first, up to four edges of a depth at a time:
      // lanes = (edge slot) x (bin).  An edge's child vector is final once the messages of the child's own
      // children (one depth down, already computed) are multiplied in, in edge order (pictorial.py:44-56).
      const int i = lane & 7;
      if (os.rsteps >= 0) {   // the schedule tables hold this tree (oc_refine_schedule)
        const int nst = os.rsteps;
        uint2 w = os.rstep[0][lane];
        for (int st = 0; st < nst; ++st) {
          const uint2 cur_w = w;
          if (st + 1 < nst) w = os.rstep[st + 1][lane];
          const bool on = (cur_w.x >> 24) != 0u;
          const int e = cur_w.x & 255u, c = (cur_w.x >> 8) & 255u, nch = (cur_w.x >> 16) & 255u;
          if (on) {
            double acc = eR[c * 8 + i];
            uint32_t ch = cur_w.y;
            for (int k = 0; k < nch; ++k, ch >>= 8) acc = acc * msgR[(ch & 255u) * 8 + i];
            eR[c * 8 + i] = acc;
          }
          __syncwarp();
          if (on) {   // message to parent bin i: the first maximum over the allowed child bins
            const unsigned pm = os.pmask[e * 8 + i];
            double best = 0.0;
            int bidx = -1;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const double val = ((pm >> jj) & 1u) ? eR[c * 8 + jj] : 0.0;
              if (bidx < 0 || val > best) { best = val; bidx = jj; }
            }
            msgR[e * 8 + i] = best;
            bpR[e * 8 + i] = (uint8_t)bidx;
          }
          __syncwarp();
        }
        double acc = 0.0;
        if (lane < 8) {
          acc = eR[p.root_idx * 8 + lane];
          uint32_t ch = os.rroot;
          for (int k = 0; k < os.rroot_n; ++k, ch >>= 8) acc = acc * msgR[(ch & 255u) * 8 + lane];
        }
        double best = acc;   // lane 0: root argmax, first maximum, from the other lanes' registers
        int bidx = 0;
#pragma unroll
        for (int b = 1; b < 8; ++b) {
          const double v = __shfl_sync(0xffffffffu, acc, b);
          if (v > best) { best = v; bidx = b; }
        }
        if (lane == 0) s.bin[p.root_idx] = bidx;
        __syncwarp();
        for (int d = 0; d < os.max_depth; ++d) {   // back-tracking: parents before children, a lane per edge
          const uint32_t wb = os.rback[d][lane];
          if (wb >> 24) s.bin[(wb >> 16) & 255u] = bpR[(wb & 255u) * 8 + s.bin[(wb >> 8) & 255u]];
          __syncwarp();
        }
      } else {
      const int sl = lane >> 3;
      for (int d = os.max_depth; d >= 1; --d)
        for (int c0 = os.bt_start[d]; c0 < os.bt_start[d + 1]; c0 += 4) {
          const bool on = c0 + sl < os.bt_start[d + 1];
          const int e = on ? os.bt_edge[c0 + sl] : 0, c = s.edge_c[e];
          if (on) {
            double acc = eR[c * 8 + i];
            for (int ce = os.child_start[c]; ce < os.child_start[c + 1]; ++ce) acc = acc * msgR[os.child_edge[ce] * 8 + i];
            eR[c * 8 + i] = acc;
          }
          __syncwarp();
          if (on) {   // message to parent bin i: the first maximum over the allowed child bins
            const unsigned pm = os.pmask[e * 8 + i];
            double best = 0.0;
            int bidx = -1;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const double val = ((pm >> jj) & 1u) ? eR[c * 8 + jj] : 0.0;
              if (bidx < 0 || val > best) { best = val; bidx = jj; }
            }
            msgR[e * 8 + i] = best;
            bpR[e * 8 + i] = (uint8_t)bidx;
          }
          __syncwarp();
        }
      if (lane < 8) {
        const int c = p.root_idx;
        double acc = eR[c * 8 + lane];
        for (int ce = os.child_start[c]; ce < os.child_start[c + 1]; ++ce) acc = acc * msgR[os.child_edge[ce] * 8 + lane];
        eR[c * 8 + lane] = acc;
      }
      __syncwarp();
      if (lane == 0) {
        const double* er = eR + p.root_idx * 8;
        double best = er[0];
        int bidx = 0;
        for (int b = 1; b < 8; ++b)
          if (er[b] > best) { best = er[b]; bidx = b; }
        s.bin[p.root_idx] = bidx;
      }
      __syncwarp();
      for (int d = 1; d <= os.max_depth; ++d) {   // back-tracking: parents before children, a lane per edge
        for (int k = os.bt_start[d] + lane; k < os.bt_start[d + 1]; k += 32) {
          const int e = os.bt_edge[k];
          s.bin[s.edge_c[e]] = bpR[e * 8 + s.bin[s.edge_p[e]]];
        }
        __syncwarp();
      }
      }
      if (lane < J) {
        const int b = s.bin[lane];
        if (p.out_trace) p.out_trace[((size_t)f * (p.depth + 1) + lvl) * J + lane] = b;
        double X[3];
        bin_to_point(cur, 2, b, s.pose[lane], X);
        s.pose[lane][0] = X[0]; s.pose[lane][1] = X[1]; s.pose[lane][2] = X[2];
      }
    }
    __syncthreads();
    cur = cur / 2.0;
  }
}

// What the reference's product with the 0/1 matrix does once the first maximum over the ALLOWED children
// is known (`found`, value `best`): disallowed children are candidates with value 0 at their own index.
__device__ __forceinline__ void oc_finish_max(const uint32_t* __restrict__ row0, int n0, int nb0, int iy, int ix,
                                              int iz, int found, double best, double& val, int& arg) {
  const double mA = found >= 0 ? best : 0.0;
  if (found >= 0 && mA > 0.0) {
    val = mA;
    arg = found;
    return;
  }
  int first_dis = -1;   // rare path: every allowed energy is <= 0
  for (int j = 0; j < nb0; ++j) {
    int jy, jx, jz;
    bin_coords(n0, j, jy, jx, jz);
    const int d = (abs(iy - jy) * n0 + abs(ix - jx)) * n0 + abs(iz - jz);
    if (!((__ldg(row0 + (d >> 5)) >> (d & 31)) & 1u)) { first_dis = j; break; }
  }
  if (found < 0) { val = 0.0; arg = 0; }
  else if (mA == 0.0) { val = 0.0; arg = (first_dis >= 0 && first_dis < found) ? first_dis : found; }
  else if (first_dis >= 0) { val = 0.0; arg = first_dis; }
  else { val = mA; arg = found; }
}

// ---- shared-memory accessors by 32-bit address (keeps the hot loops free of generic-pointer arithmetic) ----
__device__ __forceinline__ double oc_lds64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void oc_sts64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// Energy vectors of the on-chip kernel are stored x-FASTEST on the 16^3 grid: memory index m = (iy, iz, ix),
// i.e. the logical bin index (iy, ix, iz) of pictorial.py:108-119 with its two low nibbles swapped.  The unary
// samples arrive with lanes along grid x (fewer shared-memory bank conflicts when reading the staged maps, see
// the sampling loop), so they are stored with unit stride, and the max-product walks along x just as well as
// along z.  Everything visible outside the kernel (traces, poses) uses logical indices.
__device__ __forceinline__ int oc_swap(int j) { return (j & ~0xff) | ((j & 15) << 4) | ((j >> 4) & 15); }

__device__ __forceinline__ uint4 oc_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
// One candidate child from an edge's offset list.  `o2` = (ox + 8) | (k + 8) << 8, `d8` = (k*16 + ox)*8;
// p2 = (ix + 8) | (iz + 8) << 8 of the lane's parent (0 for a lane that sits this task out).  Both child
// coordinates are inside the 16-wide grid iff bits 4-5 of both byte sums read 01.  The forward pass only needs
// the VALUE of the maximum: which child attains it is resolved during back-tracking, for the one parent bin per
// edge that lies on the chosen pose (oc_backtrack).  Lanes whose child is outside do not load at all (a
// redirected load to a cell holding -inf costs a bank conflict per half-warp: measured 133 k frames/s vs 168 k).
// `v` is the caller's scratch for the loaded energy: a lane outside keeps whatever it held (the compare is masked),
// and naming it -- four of them, used in rotation -- keeps the compiler from inventing a loop-carried register
// per candidate for that "old value", which is what spilled.
__device__ __forceinline__ void oc_cand(uint32_t o2, uint32_t d8, uint32_t p2, uint32_t base, double& best, double& v) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q, f;\n\t"
      ".reg .b32 s, d;\n\t"
      "add.u32 s, %2, %4;\n\t"
      "setp.ne.u32 f, 0, 0;\n\t"
      "lop3.or.b32 s|q, s, 0x3030, 0x1010, 0x6A, f;\n\t"   /* q = ((s & 0x3030) ^ 0x1010) != 0: OUTSIDE */
      "add.u32 d, %5, %3;\n\t"
      "@!q ld.shared.f64 %1, [d];\n\t"
      "setp.gt.and.f64 p, %1, %0, !q;\n\t"
      "@p mov.f64 %0, %1;\n\t"
      "}" : "+d"(best), "+d"(v) : "r"(o2), "r"(d8), "r"(p2), "r"(base) : "memory");
}

// The factor a parent's energy is multiplied with, from the maximum over its allowed in-grid children
// (`best`, -inf when it has none): max_j ( P[i,j] ? S[j] : 0 ) of pictorial.py:50-56 with the zeros of the
// disallowed children taken into account.
__device__ __forceinline__ double oc_factor(const uint32_t* __restrict__ row0, int n0, int nb0, int iy, int ix, int iz,
                                            bool any, double best) {
  if (any && best > 0.0) return best;
  double val;
  int arg;
  oc_finish_max(row0, n0, nb0, iy, ix, iz, any ? 0 : -1, best, val, arg);   // (val does not depend on `found`'s value)
  return val;
}

__device__ __forceinline__ void oc_cand_check(uint32_t o2, uint32_t d8, uint32_t p2, uint32_t base, uint32_t sS) {
#if PB200_DEBUG_CHECKS   // an in-grid candidate must lie inside the source vector
  const bool inside = (((o2 + p2) & 0x3030u) ^ 0x1010u) == 0u;
  const uint32_t a = base + d8;
  PB_DCHECK(!inside || (a >= sS && a < sS + 4096u * 8u), kDbgRpsmCandAddr);
#endif
}

// One warp task of the max-product, FAST FORM (n0 = 16, source and destination in shared memory, finite frame):
//   D[i] <- D[i] * max_j ( P[i,j] ? S[j] : 0 )                                  (pictorial.py:50-56)
// for the 64 parents (iy, z in [4q, 4q+4), all ix) of task u = 4 iy + q.  A lane owns TWO parents, A = (iy, z0, ix)
// and B = (iy, z0+1, ix) with z0 = 4q + 2 (lane / 16): a child at z-offset k' from z0 is neighbour k' of A and
// k'-1 of B, and the allowed z-offsets of a row (oy, ox) come in runs, so most children serve both -- one
// shared-memory read and ONE comparison, into a running maximum that both parents take at the end.  The allowed
// children of the edge are LISTS of offsets (o2, d8 as in oc_cand, k' in place of k), cut into slices by |oy|
// and, within a slice, into the runs' ends (A's own child and B's own child of each run, alternating) and the
// children of both; they are the same for every lane, so all lanes walk them in lock step and a lane only
// masks the children that fall outside the grid.  The order of the walk is free: only the VALUE of the maximum
// is needed here (oc_cand).  sList: shared address of the block's lists, loff / lcnt: the edge's sub-lists.
__device__ __forceinline__ bool oc_maxprod_unit_flat(uint32_t sS, uint32_t sD, uint32_t sList,
                                                     const uint16_t (*__restrict__ loff)[2],
                                                     const uint16_t (*__restrict__ lcnt)[2],
                                                     const uint32_t* __restrict__ row0, int r, int u) {
  const int lane = threadIdx.x & 31;
  const int ix = lane & 15, z0 = 4 * (u & 3) + 2 * (lane >> 4), iy = u >> 2;
  const int iA = iy * 256 + z0 * 16 + ix;             // memory index (iy, iz, ix) of parent A; B = iA + 16
  const double accA = oc_lds64(sD + (uint32_t)iA * 8u), accB = oc_lds64(sD + (uint32_t)(iA + 16) * 8u);
  const bool skipA = accA == 0.0, skipB = accB == 0.0;   // 0 * (finite max) = 0
  double bestA = -INFINITY, bestB = -INFINITY, bestM = -INFINITY;
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  if (__any_sync(0xffffffffu, !(skipA && skipB))) {
    const uint32_t p2 = (skipA && skipB) ? 0u : (uint32_t)((ix + 8) | ((z0 + 8) << 8));
    for (int oy = -r; oy <= r; ++oy) {
      if ((unsigned)(iy + oy) >= 16u) continue;       // (all parents of a task share iy)
      const int aoy = oy < 0 ? -oy : oy;
      const uint32_t base = sS + (uint32_t)((iA + oy * 256) * 8);
      // each sub-list four entries per trip (its length is even), the loaded energies in v0..v3
#define OC_WALK(SUB, EVEN, ODD)                                                                 \
      {                                                                                           \
        uint32_t la = sList + (uint32_t)loff[aoy][SUB] * 8u;                                      \
        const int n = lcnt[aoy][SUB];                                                             \
        int t = 0;                                                                                \
        for (; t + 4 <= n; t += 4, la += 32u) {                                                   \
          const uint4 e = oc_lds128(la), g = oc_lds128(la + 16u); /* the same for every lane */  \
          oc_cand_check(e.x, e.y, p2, base, sS);                                                  \
          oc_cand_check(e.z, e.w, p2, base, sS);                                                  \
          oc_cand_check(g.x, g.y, p2, base, sS);                                                  \
          oc_cand_check(g.z, g.w, p2, base, sS);                                                  \
          oc_cand(e.x, e.y, p2, base, EVEN, v0);                                                  \
          oc_cand(e.z, e.w, p2, base, ODD, v1);                                                   \
          oc_cand(g.x, g.y, p2, base, EVEN, v2);                                                  \
          oc_cand(g.z, g.w, p2, base, ODD, v3);                                                   \
        }                                                                                         \
        if (t < n) {                                                                              \
          const uint4 e = oc_lds128(la);                                                          \
          oc_cand_check(e.x, e.y, p2, base, sS);                                                  \
          oc_cand_check(e.z, e.w, p2, base, sS);                                                  \
          oc_cand(e.x, e.y, p2, base, EVEN, v0);                                                  \
          oc_cand(e.z, e.w, p2, base, ODD, v1);                                                   \
        }                                                                                         \
      }
      OC_WALK(0, bestA, bestB)   // a run's first child is A's alone, the one past its end B's alone: they alternate
      OC_WALK(1, bestM, bestM)   // the children in between are shared
#undef OC_WALK
    }
    // the shared children went into one running maximum (one comparison each); both parents take it at the end
    bestA = bestM > bestA ? bestM : bestA;
    bestB = bestM > bestB ? bestM : bestB;
  }
  // (a finite frame: -inf can only be the initial value, i.e. no allowed child inside the grid)
  bool bad = false;
  {
    const double out = skipA ? 0.0 : accA * oc_factor(row0, 16, 4096, iy, ix, z0, bestA != -INFINITY, bestA);
    bad |= !(fabs(out) <= 1.79769313486231570e308);
    oc_sts64(sD + (uint32_t)iA * 8u, out);
  }
  {
    const double out = skipB ? 0.0 : accB * oc_factor(row0, 16, 4096, iy, ix, z0 + 1, bestB != -INFINITY, bestB);
    bad |= !(fabs(out) <= 1.79769313486231570e308);
    oc_sts64(sD + (uint32_t)(iA + 16) * 8u, out);
  }
  return bad;
}

// EXACT PER-LANE FORM (vectors spilled to scratch, |dz| sets with gaps, or non-finite energies): every lane
// enumerates its own allowed children with the reference's first-candidate rule.
__device__ __forceinline__ bool oc_maxprod_unit_exact(const double* S, double* D, const uint16_t* __restrict__ dz,
                                                      const uint32_t* __restrict__ row0, int n0, int nb0, int r,
                                                      int u, bool skip_ok) {
  const int lane = threadIdx.x & 31;
  const int i = 32 * u + lane;                        // memory index
  if (i >= nb0) return false;
  const bool xfast = n0 == 16;
  const int il = xfast ? oc_swap(i) : i;              // logical bin of this parent
  const int iz = il % n0, qi = il / n0, ix = qi % n0, iy = qi / n0;
  const double acc = D[i];
  if (skip_ok && acc == 0.0) {
    D[i] = 0.0;
    return false;
  }
  const int w = 2 * r + 1;
  int found = -1;
  double best = 0.0;
  const int jy1 = min(iy + r, n0 - 1), jx0 = max(ix - r, 0), jx1 = min(ix + r, n0 - 1);
  for (int jy = max(iy - r, 0); jy <= jy1; ++jy)
    for (int jx = jx0; jx <= jx1; ++jx) {
      const int row = jy * n0 + jx;
      const unsigned dm = dz[abs(iy - jy) * w + (jx - ix) + r];
      const int base = row * n0;
      for (int jz = 0; jz < n0; ++jz)
        if ((dm >> abs(iz - jz)) & 1u) {
          const double v = S[xfast ? oc_swap(base + jz) : base + jz];
          if (found < 0 || v > best) { best = v; found = base + jz; }
        }
    }
  double val;
  int arg;
  oc_finish_max(row0, n0, nb0, iy, ix, iz, found, best, val, arg);
  const double out = acc * val;
  D[i] = out;
  return !(fabs(out) <= 1.79769313486231570e308);
}

// Back-tracking of one edge: the first argmax over the children of parent bin `par` (logical index), from the
// child's final energy vector Sg (scratch, memory order).  Same rule as the sequential enumeration above -- the
// first allowed child is taken unconditionally, a later one only if STRICTLY larger -- evaluated in parallel:
// a partial result is (best, found) = first maximum among the non-NaN children seen, and (first_v, first) = the
// lowest child seen; partial results merge associatively, and a NaN in the very first position wins at the end.
struct OcPick {
  double best, first_v;
  int found, first;
};
__device__ __forceinline__ void oc_pick_merge(OcPick& a, double ob, int of, double ofv, int ofi) {
  if (of >= 0 && (a.found < 0 || ob > a.best || (ob == a.best && of < a.found))) { a.best = ob; a.found = of; }
  if (ofi < a.first) { a.first = ofi; a.first_v = ofv; }
}
template <int WIDTH>
__device__ __forceinline__ void oc_pick_reduce(OcPick& a) {
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, a.best, o), ofv = __shfl_xor_sync(0xffffffffu, a.first_v, o);
    const int of = __shfl_xor_sync(0xffffffffu, a.found, o), ofi = __shfl_xor_sync(0xffffffffu, a.first, o);
    oc_pick_merge(a, ob, of, ofv, ofi);
  }
}
// One thread's share: row t = (jy, jx) of the (2r+1)^2 neighbourhood of `par`, all jz; 8 independent loads at a time.
__device__ __forceinline__ OcPick oc_pick_row(const double* __restrict__ Sg, const uint16_t* __restrict__ dz, int n0,
                                              int r, int par, int t) {
  OcPick a;
  a.best = -INFINITY; a.first_v = 0.0; a.found = -1; a.first = 0x7fffffff;
  const bool xfast = n0 == 16;
  const int iz = par % n0, qi = par / n0, ix = qi % n0, iy = qi / n0;
  const int w = 2 * r + 1;
  if (t >= w * w) return a;
  const int ty = t / w, jy = iy - r + ty, jx = ix - r + (t - ty * w);
  if ((unsigned)jy >= (unsigned)n0 || (unsigned)jx >= (unsigned)n0) return a;
  const unsigned dm = dz[abs(iy - jy) * w + (jx - ix) + r];
  const int base = (jy * n0 + jx) * n0;
  for (int z0 = 0; z0 < n0; z0 += 8) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int jz = z0 + q;
      const bool ok = jz < n0 && ((dm >> abs(iz - jz)) & 1u);
      v[q] = ok ? __ldcg(Sg + (xfast ? oc_swap(base + jz) : base + jz)) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int jz = z0 + q;
      if (jz < n0 && ((dm >> abs(iz - jz)) & 1u)) {   // ascending child index
        if (a.first == 0x7fffffff) { a.first = base + jz; a.first_v = v[q]; }
        if ((a.found < 0 && v[q] == v[q]) || v[q] > a.best) { a.best = v[q]; a.found = base + jz; }
      }
    }
  }
  return a;
}
__device__ __forceinline__ int oc_pick_finish(OcPick a, const uint32_t* __restrict__ row0, int n0, int nb0, int par) {
  if (a.first != 0x7fffffff && a.first_v != a.first_v) { a.best = a.first_v; a.found = a.first; }   // a NaN in front stays
  const int iz = par % n0, qi = par / n0, ix = qi % n0, iy = qi / n0;
  double val;
  int arg;
  oc_finish_max(row0, n0, nb0, iy, ix, iz, a.found, a.best, val, arg);
  return arg;
}

__global__ void __launch_bounds__(kOcThreads, 1) rpsm_onchip_kernel(const RpsmParams p, const OcLayout L) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  OcShared& os = *reinterpret_cast<OcShared*>(smem_raw);
  RpsmShared& s = os.base;
  constexpr int T = kOcThreads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = p.n0, nb0 = n0 * n0 * n0, words0 = (nb0 + 31) / 32;
  const int J = p.J, E = J - 1, V = p.V;
  const int HWm = p.H * p.W;
  const int nbR_ = p.nR * p.nR * p.nR;
  const int nunits = (nb0 + 31) / 32;
  float* stage = reinterpret_cast<float*>(smem_raw + L.stage_off);
  uint16_t* dzm = reinterpret_cast<uint16_t*>(smem_raw + L.dzm_off);
  double* vec_sm = reinterpret_cast<double*>(smem_raw + L.vec_off);
  double* gp = reinterpret_cast<double*>(smem_raw + L.refine_off);
  double* eR = gp + (size_t)J * nbR_ * 3;
  double* sv = eR + (size_t)J * nbR_;
  double* msgR = sv + (size_t)V * J * nbR_;   // (refine_levels8) messages per (edge, parent bin)
  uint8_t* bpR = reinterpret_cast<uint8_t*>(msgR + (size_t)E * nbR_);
  double* coords = L.coords_ws + (size_t)blockIdx.x * V * nb0 * 2;   // (fx, fy) of every (view, bin)
  int32_t* tappos = L.tap_ws + (size_t)blockIdx.x * V * nb0;         // top-left tap, or outside / NaN
  double* spill = L.spill_ws + (size_t)blockIdx.x * L.nspill * L.vec_stride;
  double* sfin = L.sfin_ws + (size_t)blockIdx.x * E * L.vec_stride;
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&os.mbar);
  const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
  const uint32_t vec_s = (uint32_t)__cvta_generic_to_shared(vec_sm);
  const uint32_t list_s = (uint32_t)__cvta_generic_to_shared(smem_raw + L.list_off);
  const uint64_t pol_stream = l2_policy_evict_first();
  const uint64_t pol_keep = l2_policy_evict_last();
  auto vec = [&](int b) -> double* {
    return b < L.nsm ? vec_sm + (size_t)b * L.vec_stride : spill + (size_t)(b - L.nsm) * L.vec_stride;
  };

  // ---- once per block: tree program, shell reach and |dz| sets of every edge, task order -------------
  if (tid < E) { s.edge_p[tid] = p.edges[2 * tid]; s.edge_c[tid] = p.edges[2 * tid + 1]; }
  if (tid < J) s.order[tid] = p.order[tid];
  if (tid < kRpsmMaxJ) os.reach[tid] = 0;
  if (tid == 0) {
    os.stage_tag = -1;
    os.stage_seq = 0;
    oc_mbar_init(mbar);
  }
  __syncthreads();
  if (warp == 0) {   // thread 0 builds the tree program while the other warps scan row 0 of every edge for its reach
    if (tid == 0) oc_build_program(os, J, E, p.root_idx, L.nsm + L.nspill);
    __syncwarp();
    oc_refine_schedule(os, J, E, p.root_idx);
  } else {
    for (int t = tid - 32; t < E * nb0; t += T - 32) {
      const int e = t / nb0, d = t - e * nb0;
      const uint32_t* row0 = p.pair_bits + (size_t)e * nb0 * words0;
      if ((__ldg(row0 + (d >> 5)) >> (d & 31)) & 1u) {
        int dy, dx, dz;
        bin_coords(n0, d, dy, dx, dz);
        atomicMax(&os.reach[e], max(dy, max(dx, dz)));
      }
    }
  }
  __syncthreads();
  if (tid == 0) oc_table_offsets(os, E);
  __syncthreads();
  if (os.prog_err || os.doff[E] > L.dzm_cap) {
    // edges that do not form a tree over all joints, or a max_reach smaller than the table's real reach: the
    // arguments are inconsistent.  Answer NaN for every frame of this block instead of corrupting memory (and
    // instead of __trap(), which would poison the caller's CUDA context).
    for (int f = blockIdx.x; f < p.B; f += gridDim.x)
      for (int t = tid; t < J * 3; t += T) p.out_pose[(size_t)f * J * 3 + t] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  for (int t = tid; t < os.doff[E]; t += T) {
    int e = 0;
    while (t >= os.doff[e + 1]) ++e;
    const int w = 2 * os.reach[e] + 1, local = t - os.doff[e];
    const int adx = abs(local % w - os.reach[e]), ady = local / w;
    const uint32_t* row0 = p.pair_bits + (size_t)e * nb0 * words0;
    uint32_t m = 0u;
    for (int dz = 0; dz < n0; ++dz) {
      const int d = (ady * n0 + adx) * n0 + dz;
      m |= ((__ldg(row0 + (d >> 5)) >> (d & 31)) & 1u) << dz;
    }
    dzm[t] = (uint16_t)m;
  }
  // warp tasks sorted by how much of their neighbourhood lies inside the grid (interior first): the
  // dynamic hand-out below then finishes all warps together
  for (int u = tid; u < nunits; u += T) {
    auto key = [&](int w) {
      const int q = (32 * w + 16) / n0, cx = q % n0, cy = q / n0;
      return abs(2 * cx - (n0 - 1)) + abs(2 * cy - (n0 - 1));
    };
    const int ku = key(u);
    int rank = 0;
    for (int w = 0; w < nunits; ++w) {
      const int kw = key(w);
      rank += (kw < ku) || (kw == ku && w < u);
    }
    os.unit_order[rank] = (uint8_t)u;
  }
  __syncthreads();
  // ---- the edges' child-offset lists (fast form of the max-product, n0 = 16) ------------------------
  // Row (e, |oy|, ox) has the allowed z-offsets K (from its |dz| set), a union of runs.  For the parent pair
  // (z0, z0+1) of a lane the children at z0 + k' are: of A only, k' in K \ (K+1) -- the first of each run; of B
  // only, (K+1) \ K -- one past the end of each run; of both, K & (K+1).  Each slice (e, |oy|) stores two
  // sub-lists, rows in order: the runs' (A-only, B-only) pairs, and the shared children, padded to an even number
  // of entries with one that is outside for every parent.  Bit (k' + 15) of a 32-bit word stands for k' = -15..16.
  {
    int32_t* rowoff = reinterpret_cast<int32_t*>(vec_sm);   // scratch: the energy vectors are not in use yet
    const int rows = os.doff[E];
    auto sub_bits = [](unsigned m, int which) -> uint32_t {   // 0: A only, 1: both, 2: B only
      uint32_t K = 0u;
      for (int d = 0; d < 16; ++d)
        if ((m >> d) & 1u) K |= (1u << (15 + d)) | (1u << (15 - d));
      const uint32_t K1 = K << 1;
      return which == 0 ? (K & ~K1) : which == 1 ? (K & K1) : (K1 & ~K);
    };
    auto entry = [](int ox, int k) {
      return make_uint2((uint32_t)((ox + 8) + (k + 8) * 256), (uint32_t)((k * 16 + ox) * 8));
    };
    const bool can = n0 == 16 && 4 * rows * 4 <= L.nsm * L.vec_stride * 8;   // block-uniform
    int32_t* pre = rowoff + 2 * rows;
    auto row_slot = [&](int row, int& e, int& a, int& x, int& w) {   // row -> (edge, |oy|, column, columns)
      e = 0;
      while (row >= os.doff[e + 1]) ++e;
      w = 2 * os.reach[e] + 1;
      const int local = row - os.doff[e];
      a = local / w;
      x = local - a * w;
    };
    if (can)
      for (int t = tid; t < 2 * rows; t += T) {   // entries per row: two per run, or the shared ones
        const unsigned m = dzm[t % rows];
        rowoff[t] = t < rows ? 2 * __popc(sub_bits(m, 0)) : __popc(sub_bits(m, 1));
      }
    __syncthreads();
    if (can)
      for (int t = tid; t < 2 * rows; t += T) {   // a row's offset within its sub-list; the last row knows the length
        const int sub = t / rows, row = t - sub * rows;
        int e, a, x, w;
        row_slot(row, e, a, x, w);
        int before = 0;
        for (int k = 1; k <= x; ++k) before += rowoff[t - k];
        pre[t] = before;
        if (x == w - 1) {
          const int n = before + rowoff[t];
          os.lcnt[e][a][sub] = (uint16_t)min(n + (n & 1), 65535);
        }
      }
    __syncthreads();
    if (tid == 0 && !can) os.use_flat = 0u;
    else if (tid == 0) {   // the sub-lists one after the other; an edge whose lists no longer fit goes without
      int off = 0;         // (it takes the per-lane enumeration: same results, slower)
      uint32_t have = 0u;
      const int cap = min(L.list_cap, 65535);
      for (int e = 0; e < E; ++e) {
        int need = 0;
        for (int a = 0; a <= os.reach[e]; ++a)
          for (int sub = 0; sub < 2; ++sub) need += os.lcnt[e][a][sub];
        if (off + need > cap) continue;
        have |= 1u << e;
        for (int a = 0; a <= os.reach[e]; ++a)
          for (int sub = 0; sub < 2; ++sub) {
            os.loff[e][a][sub] = (uint16_t)off;
            off += os.lcnt[e][a][sub];
          }
      }
      os.use_flat = have;
      PB_DCHECK(off <= L.list_cap, kDbgRpsmList);
    }
    __syncthreads();
    if (os.use_flat) {
      uint2* list = reinterpret_cast<uint2*>(smem_raw + L.list_off);
      for (int t = tid; t < 2 * rows; t += T) {
        const int sub = t / rows, row = t - sub * rows;
        int e, a, x, w;
        row_slot(row, e, a, x, w);
        if (!((os.use_flat >> e) & 1u)) continue;
        const int ox = x - os.reach[e];
        const int o0 = os.loff[e][a][sub];
        int o = o0 + pre[t];
        if (sub == 0) {
          uint32_t A = sub_bits(dzm[row], 0), B = sub_bits(dzm[row], 2);   // as many bits in one as in the other
          while (A != 0u && B != 0u) {
            list[o++] = entry(ox, __ffs(A) - 16);
            list[o++] = entry(ox, __ffs(B) - 16);
            A &= A - 1u;
            B &= B - 1u;
          }
        } else {
          const uint32_t bits = sub_bits(dzm[row], 1);
          for (int b = 0; b < 32; ++b)
            if ((bits >> b) & 1u) list[o++] = entry(ox, b - 15);
          if (x == w - 1 && ((o - o0) & 1)) list[o] = make_uint2(0x3030u, 0u);   // padding
        }
      }
    }
    // the fast form's 64 warp tasks, longest first: a task walks the slices |oy| its plane iy has inside the grid
    // (z only masks lanes), so its length falls with the plane's distance from the middle
    for (int u = tid; u < 64; u += T) {
      auto key = [](int w) { return abs(2 * (w >> 2) - 15); };
      const int ku = key(u);
      int rank = 0;
      for (int w = 0; w < 64; ++w) {
        const int kw = key(w);
        rank += (kw < ku) || (kw == ku && w < u);
      }
      os.pair_order[rank] = (uint8_t)u;
    }
    __syncthreads();
  }

  const int ngroups = L.stage_views > 0 ? (V + L.stage_views - 1) / L.stage_views : 0;
  // thread 0: hand group g of sampling slot si of frame fr to the copy engine
  auto stage_issue = [&](int fr, int si, int g) {
    const int j = os.samp_joint[si];
    const int v0 = g * L.stage_views, nv = min(V - v0, L.stage_views);
    oc_mbar_expect(mbar, (uint32_t)(nv * HWm * 4));
    for (int v = 0; v < nv; ++v)
      oc_bulk_g2s(stage_s + (uint32_t)(v * HWm * 4),
                  p.hm + (((size_t)fr * V + (v0 + v)) * J + j) * (size_t)HWm, (uint32_t)(HWm * 4), mbar, pol_stream);
    os.stage_tag = ((long long)fr << 16) | ((long long)si << 8) | (long long)g;
    os.stage_seq += 1;
  };
  // all threads: wait until group g of slot si of frame fr is in the stage
  auto stage_acquire = [&](int fr, int si, int g) {
    const long long tag = ((long long)fr << 16) | ((long long)si << 8) | (long long)g;
    if (os.stage_tag != tag) {   // block-uniform: cold start
      if (os.stage_seq > 0) oc_mbar_wait(mbar, (os.stage_seq - 1) & 1u);   // drain the copy in flight
      __syncthreads();
      if (tid == 0) stage_issue(fr, si, g);
      __syncthreads();
    }
    oc_mbar_wait(mbar, (os.stage_seq - 1) & 1u);
    PB_DCHECK(os.stage_tag == tag, kDbgRpsmStage);   // what is about to be sampled is what was copied
  };

  for (int f = blockIdx.x; f < p.B; f += gridDim.x) {
    __syncthreads();
    if (tid < V) {
      load_cam(p.campack + (size_t)p.cam_index[(size_t)f * V + tid] * PB200_CAM_STRIDE, s.cam[tid]);
      for (int k = 0; k < 6; ++k) s.aff[tid][k] = p.box_affine[((size_t)f * V + tid) * 6 + k];
    }
    if (tid < E) s.limb[tid] = p.limb[(size_t)f * E + tid];
    if (tid == 0) os.nonfinite = 0;
    __syncthreads();
    const double centre[3] = {p.root[3 * (size_t)f], p.root[3 * (size_t)f + 1], p.root[3 * (size_t)f + 2]};

    // ---- heatmap coordinates of every (bin, view), once per frame ---------------------------------
    // Slot s = tid + k*T is memory position s of the energy vectors = logical bin slot_bin(s).  On the 16^3 grid
    // that is x-fastest, so the 16 lanes of a half-warp walk along grid x instead of grid z: the vertical world
    // axis projects to an image column, and with the 64-float pitch of a staged map the 16 taps of a column share
    // ONE shared-memory bank; along x they spread over the banks.  The parked taps are stored by slot.
    auto slot_bin = [n0](int sl) { return n0 == 16 ? oc_swap(sl) : sl; };
    for (int sl = tid; sl < nb0; sl += T) {
      const int b = slot_bin(sl);
      double X[3];
      bin_to_point(p.grid_size, n0, b, centre, X);
      for (int v = 0; v < V; ++v) {
        double hx, hy, fx, fy;
        int pos;
        grid_to_heatmap(s.cam[v], s.aff[v], X, p.W, p.H, p.img_w, p.img_h, hx, hy);
        bilinear_prepare(p.W, p.H, hx, hy, pos, fx, fy);   // the joint-independent half of the sample
        st_keep_f64x2(coords + ((size_t)v * nb0 + sl) * 2, fx, fy, pol_keep);
        st_keep_s32(tappos + (size_t)v * nb0 + sl, pos, pol_keep);
      }
    }
    // (every thread reads back only the coordinates it wrote itself: no barrier needed)

    for (int oi = 0; oi < os.nops; ++oi) {
      const OcOp op = os.ops[oi];
      double* D = vec(op.dst);
      if (op.kind != kOpAcc) {   // sample the unary of op.joint: float64 sum over views, in order
        double u[kOcMaxPer];
#pragma unroll
        for (int k = 0; k < kOcMaxPer; ++k) u[k] = 0.0;
        const int W = p.W;
        if (ngroups > 0) {
          for (int g = 0; g < ngroups; ++g) {
            stage_acquire(f, op.samp, g);
            const int v0 = g * L.stage_views, v1 = min(V, v0 + L.stage_views);
#pragma unroll
            for (int k = 0; k < kOcMaxPer; ++k) {
              const int sl = tid + k * T;
              if (sl < nb0)
                for (int v = v0; v < v1; ++v) {
                  const double2 fr = __ldcg(reinterpret_cast<const double2*>(coords + ((size_t)v * nb0 + sl) * 2));
                  const int pos = __ldcg(tappos + (size_t)v * nb0 + sl);
                  const float* m = stage + (size_t)(v - v0) * HWm;
                  u[k] = u[k] + bilinear_apply([m](int t) { return m[t]; }, W, pos, fr.x, fr.y);
                }
            }
            __syncthreads();   // every thread is done with the stage
            if (tid == 0) {    // request what is needed next: it lands under the max-product below
              if (g + 1 < ngroups) stage_issue(f, op.samp, g + 1);
              else if (op.samp + 1 < os.nsamp) stage_issue(f, op.samp + 1, 0);
              else if (f + (int)gridDim.x < p.B) stage_issue(f + (int)gridDim.x, 0, 0);
            }
            if (g + 1 < ngroups) __syncthreads();
          }
        } else {
#pragma unroll
          for (int k = 0; k < kOcMaxPer; ++k) {
            const int sl = tid + k * T;
            if (sl < nb0)
              for (int v = 0; v < V; ++v) {
                const double2 fr = __ldcg(reinterpret_cast<const double2*>(coords + ((size_t)v * nb0 + sl) * 2));
                const int pos = __ldcg(tappos + (size_t)v * nb0 + sl);
                const float* m = p.hm + (((size_t)f * V + v) * J + op.joint) * (size_t)HWm;
                u[k] = u[k] + bilinear_apply([m](int t) { return __ldg(m + t); }, W, pos, fr.x, fr.y);
              }
          }
        }
#pragma unroll
        for (int k = 0; k < kOcMaxPer; ++k) {
          const int sl = tid + k * T;
          if (sl < nb0) {
            if (!(fabs(u[k]) <= 1.79769313486231570e308)) os.nonfinite = 1;   // inf / NaN: no shortcuts
            D[sl] = u[k];   // slot order IS the memory order of the energy vectors
          }
        }
      }
      if (op.kind != kOpLeaf) {
        double* S = vec(op.src);
        if (tid == 0) os.unit_next = 0;
        __syncthreads();
        const int e = op.edge;
        // The child's vector is final: keep a copy for the back-tracking, which asks for the argmax of ONE
        // parent bin per edge once the pose is known (the stores drain underneath the max-product).
        {
          double* keep = sfin + (size_t)e * L.vec_stride;
#pragma unroll
          for (int k = 0; k < kOcMaxPer; ++k) {
            const int sl = tid + k * T;
            if (sl < nb0) __stcg(keep + sl, S[sl]);
          }
        }
        const bool finite = os.nonfinite == 0;
        const bool flat = finite && ((os.use_flat >> e) & 1u) != 0u && op.src < L.nsm && op.dst < L.nsm;
        const uint32_t* row0 = p.pair_bits + (size_t)e * nb0 * words0;
        bool bad = false;
        for (;;) {
          int t = 0;
          if (lane == 0) t = atomicAdd(&os.unit_next, 1);
          t = __shfl_sync(0xffffffffu, t, 0);
          if (t >= (flat ? 64 : nunits)) break;
          const int u = flat ? os.pair_order[t] : os.unit_order[t];
          PB_DCHECK(u >= 0 && u < (flat ? 64 : nunits), kDbgRpsmUnit);
          if (flat)   // everything on chip, every lane walks the edge's offset list
            bad |= oc_maxprod_unit_flat(vec_s + (uint32_t)((size_t)op.src * L.vec_stride * 8),
                                        vec_s + (uint32_t)((size_t)op.dst * L.vec_stride * 8), list_s,
                                        os.loff[e], os.lcnt[e], row0, os.reach[e], u);
          else
            bad |= oc_maxprod_unit_exact(S, D, dzm + os.doff[e], row0, n0, nb0, os.reach[e], u, finite);
        }
        if (bad) os.nonfinite = 1;
      }
      __syncthreads();
    }

    // ---- root argmax (first maximum), then back-tracking level by level, one warp per edge ---------
    {
      const double* er = vec(os.root_buf);
      double best = -INFINITY;
      int bidx = 0x7fffffff;
      for (int m = tid; m < nb0; m += T) {   // memory order; the first maximum is by LOGICAL index
        const double v = er[m];
        const int l = slot_bin(m);
        if (bidx == 0x7fffffff || v > best || (v == best && l < bidx)) { best = v; bidx = l; }
      }
      warp_first_max(best, bidx);
      if (lane == 0) { s.red_val[warp] = best; s.red_idx[warp] = bidx; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < T / 32; ++w)
          if (s.red_val[w] > best || (s.red_val[w] == best && s.red_idx[w] < bidx)) {
            best = s.red_val[w];
            bidx = s.red_idx[w];
          }
        s.bin[p.root_idx] = bidx;
      }
      __syncthreads();
      // parents before children; up to four edges of one depth at a time, eight warps (256 rows) each
      constexpr int kGroup = 8, kGroups = T / 32 / kGroup;
      const int g = warp / kGroup, tg = tid - g * (kGroup * 32);
      for (int d = 1; d <= os.max_depth; ++d)
        for (int c = os.bt_start[d]; c < os.bt_start[d + 1]; c += kGroups) {
          const bool on = c + g < os.bt_start[d + 1];
          const int e = on ? os.bt_edge[c + g] : 0;
          const int par = s.bin[s.edge_p[e]];
          if (on) {
            const int r = os.reach[e], ww = (2 * r + 1) * (2 * r + 1);
            OcPick a;
            a.best = -INFINITY; a.first_v = 0.0; a.found = -1; a.first = 0x7fffffff;
            for (int t = tg; t < ww; t += kGroup * 32) {
              const OcPick b = oc_pick_row(sfin + (size_t)e * L.vec_stride, dzm + os.doff[e], n0, r, par, t);
              oc_pick_merge(a, b.best, b.found, b.first_v, b.first);
            }
            oc_pick_reduce<32>(a);
            if (lane == 0) { os.bt_val[warp] = a.best; os.bt_idx[warp] = a.found; os.bt_fval[warp] = a.first_v; os.bt_fidx[warp] = a.first; }
          }
          __syncthreads();
          if (on && warp == g * kGroup) {
            OcPick a;
            const int w8 = g * kGroup + (lane & (kGroup - 1));
            a.best = os.bt_val[w8]; a.found = os.bt_idx[w8]; a.first_v = os.bt_fval[w8]; a.first = os.bt_fidx[w8];
            oc_pick_reduce<kGroup>(a);
            if (lane == 0) {
              const int b = oc_pick_finish(a, p.pair_bits + (size_t)e * nb0 * words0, n0, nb0, par);
              PB_DCHECK(b >= 0 && b < nb0, kDbgRpsmArg);
              s.bin[s.edge_c[e]] = b;
            }
          }
          __syncthreads();
        }
    }

    if (tid < J) {
      double X[3];
      bin_to_point(p.grid_size, n0, s.bin[tid], centre, X);
      s.pose[tid][0] = X[0]; s.pose[tid][1] = X[1]; s.pose[tid][2] = X[2];
      if (p.out_trace) p.out_trace[((size_t)f * (p.depth + 1)) * J + tid] = s.bin[tid];
    }
    __syncthreads();
    if (p.nR == 2) refine_levels8<T>(p, os, f, eR, sv, msgR, bpR);
    else refine_levels<T>(p, s, f, gp, eR, sv, bpR);
    if (tid < J) {
      double* o = p.out_pose + ((size_t)f * J + tid) * 3;
      o[0] = s.pose[tid][0]; o[1] = s.pose[tid][1]; o[2] = s.pose[tid][2];
    }
  }
  // every bulk copy that was issued has been waited for: stage_issue only looks ahead to frames this
  // block will process, and stage_acquire waits for each before its data is read
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

// *flag |= 1 unless P[e][i][j] depends on (|dy|,|dx|,|dz|) only, i.e. equals row 0 at that offset
__global__ void pairwise_lut_check_kernel(const uint32_t* __restrict__ bits, int E, int n, int* flag) {
  const int nb = n * n * n, words = (nb + 31) / 32;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)E * nb * words) return;
  const int w = (int)(t % words);
  const int i = (int)((t / words) % nb);
  const int e = (int)(t / ((long long)words * nb));
  const uint32_t* row0 = bits + (size_t)e * nb * words;
  const uint32_t m = bits[t];
  int iy, ix, iz;
  bin_coords(n, i, iy, ix, iz);
  bool bad = false;
  for (int k = 0; k < 32; ++k) {
    const int j = w * 32 + k;
    if (j >= nb) break;
    int jy, jx, jz;
    bin_coords(n, j, jy, jx, jz);
    const int d = (abs(iy - jy) * n + abs(ix - jx)) * n + abs(iz - jz);
    const uint32_t want = (row0[d >> 5] >> (d & 31)) & 1u;
    bad |= ((m >> k) & 1u) != want;
  }
  if (bad) atomicOr(flag, 1);
  // flag[1] = largest |index offset| of an allowed pair in row 0 of any edge (the shell "reach")
  if (i == 0 && m != 0u) {
    int r = 0;
    for (int k = 0; k < 32; ++k) {
      const int j = w * 32 + k;
      if (j >= nb || !((m >> k) & 1u)) continue;
      int jy, jx, jz;
      bin_coords(n, j, jy, jx, jz);
      r = max(r, max(jy, max(jx, jz)));
    }
    atomicMax(flag + 1, r);
  }
}

static int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

static size_t rpsm_smem_bytes(int nb0, int J, int V, int nbR) {
  const size_t words0 = (nb0 + 31) / 32;
  size_t b = ((sizeof(RpsmShared) + 15) / 16) * 16 + (size_t)nb0 * (sizeof(double) + 2 * sizeof(uint32_t)) +
             words0 * sizeof(uint32_t);
  b += (size_t)next_pow2(nb0) * sizeof(uint16_t) + 32;   // + alignment slack
  b += (size_t)J * nbR * (3 + 1 + V) * sizeof(double) + (size_t)(J - 1) * nbR;
  return b;
}

static int rpsm_slots(int B, int n_sm) {
  const int cap = n_sm * 2;
  return B < cap ? (B > 0 ? B : 1) : cap;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// live energy vectors the depth-first program can need for ANY tree of J joints: one accumulator per
// ancestor with a second child pending (each costs the tree two more nodes) + source + destination
static int onchip_max_vectors(int J) { return (J - 1) / 2 + 2; }

static size_t onchip_workspace_bytes(int blocks, int V, int J, int nb0, int n0, int nspill) {
  const size_t vec_stride = align_up((size_t)nb0, 2);
  size_t b = align_up((size_t)blocks * V * nb0 * 2 * sizeof(double), 256);
  b += align_up((size_t)blocks * V * nb0 * sizeof(int32_t), 256);
  b += align_up((size_t)blocks * nspill * vec_stride * sizeof(double), 256);
  b += align_up((size_t)blocks * (J - 1) * vec_stride * sizeof(double), 256);
  return b;
}

// Shared-memory plan of rpsm_onchip_kernel; returns false when the shape does not fit the on-chip path.
static bool rpsm_onchip_layout(const float* hm, int V, int J, int H, int W, int n0, int nbR, int max_reach,
                               OcLayout& L) {
  const int nb0 = n0 * n0 * n0, E = J - 1;
  if (n0 > 16 || nb0 > kOcMaxPer * kOcThreads || max_reach < 0 || max_reach > kRpsmEnumReach) return false;
  const size_t map_bytes = (size_t)H * W * 4;
  size_t off = align_up(sizeof(OcShared), 128);
  L.stage_views = 0;
  if ((H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(hm) & 15u) == 0 && map_bytes <= (size_t)kOcStageBytes) {
    L.stage_views = (int)((size_t)kOcStageBytes / map_bytes);
    if (L.stage_views > V) L.stage_views = V;
  }
  L.stage_off = (int)off;
  off = align_up(off + (size_t)L.stage_views * map_bytes, 128);
  L.dzm_off = (int)off;
  L.dzm_cap = E * (max_reach + 1) * (2 * max_reach + 1);
  off = align_up(off + (size_t)L.dzm_cap * sizeof(uint16_t), 16);
  // the refinement's arrays (grid points, energies, samples, messages, back pointers): they are only alive
  // between the back-tracking of a frame and the first sample of the next, when the energy vectors are dead, so
  // they share the vectors' memory whenever they fit there (always on the reference's 16^3 grid)
  const size_t refine_bytes =
      align_up((size_t)J * nbR * (3 + 1 + V) * sizeof(double) + (size_t)E * nbR * sizeof(double) + (size_t)E * nbR, 16);
  L.vec_stride = (int)align_up((size_t)nb0, 2);
  const size_t vec_bytes = (size_t)L.vec_stride * sizeof(double);
  const int need = onchip_max_vectors(J);
  int nsm = 0;
  for (int alias = 1; alias >= 0; --alias) {
    const size_t voff = alias ? off : off + refine_bytes;
    if (voff + 2 * vec_bytes > (size_t)kOcSmemBudget) {   // source + destination must be on chip
      if (alias) continue;
      return false;
    }
    const int fit = (int)(((size_t)kOcSmemBudget - voff) / vec_bytes);
    // leave ~16 KiB for the child-offset lists unless that would push one of the 4 vectors the reference
    // skeletons need off the chip
    nsm = (int)(((size_t)kOcSmemBudget - voff - 16 * 1024) / vec_bytes);
    if (nsm < 4) nsm = fit < 4 ? fit : 4;
    if (nsm > need) nsm = need;
    if (alias && (size_t)nsm * vec_bytes < refine_bytes) continue;
    L.refine_off = (int)(alias ? voff : off);
    off = voff;
    break;
  }
  L.vec_off = (int)off;
  L.nsm = nsm;
  L.nspill = need - nsm;
  L.list_off = (int)(off + (size_t)nsm * vec_bytes);
  L.list_cap = (int)(((size_t)kOcSmemBudget - (size_t)L.list_off) / 8);
  L.smem_bytes = (size_t)L.list_off + (size_t)L.list_cap * 8;
  return true;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_rpsm_workspace_bytes(int B, int J, int first_nbins, int n_sm) {
  if (B <= 0 || J < 2 || first_nbins < 1 || n_sm < 1) return 0;
  const size_t nb0 = (size_t)first_nbins * first_nbins * first_nbins;
  const size_t slots = (size_t)rpsm_slots(B, n_sm);
  const size_t e_bytes = ((slots * J * nb0 * sizeof(double) + 255) / 256) * 256;
  const size_t generic = e_bytes + slots * (size_t)(J - 1) * nb0 * sizeof(uint16_t);
  // the on-chip path: one block per SM, coordinates for up to PB200_MAX_VIEWS views, every vector spilled
  const size_t onchip = onchip_workspace_bytes(B < n_sm ? B : n_sm, PB200_MAX_VIEWS, J, (int)nb0, first_nbins,
                                               onchip_max_vectors(J));
  return generic > onchip ? generic : onchip;
}

extern "C" int pb200_rpsm(const float* hm, int B, int V, int J, int H, int W, const double* campack,
                          const int32_t* cam_index, const double* box_affine, int img_w, int img_h,
                          const double* root, const double* limb, const int32_t* edges,
                          const int32_t* order, int root_idx, const uint32_t* pair_bits, int use_lut,
                          int max_reach, int first_nbins, int recur_nbins, int recur_depth, double grid_size,
                          double tolerance, void* workspace, size_t workspace_bytes,
                          double* out_pose, int32_t* out_trace, void* stream) {
  PB_REQUIRE(B >= 0 && H >= 2 && W >= 2, "bad shape B=%d H=%d W=%d", B, H, W);
  if (B == 0) return PB200_OK;
  PB_REQUIRE(hm && campack && cam_index && box_affine && root && limb && edges && order && pair_bits,
             "null input pointer");
  PB_REQUIRE(out_pose && workspace, "null output / workspace pointer");
  PB_REQUIRE(V >= 1 && V <= PB200_MAX_VIEWS, "V=%d outside [1,%d]", V, PB200_MAX_VIEWS);
  PB_REQUIRE(J >= 2 && J <= kRpsmMaxJ, "J=%d outside [2,%d]", J, kRpsmMaxJ);
  PB_REQUIRE(root_idx >= 0 && root_idx < J, "root_idx out of range");
  PB_REQUIRE(first_nbins >= 1 && first_nbins * first_nbins * first_nbins <= kRpsmMaxBins0,
             "first_nbins^3 must be <= %d", kRpsmMaxBins0);
  PB_REQUIRE(recur_nbins >= 1 && recur_nbins * recur_nbins * recur_nbins <= kRpsmMaxBinsR,
             "recur_nbins^3 must be <= %d", kRpsmMaxBinsR);
  PB_REQUIRE(recur_depth >= 0, "recur_depth < 0");
  PB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  const int nb0 = first_nbins * first_nbins * first_nbins;
  const int nbR = recur_nbins * recur_nbins * recur_nbins;
  PB_REQUIRE(workspace_bytes >= pb200_rpsm_workspace_bytes(B, J, first_nbins, sm),
             "workspace too small: %zu < %zu", workspace_bytes, pb200_rpsm_workspace_bytes(B, J, first_nbins, sm));
  RpsmParams p;
  p.hm = hm; p.B = B; p.V = V; p.J = J; p.H = H; p.W = W;
  p.campack = campack; p.cam_index = cam_index; p.box_affine = box_affine;
  p.img_w = (double)img_w; p.img_h = (double)img_h;
  p.root = root; p.limb = limb; p.edges = edges; p.order = order; p.root_idx = root_idx;
  p.pair_bits = pair_bits; p.use_lut = use_lut;
  p.n0 = first_nbins; p.nR = recur_nbins; p.depth = recur_depth; p.npad = next_pow2(nb0);
  p.grid_size = grid_size; p.tol = tolerance;
  p.energy_ws = nullptr; p.bp_ws = nullptr;
  p.out_pose = out_pose; p.out_trace = out_trace;

  // ---- on-chip path: offset-only table, shells within reach, n0 <= 16 ------------------------------
  OcLayout L;
  if (use_lut && rpsm_onchip_layout(hm, V, J, H, W, first_nbins, nbR, max_reach, L)) {
    const int blocks = B < sm ? B : sm;   // one persistent block per SM
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    L.coords_ws = reinterpret_cast<double*>(w);
    w += align_up((size_t)blocks * V * nb0 * 2 * sizeof(double), 256);
    L.tap_ws = reinterpret_cast<int32_t*>(w);
    w += align_up((size_t)blocks * V * nb0 * sizeof(int32_t), 256);
    L.spill_ws = reinterpret_cast<double*>(w);
    w += align_up((size_t)blocks * L.nspill * L.vec_stride * sizeof(double), 256);
    L.sfin_ws = reinterpret_cast<double*>(w);
    static PerDevice<size_t> attr_set;
    size_t* have = attr_set.slot();
    if (have == nullptr) return PB200_ERR_CUDA;
    if (*have < L.smem_bytes) {
      PB_CUDA(cudaFuncSetAttribute(rpsm_onchip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOcSmemBudget));
      *have = (size_t)kOcSmemBudget;
    }
    rpsm_onchip_kernel<<<blocks, kOcThreads, L.smem_bytes, (cudaStream_t)stream>>>(p, L);
    PB_LAUNCH_CHECK("rpsm_onchip_kernel");
    return PB200_OK;
  }

  // ---- generic path: arbitrary bit matrices, wide shells, large grids ------------------------------
  const size_t smem = rpsm_smem_bytes(nb0, J, V, nbR);
  PB_REQUIRE(smem <= 227 * 1024, "first_nbins=%d needs %zu bytes of shared memory", first_nbins, smem);
  static PerDevice<size_t> generic_attr;
  size_t* have = generic_attr.slot();
  if (have == nullptr) return PB200_ERR_CUDA;
  if (*have < smem) {
    PB_CUDA(cudaFuncSetAttribute(rpsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *have = smem;
  }
  const int slots = rpsm_slots(B, sm);
  p.energy_ws = reinterpret_cast<double*>(workspace);
  const size_t e_bytes = (((size_t)slots * J * nb0 * sizeof(double) + 255) / 256) * 256;
  p.bp_ws = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(workspace) + e_bytes);
  rpsm_kernel<<<slots, kRpsmThreads, smem, (cudaStream_t)stream>>>(p);
  PB_LAUNCH_CHECK("rpsm_kernel");
  return PB200_OK;
}

extern "C" int pb200_pairwise_level0(const double* avg_limb, int E, int nbins, double box_size,
                                     uint32_t* pair_bits, void* stream) {
  PB_REQUIRE(avg_limb && pair_bits, "null pointer");
  PB_REQUIRE(E >= 1 && nbins >= 1 && nbins * nbins * nbins <= kRpsmMaxBins0, "bad E=%d nbins=%d", E, nbins);
  const long long nb = (long long)nbins * nbins * nbins, words = (nb + 31) / 32;
  const long long n = (long long)E * nb * words;
  pairwise_level0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      avg_limb, E, nbins, box_size, pair_bits);
  PB_LAUNCH_CHECK("pairwise_level0_kernel");
  return PB200_OK;
}

extern "C" int pb200_pairwise_lut_check(const uint32_t* pair_bits, int E, int nbins, int32_t* out_flag,
                                        void* stream) {
  PB_REQUIRE(pair_bits && out_flag, "null pointer");
  PB_REQUIRE(E >= 1 && nbins >= 1 && nbins * nbins * nbins <= kRpsmMaxBins0, "bad E=%d nbins=%d", E, nbins);
  const long long nb = (long long)nbins * nbins * nbins, words = (nb + 31) / 32;
  const long long n = (long long)E * nb * words;
  pairwise_lut_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pair_bits, E, nbins, out_flag);
  PB_LAUNCH_CHECK("pairwise_lut_check_kernel");
  return PB200_OK;
}

PB_DEFINE_DEBUG_READER(rpsm)
