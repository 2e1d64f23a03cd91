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
  double red_val[kRpsmThreads / 32];
  int red_idx[kRpsmThreads / 32];
  int reach;
};

__device__ __forceinline__ double sample_view(const RpsmParams& p, const RpsmShared& s, int f, int v,
                                              int j, const double* X) {
  double hx, hy;
  grid_to_heatmap(s.cam[v], s.aff[v], X, p.W, p.H, p.img_w, p.img_h, hx, hy);
  const float* m = p.hm + (((size_t)f * p.V + v) * p.J + j) * (size_t)(p.H * p.W);
  const int W = p.W;
  return bilinear_zero_outside([m, W](int y, int x) { return __ldg(m + y * W + x); }, p.W, p.H, hx, hy);
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
          u = u + bilinear_zero_outside([m, W](int y, int x) { return __ldg(m + y * W + x); }, p.W, p.H,
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

    // ---- refinement levels ----------------------------------------------------------
    double cur = p.grid_size / (double)n0;
    for (int lvl = 1; lvl <= p.depth; ++lvl) {
      for (int t = tid; t < J * nbR; t += kRpsmThreads) {
        const int j = t / nbR, b = t - j * nbR;
        double X[3];
        bin_to_point(cur, nR, b, s.pose[j], X);
        gp[3 * t] = X[0]; gp[3 * t + 1] = X[1]; gp[3 * t + 2] = X[2];
      }
      __syncthreads();
      // one thread per (view, joint, bin) sample, then the ordered sum over views
      for (int t = tid; t < V * J * nbR; t += kRpsmThreads) {
        const int v = t / (J * nbR), r = t - v * (J * nbR);
        sv[t] = sample_view(p, s, f, v, r / nbR, gp + 3 * r);
      }
      __syncthreads();
      for (int t = tid; t < J * nbR; t += kRpsmThreads) {
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
                          const int32_t* order, int root_idx, const uint32_t* pair_bits, int use_lut,
                          int first_nbins, int recur_nbins, int recur_depth, double grid_size,
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
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  const int nb0 = first_nbins * first_nbins * first_nbins;
  PB_REQUIRE(workspace_bytes >= pb200_rpsm_workspace_bytes(B, J, first_nbins, sm),
             "workspace too small: %zu < %zu", workspace_bytes, pb200_rpsm_workspace_bytes(B, J, first_nbins, sm));
  const size_t smem = rpsm_smem_bytes(nb0, J, V, recur_nbins * recur_nbins * recur_nbins);
  PB_REQUIRE(smem <= 227 * 1024, "first_nbins=%d needs %zu bytes of shared memory", first_nbins, smem);
  PB_CUDA(cudaFuncSetAttribute(rpsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int slots = rpsm_slots(B, sm);
  RpsmParams p;
  p.hm = hm; p.B = B; p.V = V; p.J = J; p.H = H; p.W = W;
  p.campack = campack; p.cam_index = cam_index; p.box_affine = box_affine;
  p.img_w = (double)img_w; p.img_h = (double)img_h;
  p.root = root; p.limb = limb; p.edges = edges; p.order = order; p.root_idx = root_idx;
  p.pair_bits = pair_bits; p.use_lut = use_lut;
  p.n0 = first_nbins; p.nR = recur_nbins; p.depth = recur_depth; p.npad = next_pow2(nb0);
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
