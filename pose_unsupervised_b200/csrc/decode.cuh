// decode.cuh -- one warp decodes one heatmap: flat argmax with numpy's tie rules,
// quarter-pixel shift, inverse crop affine.  Shared by decode.cu and lift_fused.cu.
//
// Reference: lib/core/inference.py:19-75.
//   * np.argmax returns the FIRST maximal flat index and treats NaN as maximal
//     (first NaN wins); -0.0 == +0.0.
//   * coordinates are zeroed where maxval <= 0 (or NaN);
//   * POST_PROCESS moves by 0.25*sign(neighbour difference) when 1 < px < W-1 and
//     1 < py < H-1;
//   * the result is [x, y, 1] @ trans.T in float64, stored as float32.
#ifndef PB200_DECODE_CUH_
#define PB200_DECODE_CUH_

#include "pb_common.cuh"

namespace pb200 {

// Streaming 128-bit load: read-only path, do not allocate in L1 (each heatmap byte
// is used once; the few re-reads of the refinement come from L2).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p));
  return v;
}

struct ArgMax {
  float val;
  int idx;
};

// candidate b beats candidate a?  (value descending, then index ascending; NaN-free)
__device__ __forceinline__ bool beats_fast(float bv, int bi, float av, int ai) {
  return (bv > av) || (bv == av && bi < ai);
}

// numpy order with NaN: any NaN beats any number; among NaNs the lower index wins.
__device__ __forceinline__ bool beats_nan(float bv, int bi, float av, int ai) {
  const bool bn = bv != bv, an = av != av;
  if (bn != an) return bn;
  if (bn) return bi < ai;
  return (bv > av) || (bv == av && bi < ai);
}

template <bool kNan>
__device__ __forceinline__ ArgMax warp_argmax(float v, int i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, i, off);
    const bool take = kNan ? beats_nan(ov, oi, v, i) : beats_fast(ov, oi, v, i);
    if (take) { v = ov; i = oi; }
  }
  return ArgMax{v, i};
}

#define PB_DECODE_UNROLL 8

// Fast path: HW % 4 == 0 and a 16-byte aligned map.  Lane l owns float4 number
// l, l+32, ... so its elements are visited in ascending index order and a strict
// '>' keeps the lane-local first maximum.  NaNs are only detected here (flag);
// the caller falls back to scan_map_exact for such maps.
__device__ __forceinline__ void scan_map_vec4(const float* __restrict__ base, int nvec, int lane,
                                              float& best, int& bidx, bool& saw_nan) {
  const float4* p = reinterpret_cast<const float4*>(base);
  best = -INFINITY;
  bidx = lane < nvec ? 4 * lane : 0x7fffffff;
  bool nanp = false;
  constexpr int U = PB_DECODE_UNROLL;
  int v0 = 0;
  for (; v0 + 32 * U <= nvec; v0 += 32 * U) {
    float4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = ld_stream_f4(p + v0 + u * 32 + lane);
    const int e0 = 4 * (v0 + lane);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + 128 * u;
      const float4 q = r[u];
      nanp |= (q.x != q.x) | (q.y != q.y) | (q.z != q.z) | (q.w != q.w);
      if (q.x > best) { best = q.x; bidx = e; }
      if (q.y > best) { best = q.y; bidx = e + 1; }
      if (q.z > best) { best = q.z; bidx = e + 2; }
      if (q.w > best) { best = q.w; bidx = e + 3; }
    }
  }
  if (v0 < nvec) {  // ragged tail (e.g. 80x80 maps): guarded loads
    float4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int vi = v0 + u * 32 + lane;
      r[u] = vi < nvec ? ld_stream_f4(p + vi) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    const int e0 = 4 * (v0 + lane);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + 128 * u;
      const float4 q = r[u];
      nanp |= (q.x != q.x) | (q.y != q.y) | (q.z != q.z) | (q.w != q.w);
      if (q.x > best) { best = q.x; bidx = e; }
      if (q.y > best) { best = q.y; bidx = e + 1; }
      if (q.z > best) { best = q.z; bidx = e + 2; }
      if (q.w > best) { best = q.w; bidx = e + 3; }
    }
  }
  saw_nan = nanp;
}

// Exact scalar path: any size / alignment, full numpy NaN semantics.  Lane l owns
// elements l, l+32, ...
__device__ __forceinline__ void scan_map_exact(const float* __restrict__ base, int hw, int lane,
                                               float& best, int& bidx) {
  best = -INFINITY;
  bidx = lane < hw ? lane : 0x7fffffff;
  for (int e = lane; e < hw; e += 32) {
    const float v = __ldg(base + e);
    const bool best_nan = best != best;
    if (!best_nan && ((v > best) || (v != v))) { best = v; bidx = e; }
  }
}

struct DecodeOut {
  float x, y, maxval;
  int idx;
};

// Whole-warp flat argmax of one map straight from global memory; every lane gets the result.
__device__ __forceinline__ ArgMax scan_map(const float* __restrict__ base, int hw, bool vec_ok, int lane) {
  float best;
  int bidx;
  bool need_exact = !vec_ok;
  if (vec_ok) {
    bool saw_nan;
    scan_map_vec4(base, hw >> 2, lane, best, bidx, saw_nan);
    need_exact = __any_sync(0xffffffffu, saw_nan);
    if (!need_exact) return warp_argmax<false>(best, bidx);
  }
  scan_map_exact(base, hw, lane, best, bidx);
  return warp_argmax<true>(best, bidx);
}

struct Affine6 {
  double t[6];
};

__device__ __forceinline__ Affine6 load_affine(const double* __restrict__ trans) {
  Affine6 a;
#pragma unroll
  for (int k = 0; k < 6; ++k) a.t[k] = __ldg(trans + k);
  return a;
}

// Coordinates from the argmax: mask, quarter-pixel shift, inverse crop affine.
__device__ __forceinline__ DecodeOut finish_map(const ArgMax am, const float* __restrict__ base, int H,
                                                int W, bool has_trans, const Affine6& a,
                                                bool post_process) {
  DecodeOut o;
  o.idx = am.idx;
  o.maxval = am.val;
  float fx = (float)(am.idx % W);
  float fy = (float)(am.idx / W);
  if (!(am.val > 0.0f)) { fx = 0.0f; fy = 0.0f; }  // pred_mask (inference.py:43-46)
  if (has_trans) {
    if (post_process) {
      const int px = (int)floorf(fx + 0.5f), py = (int)floorf(fy + 0.5f);
      if (1 < px && px < W - 1 && 1 < py && py < H - 1) {
        const float* c = base + py * W + px;
        const float dx = __ldg(c + 1) - __ldg(c - 1);
        const float dy = __ldg(c + W) - __ldg(c - W);
        // np.sign: -1, 0, +1, NaN for NaN
        const float sx = dx > 0.f ? 1.f : (dx < 0.f ? -1.f : (dx == 0.f ? 0.f : dx));
        const float sy = dy > 0.f ? 1.f : (dy < 0.f ? -1.f : (dy == 0.f ? 0.f : dy));
        fx += sx * 0.25f;
        fy += sy * 0.25f;
      }
    }
    // [x, y, 1] @ trans.T in float64 (BLAS accumulation order), stored float32
    const double dxx = (double)fx, dyy = (double)fy;
    const double ox = fma(dyy, a.t[1], dxx * a.t[0]) + a.t[2];
    const double oy = fma(dyy, a.t[4], dxx * a.t[3]) + a.t[5];
    fx = (float)ox;
    fy = (float)oy;
  }
  o.x = fx;
  o.y = fy;
  return o;
}

// Whole-warp decode of one map; every lane returns the same result.
__device__ __forceinline__ DecodeOut decode_map(const float* __restrict__ base, int H, int W, bool vec_ok,
                                                const double* __restrict__ trans /* 6 or null */,
                                                bool post_process, int lane) {
  Affine6 a;
  if (trans != nullptr) a = load_affine(trans);  // issued before the scan: its latency is hidden
  const ArgMax am = scan_map(base, H * W, vec_ok, lane);
  return finish_map(am, base, H, W, trans != nullptr, a, post_process);
}

}  // namespace pb200
#endif
