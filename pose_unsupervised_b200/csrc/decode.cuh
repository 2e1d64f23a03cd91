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
// kReadOnly: the map was not written by this kernel, so the neighbour loads may use the
// non-coherent read-only path (__ldg).
template <bool kReadOnly = true>
__device__ __forceinline__ DecodeOut finish_map(const ArgMax am, const float* base, int H,
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
        float dx, dy;
        if (kReadOnly) {
          dx = __ldg(c + 1) - __ldg(c - 1);
          dy = __ldg(c + W) - __ldg(c - W);
        } else {
          const volatile float* vc = c;
          dx = vc[1] - vc[-1];
          dy = vc[W] - vc[-W];
        }
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

// The same epilogue split in two so that the TMA front end can overlap it with the next map:
// begin_finish issues the four neighbour loads of the quarter-pixel shift (L2 hits) and
// returns without waiting for them; end_finish consumes them.
struct PendingMap {
  int m;  // map index, -1 = none
  const float* base;
  ArgMax am;
  float fx, fy;          // masked heatmap coordinates
  float xp, xm, yp, ym;  // hm[py][px+1], hm[py][px-1], hm[py+1][px], hm[py-1][px]
  bool refine;
  Affine6 aff;
  int ticket;            // whatever the caller's publish step returned (e.g. an atomic's result)
};

__device__ __forceinline__ void begin_finish(PendingMap& pd, int m, const float* __restrict__ base,
                                             const ArgMax am, int H, int W, bool has_trans,
                                             bool post_process, const Affine6& aff) {
  pd.m = m;
  pd.base = base;
  pd.am = am;
  pd.aff = aff;
  float fx = (float)(am.idx % W);
  float fy = (float)(am.idx / W);
  if (!(am.val > 0.0f)) { fx = 0.0f; fy = 0.0f; }  // pred_mask (inference.py:43-46)
  pd.fx = fx;
  pd.fy = fy;
  pd.refine = false;
  pd.xp = pd.xm = pd.yp = pd.ym = 0.0f;
  if (has_trans && post_process) {
    const int px = (int)floorf(fx + 0.5f), py = (int)floorf(fy + 0.5f);
    if (1 < px && px < W - 1 && 1 < py && py < H - 1) {
      const float* c = base + py * W + px;
      pd.refine = true;
      pd.xp = __ldg(c + 1); pd.xm = __ldg(c - 1);
      pd.yp = __ldg(c + W); pd.ym = __ldg(c - W);
    }
  }
}

__device__ __forceinline__ DecodeOut end_finish(const PendingMap& pd, bool has_trans) {
  DecodeOut o;
  o.idx = pd.am.idx;
  o.maxval = pd.am.val;
  float fx = pd.fx, fy = pd.fy;
  if (has_trans) {
    if (pd.refine) {
      const float dx = pd.xp - pd.xm, dy = pd.yp - pd.ym;
      const float sx = dx > 0.f ? 1.f : (dx < 0.f ? -1.f : (dx == 0.f ? 0.f : dx));
      const float sy = dy > 0.f ? 1.f : (dy < 0.f ? -1.f : (dy == 0.f ? 0.f : dy));
      fx += sx * 0.25f;
      fy += sy * 0.25f;
    }
    const double dxx = (double)fx, dyy = (double)fy;
    const double ox = fma(dyy, pd.aff.t[1], dxx * pd.aff.t[0]) + pd.aff.t[2];
    const double oy = fma(dyy, pd.aff.t[4], dxx * pd.aff.t[3]) + pd.aff.t[5];
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

// ---------------------------------------------------------------------------
// TMA front end: every warp owns a ring of kStages x 4 KiB in shared memory that the
// copy engine fills with cp.async.bulk (SASS: UBLKCP) and signals through an mbarrier
// per stage.  The warp that consumes a stage re-arms it with the next chunk of its
// current map or, once that is fully issued, of its NEXT map, so kStages chunks per
// warp stay in flight across map boundaries and across the per-map epilogue,
// independent of registers.  Needs 16-byte aligned maps with H*W % 4 == 0.
// ---------------------------------------------------------------------------
#ifndef PB_CHUNK_FLOATS
#define PB_CHUNK_FLOATS 1024  // 4 KiB per bulk copy
#endif
#ifndef PB_STAGES
#define PB_STAGES 2
#endif
#ifndef PB_PIPE_EPILOGUE
#define PB_PIPE_EPILOGUE 0  // 0: epilogue in line; 1: overlapped with the next map's chunks;
                            // 2: in line, but complete(m, ticket) runs one map late
#endif
#ifndef PB_SCAN_HOIST
#define PB_SCAN_HOIST 0     // 1: unguarded fast path for full chunks (lets ptxas batch the LDS)
#endif
constexpr int kChunkFloats = PB_CHUNK_FLOATS;
constexpr int kChunkBytes = kChunkFloats * 4;
constexpr int kStages = PB_STAGES;

__host__ __device__ constexpr size_t tma_ring_smem_bytes(int warps) {
  return (size_t)warps * kStages * kChunkBytes + (size_t)warps * kStages * 8;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct MapCursor {
  int m;              // claimed map index (>= total: none)
  const float* base;  // its first element
  int issued;         // chunks handed to the copy engine so far
};

// Streams the maps handed out by `claim` (warp-uniform: returns the next map index, >= total
// when there is none) through this warp's ring and decodes them.  The per-map epilogue is
// software-pipelined against the NEXT map's chunks:
//   after the warp reduction   begin_finish: neighbour loads issued, not awaited
//   after the next chunk 0     end_finish + publish(m, out) -> ticket   (stores / atomic issued)
//   after the next chunk 1     complete(m, ticket)                      (ticket consumed, e.g. lift)
// so the L2 and atomic round trips never stall the stream.
//   affine: [rows][6] or null (get_max_preds); warps_per_block: ring slots in `smem_raw`.
template <typename Claim, typename Publish, typename Complete>
__device__ __forceinline__ void stream_maps_tma(unsigned char* smem_raw, int warps_per_block,
                                                const HmViews& hv, int J, int H, int W, int total,
                                                const double* __restrict__ affine, bool post_process,
                                                Claim claim, Publish publish, Complete complete) {
  const int HW = H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kStages * kChunkFloats;
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(smem_raw + (size_t)warps_per_block * kStages * kChunkBytes) + warp * kStages;
  const uint32_t ring_s = smem_u32(ring), bars_s = smem_u32(bars);
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bars_s + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int nchunk = (HW + kChunkFloats - 1) / kChunkFloats;
  const int last_floats = HW - (nchunk - 1) * kChunkFloats;
  unsigned q_issue = 0, q_cons = 0;  // chunk sequence numbers of this warp (stage = q % kStages)

  auto set_map = [&](MapCursor& c, int m) {
    c.m = m;
    c.issued = 0;
    c.base = nullptr;
    if (m < total) {
      const int row = m / J;
      c.base = map_base(hv, row, m - row * J, J, HW);
    }
  };
  // hand the next chunk of `c` to the copy engine (lane 0 issues; cursors are warp-uniform)
  auto issue = [&](MapCursor& c) {
    const int k = c.issued;
    const uint32_t bytes = (k == nchunk - 1 ? last_floats : kChunkFloats) * 4;
    const unsigned s = q_issue % kStages;
    if (lane == 0) {
      mbar_expect_tx(bars_s + 8 * s, bytes);
      bulk_g2s(ring_s + s * kChunkBytes, c.base + (size_t)k * kChunkFloats, bytes, bars_s + 8 * s);
    }
    ++c.issued;
    ++q_issue;
  };

  MapCursor cur, nxt;
  set_map(cur, claim());
  set_map(nxt, total);
  bool nxt_claimed = false;
  // keep the ring full: chunks of the current map first, then of the next claimed map
  auto top_up = [&]() {
    while (q_issue - q_cons < (unsigned)kStages) {
      if (cur.m < total && cur.issued < nchunk) { issue(cur); continue; }
      if (!nxt_claimed) {
        set_map(nxt, claim());
        nxt_claimed = true;
      }
      if (nxt.m < total && nxt.issued < nchunk) { issue(nxt); continue; }
      break;
    }
  };

  PendingMap pend;
  pend.m = -1;
  int late_m = -1, late_ticket = 0;
  const bool has_trans = affine != nullptr;
  const int k_complete = nchunk > 1 ? 1 : 0;
  while (cur.m < total) {
    Affine6 aff;
    if (has_trans) aff = load_affine(affine + 6 * (size_t)(cur.m / J));  // consumed two stages later
    float best = -INFINITY;
    int bidx = 4 * lane < HW ? 4 * lane : 0x7fffffff;
    bool nanp = false;
    for (int k = 0; k < nchunk; ++k) {
      top_up();
      const unsigned s = q_cons % kStages;
      mbar_wait(bars_s + 8 * s, (q_cons / kStages) & 1u);
      // The compare chain reads the stage directly, one float4 at a time.  Two "obvious"
      // improvements were measured on B200 and rejected (profiles/r01_sweep_ring.txt): copying
      // the chunk to registers and re-arming the stage before the compares (+37 % time), and
      // an unguarded fast path for full chunks, which lets ptxas hoist all eight LDS.128 ahead
      // of the compares (+35 % time at 64x64).
      const float4* src = reinterpret_cast<const float4*>(ring + s * kChunkFloats);
      const int nvec = (k == nchunk - 1 ? last_floats : kChunkFloats) >> 2;
      const int e0 = k * kChunkFloats + 4 * lane;
#define PB_SCAN_F4(q, e)                                                        \
  nanp |= (q.x != q.x) | (q.y != q.y) | (q.z != q.z) | (q.w != q.w);           \
  if (q.x > best) { best = q.x; bidx = (e); }                                   \
  if (q.y > best) { best = q.y; bidx = (e) + 1; }                               \
  if (q.z > best) { best = q.z; bidx = (e) + 2; }                               \
  if (q.w > best) { best = q.w; bidx = (e) + 3; }
      if (PB_SCAN_HOIST && nvec == kChunkFloats / 4) {
#pragma unroll
        for (int u = 0; u < kChunkFloats / 128; ++u) {
          const float4 q = src[u * 32 + lane];
          PB_SCAN_F4(q, e0 + 128 * u)
        }
      } else {
#pragma unroll
        for (int u = 0; u < kChunkFloats / 128; ++u) {
          const int vi = u * 32 + lane;
          if (vi < nvec) {
            const float4 q = src[vi];
            PB_SCAN_F4(q, e0 + 128 * u)
          }
        }
      }
#undef PB_SCAN_F4
      __syncwarp();  // every lane is done with stage s before it is refilled
      ++q_cons;
      if (PB_PIPE_EPILOGUE == 1 && pend.m >= 0) {  // previous map's epilogue, interleaved with this map's chunks
        if (k == 0) pend.ticket = publish(pend.m, end_finish(pend, has_trans));
        if (k == k_complete) { complete(pend.m, pend.ticket); pend.m = -1; }
      }
    }
    top_up();  // the next map's first chunks fly while this one is finished
    ArgMax am;
    if (__any_sync(0xffffffffu, nanp)) {  // rare: exact numpy NaN rules, re-read from L2
      scan_map_exact(cur.base, HW, lane, best, bidx);
      am = warp_argmax<true>(best, bidx);
    } else {
      am = warp_argmax<false>(best, bidx);
    }
    begin_finish(pend, cur.m, cur.base, am, H, W, has_trans, post_process, aff);
    if (PB_PIPE_EPILOGUE == 0) {
      pend.ticket = publish(pend.m, end_finish(pend, has_trans));
      complete(pend.m, pend.ticket);
      pend.m = -1;
    } else if (PB_PIPE_EPILOGUE == 2) {
      // publish now (its stores and atomic are issued), consume the PREVIOUS map's ticket: by
      // now that atomic has long returned, so no round trip is ever waited for
      const int ticket = publish(pend.m, end_finish(pend, has_trans));
      if (late_m >= 0) complete(late_m, late_ticket);
      late_m = pend.m;
      late_ticket = ticket;
      pend.m = -1;
    }
    if (!nxt_claimed) set_map(nxt, claim());
    cur = nxt;
    set_map(nxt, total);
    nxt_claimed = false;
  }
  if (pend.m >= 0) {  // drain
    pend.ticket = publish(pend.m, end_finish(pend, has_trans));
    complete(pend.m, pend.ticket);
  }
  if (late_m >= 0) complete(late_m, late_ticket);
}

}  // namespace pb200
#endif
