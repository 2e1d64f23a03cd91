// decode.cu -- K1: batched heatmap decode (get_max_preds / get_final_preds) and the
// crop-affine pre-pass.  HBM-bound: every heatmap byte is read exactly once.
//
// Launch shape: 256-thread blocks (8 warps), one warp per map, warps stride over the
// N*J maps; grid = min(ceil(maps/8), blocks_per_sm * #SM) so that the grid is a whole
// number of resident waves on the 148 SMs.
#include "decode.cuh"

namespace pb200 {

constexpr int kDecodeWarps = 8;
// PB200_TUNE_DECODE_SCHEDULE (pb200_set_tuning).  Dynamic (fixed strided share + claimed tail) is the default:
// 0.639 vs 0.647 ms alone on one GPU (the claimed tail evens out the end of the kernel) and it absorbs blocks that
// start late under an overlapped collective (profiles/r02_scaling.md)
static int g_decode_schedule = PB200_DECODE_DYNAMIC;

__global__ void __launch_bounds__(kDecodeWarps * 32)
decode_kernel(HmViews hv, int N, int J, int H, int W, int vec_ok,
              const double* __restrict__ affine, int post_process,
              float* __restrict__ out_xy, float* __restrict__ out_maxval,
              int32_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long total = (long long)N * J;
  const int HW = H * W;
  for (long long m = (long long)blockIdx.x * kDecodeWarps + warp; m < total;
       m += (long long)gridDim.x * kDecodeWarps) {
    const int row = (int)(m / J), j = (int)(m % J);
    const float* base = map_base(hv, row, j, J, HW);
    const DecodeOut o = decode_map(base, H, W, vec_ok != 0,
                                   affine ? affine + 6 * (size_t)row : nullptr,
                                   post_process != 0, lane);
    if (lane == 0) {
      reinterpret_cast<float2*>(out_xy)[m] = make_float2(o.x, o.y);
      out_maxval[m] = o.maxval;
      if (out_idx) out_idx[m] = o.idx;
    }
  }
}

// Same decode with the TMA ring front end (decode.cuh::stream_maps_tma).
//   claim_ctr == nullptr : warps take maps blockIdx*8+warp, +gridDim*8, ... (static)
//   claim_ctr != nullptr : after a fixed strided share, warps pull small BATCHES of map indices from
//     claim_ctr[0] (same-address atomics are served at ~0.2 per ns on B200: one claim per map for the whole
//     kernel would saturate that and cost 50 %).  A block that starts late (because another kernel, e.g. an
//     NCCL collective overlapping the start of the step, still holds its SM) then simply decodes fewer maps
//     instead of finishing late.  The last block out (claim_ctr[1] counts them) zeroes both words, so the pair
//     is clean for the next launch.
#ifndef PB_DECODE_MIN_BLOCKS
#define PB_DECODE_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(kDecodeWarps * 32, PB_DECODE_MIN_BLOCKS)
decode_tma_kernel(HmViews hv, int N, int J, int H, int W, const double* __restrict__ affine,
                  int post_process, float* __restrict__ out_xy, float* __restrict__ out_maxval,
                  int32_t* __restrict__ out_idx, int* __restrict__ claim_ctr) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long total_ll = (long long)N * J;
  const int total = (int)total_ll;
  const long long stride = (long long)gridDim.x * kDecodeWarps;
  long long next = (long long)blockIdx.x * kDecodeWarps + warp;
  // dynamic form: every warp first takes a FIXED, strided share (7/8 of an even split: map w + k * nwarps, the
  // same interleaving as the static form) and only the rest comes from the counter, in small batches:
  // [cur_m, cur_end) is the open batch (warp-uniform), ahead_m0 (lane 0) / ahead_b the batch claimed ahead of
  // time -- its atomic is issued when the previous batch is opened (the first one at kernel start), so no warp
  // ever waits for a round trip.  A block that starts up to ~1/8 of the kernel's duration late is absorbed.
  const int nwarps_total = (int)gridDim.x * kDecodeWarps;
  const int my_warp = (int)blockIdx.x * kDecodeWarps + warp;
  const int fixed_per_warp = (int)((total_ll / nwarps_total) * 7 / 8);
  const int first_free = fixed_per_warp * nwarps_total;
  int k_fixed = 0, cur_m = 0, cur_end = 0, ahead_m0 = 0;
  int ahead_b = min(4, max(1, (total - first_free) / nwarps_total));
  if (claim_ctr != nullptr && lane == 0) ahead_m0 = atomicAdd(claim_ctr, ahead_b) + first_free;
  stream_maps_tma(
      smem_raw, kDecodeWarps, hv, J, H, W, total, affine, post_process != 0,
      [&]() {
        if (claim_ctr != nullptr) {
          if (k_fixed < fixed_per_warp) return my_warp + (k_fixed++) * nwarps_total;
          if (cur_m >= cur_end) {   // open the batch claimed earlier, claim the one after it
            cur_m = __shfl_sync(0xffffffffu, ahead_m0, 0);
            cur_end = cur_m + ahead_b;
            const int left = total - cur_end;
            ahead_b = left <= 0 ? 1 : min(4, max(1, left / nwarps_total));
            if (lane == 0) ahead_m0 = atomicAdd(claim_ctr, ahead_b) + first_free;
          }
          const int m = cur_m++;
          return m < total ? m : total;
        }
        const long long m = next;
        next += stride;
        return m < total_ll ? (int)m : total;
      },
      [&](int m, const DecodeOut& o) {
        if (lane == 0) {
          PB_DCHECK(m >= 0 && m < total, kDbgDecodeMapRange);
#if PB200_DEBUG_CHECKS
          if (claim_ctr != nullptr) atomicAdd(claim_ctr + 2, 1);   // third word of the slot: maps decoded by this launch
#endif
          reinterpret_cast<float2*>(out_xy)[m] = make_float2(o.x, o.y);
          out_maxval[m] = o.maxval;
          if (out_idx) out_idx[m] = o.idx;
        }
        return 0;
      },
      [&](int, int) {});
  if (claim_ctr != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(claim_ctr + 1, 1) == (int)gridDim.x - 1) {
      claim_ctr[0] = 0;
      claim_ctr[1] = 0;
#if PB200_DEBUG_CHECKS
      __threadfence();
      PB_DCHECK(atomicExch(claim_ctr + 2, 0) == total, kDbgDecodeMapCount);   // as many decodes as maps
#endif
    }
  }
}

// ---- claim counters for the dynamic form -----------------------------------------------------------
// A pool of (counter, blocks-done, [debug build: maps decoded], pad) slots per device, zeroed once; every launch takes its own pair, so
// launches that overlap on different streams never share one: eager launches cycle through the first half
// (a pair is back to zero long before it comes round again), launches recorded into a CUDA graph take a
// pair of the second half for good (a graph cannot run concurrently with itself).  No pair left, or the
// pool not yet allocated while a capture is in progress -> the static form, same results.
namespace {
constexpr int kClaimPairs = 4096;
struct ClaimPool {
  int* dev;
  unsigned eager, captured;
};
int* take_claim_pair(cudaStream_t stream) {
  static PerDevice<ClaimPool> pools;
  ClaimPool* pool = pools.slot();
  if (pool == nullptr) return nullptr;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  const bool capturing = st != cudaStreamCaptureStatusNone;
  if (pool->dev == nullptr) {
    if (capturing) return nullptr;                       // cannot zero the pool inside a capture
    int* d = nullptr;
    if (cudaMalloc(&d, sizeof(int) * 4 * kClaimPairs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(d, 0, sizeof(int) * 4 * kClaimPairs) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return nullptr; }
    pool->dev = d;                                       // 64 KiB per device, kept for the life of the process
  }
  if (capturing) {
    const unsigned k = __sync_fetch_and_add(&pool->captured, 1u);
    if (k >= (unsigned)kClaimPairs / 2) return nullptr;
    return pool->dev + 4 * (kClaimPairs / 2 + k);
  }
  return pool->dev + 4 * (__sync_fetch_and_add(&pool->eager, 1u) % (kClaimPairs / 2));
}
}  // namespace

__global__ void crop_affine_kernel(const void* center, int c_f64, const void* scale, int s_f64,
                                   const double* __restrict__ rot_sincos, double shift_x, double shift_y,
                                   int shift_f64, int n, int out_w, int out_h, int inv,
                                   double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  CropSpec q = crop_spec_plain();
  if (rot_sincos != nullptr) { q.sn = rot_sincos[2 * i]; q.cs = rot_sincos[2 * i + 1]; }
  q.shift[0] = shift_x; q.shift[1] = shift_y; q.shift_f64 = shift_f64 != 0;
  double m[6];
  crop_affine_row(center, c_f64, scale, s_f64, i, q, out_w, out_h, inv, m);
#pragma unroll
  for (int k = 0; k < 6; ++k) out[6 * (size_t)i + k] = m[k];
}

int fill_views(const float* const* hm_views_host, int n_ptr, int N, HmViews& hv) {
  PB_REQUIRE(hm_views_host != nullptr, "hm_views_host is null");
  PB_REQUIRE(n_ptr >= 1 && n_ptr <= PB200_MAX_VIEWS, "n_ptr=%d outside [1,%d]", n_ptr, PB200_MAX_VIEWS);
  PB_REQUIRE(N % n_ptr == 0, "N=%d is not a multiple of the %d view tensors", N, n_ptr);
  hv.n = n_ptr;
  for (int v = 0; v < PB200_MAX_VIEWS; ++v) hv.ptr[v] = v < n_ptr ? hm_views_host[v] : nullptr;
  for (int v = 0; v < n_ptr; ++v) PB_REQUIRE(hv.ptr[v] != nullptr, "heatmap pointer %d is null", v);
  return PB200_OK;
}

bool views_vec_ok(const HmViews& hv, int HW) {
  if (HW % 4 != 0) return false;
  for (int v = 0; v < hv.n; ++v)
    if ((reinterpret_cast<uintptr_t>(hv.ptr[v]) & 15u) != 0) return false;
  return true;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_set_tuning(int key, int value) {
  PB_REQUIRE(key == PB200_TUNE_DECODE_SCHEDULE, "unknown tuning key %d", key);
  PB_REQUIRE(value == PB200_DECODE_STATIC || value == PB200_DECODE_DYNAMIC, "decode schedule must be 0 (static) or 1 (dynamic)");
  pb200::g_decode_schedule = value;
  return PB200_OK;
}

extern "C" int pb200_crop_affine(const void* center, int center_dtype, const void* scale,
                                 int scale_dtype, const double* rot_sincos, double shift_x,
                                 double shift_y, int shift_dtype, int n, int out_w, int out_h, int inv,
                                 double* out, void* stream) {
  PB_REQUIRE(center && scale && out, "null pointer");
  PB_REQUIRE(n >= 0 && out_w > 0 && out_h > 0, "bad sizes n=%d out=(%d,%d)", n, out_w, out_h);
  PB_REQUIRE((center_dtype | 1) == 1 && (scale_dtype | 1) == 1 && (shift_dtype | 1) == 1,
             "dtype tags must be PB200_F32/PB200_F64");
  if (n == 0) return PB200_OK;
  crop_affine_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      center, center_dtype, scale, scale_dtype, rot_sincos, shift_x, shift_y, shift_dtype, n, out_w, out_h,
      inv, out);
  PB_LAUNCH_CHECK("crop_affine_kernel");
  return PB200_OK;
}

namespace pb200 {
int launch_decode(const HmViews& hv, int N, int J, int H, int W, const double* affine, int post_process,
                  float* out_xy, float* out_maxval, int32_t* out_idx, void* stream);
}

extern "C" int pb200_decode(const float* const* hm_views_host, int n_ptr, int N, int J, int H, int W,
                            const double* affine, int post_process, float* out_xy, float* out_maxval,
                            int32_t* out_idx, void* stream) {
  PB_REQUIRE(N >= 0 && J >= 1 && H >= 1 && W >= 1, "bad shape N=%d J=%d H=%d W=%d", N, J, H, W);
  PB_REQUIRE((long long)H * W < (1LL << 24), "map of %dx%d exceeds the float32-exact index range", H, W);
  if (N == 0) return PB200_OK;
  PB_REQUIRE(out_xy && out_maxval, "null output");
  HmViews hv;
  int rc = fill_views(hm_views_host, n_ptr, N, hv);
  if (rc != PB200_OK) return rc;
  return launch_decode(hv, N, J, H, W, affine, post_process, out_xy, out_maxval, out_idx, stream);
}

int pb200::launch_decode(const HmViews& hv, int N, int J, int H, int W, const double* affine,
                         int post_process, float* out_xy, float* out_maxval, int32_t* out_idx,
                         void* stream) {
  const long long maps = (long long)N * J;
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  PB_REQUIRE(maps < (1LL << 31) - (1 << 20), "N*J too large for one launch; split the batch");
  if (views_vec_ok(hv, H * W)) {  // TMA ring front end
    const size_t smem = tma_ring_smem_bytes(kDecodeWarps);
    static PerDevice<int> tma_blocks_per_sm;
    int* per_sm = tma_blocks_per_sm.slot();
    if (per_sm == nullptr) return PB200_ERR_CUDA;
    if (*per_sm == 0) {
      PB_CUDA(cudaFuncSetAttribute(decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int n = 0;
      PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_tma_kernel, kDecodeWarps * 32, smem));
      *per_sm = n > 0 ? n : 1;
    }
    long long blocks = (maps + kDecodeWarps - 1) / kDecodeWarps;
    const long long cap = (long long)sm * *per_sm;
    if (blocks > cap) blocks = cap;
    // dynamic claims only when the grid is the full persistent wave (otherwise every warp has one map)
    int* claim = (blocks == cap && g_decode_schedule == PB200_DECODE_DYNAMIC) ? take_claim_pair((cudaStream_t)stream) : nullptr;
    decode_tma_kernel<<<(unsigned)blocks, kDecodeWarps * 32, smem, (cudaStream_t)stream>>>(
        hv, N, J, H, W, affine, post_process, out_xy, out_maxval, out_idx, claim);
    PB_LAUNCH_CHECK("decode_tma_kernel");
    return PB200_OK;
  }
  static PerDevice<int> ldg_blocks_per_sm;
  int* per_sm = ldg_blocks_per_sm.slot();
  if (per_sm == nullptr) return PB200_ERR_CUDA;
  if (*per_sm == 0) {
    int n = 0;
    PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_kernel, kDecodeWarps * 32, 0));
    *per_sm = n > 0 ? n : 1;
  }
  long long blocks = (maps + kDecodeWarps - 1) / kDecodeWarps;
  const long long cap = (long long)sm * *per_sm;  // one resident wave
  if (blocks > cap) blocks = cap;
  decode_kernel<<<(unsigned)blocks, kDecodeWarps * 32, 0, (cudaStream_t)stream>>>(
      hv, N, J, H, W, views_vec_ok(hv, H * W) ? 1 : 0, affine, post_process, out_xy, out_maxval,
      out_idx);
  PB_LAUNCH_CHECK("decode_kernel");
  return PB200_OK;
}

// ---- transform_preds on already decoded coordinates ----------------------------------
namespace pb200 {
template <typename T>
__global__ void transform_preds_kernel(const T* __restrict__ coords, const double* __restrict__ affine,
                                       long long n, int J, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* t = affine + 6 * (i / J);
  const double x = (double)coords[2 * i], y = (double)coords[2 * i + 1];
  out[2 * i] = fma(y, t[1], x * t[0]) + t[2];
  out[2 * i + 1] = fma(y, t[4], x * t[3]) + t[5];
}
}  // namespace pb200

extern "C" int pb200_transform_preds(const void* coords, int coords_dtype, const double* affine, int N,
                                     int J, double* out, void* stream) {
  PB_REQUIRE(coords && affine && out, "null pointer");
  PB_REQUIRE(N >= 0 && J >= 1, "bad shape");
  PB_REQUIRE((coords_dtype | 1) == 1, "coords_dtype must be PB200_F32/PB200_F64");
  const long long n = (long long)N * J;
  if (n == 0) return PB200_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (coords_dtype == PB200_F32)
    pb200::transform_preds_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)coords, affine, n, J, out);
  else
    pb200::transform_preds_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)coords, affine, n, J, out);
  PB_LAUNCH_CHECK("transform_preds_kernel");
  return PB200_OK;
}

// ---- flip-test averaging fused with the decode (lib/core/function.py:567-583 + :632-640) ----------
// validate() runs the network on the mirrored input, flips the result back (mirror the columns, swap
// left/right joints), optionally shifts it one column to the right (TEST.SHIFT_HEATMAP) and averages
// it with the plain output before decoding.  Here one warp per (row, joint) reads both maps once,
// writes the averaged map (it is also what goes into the h5 file, function.py:640,673) and decodes it
// on the fly -- instead of flip / index_select / clone / add / mul kernels plus a second pass.
namespace pb200 {
__global__ void __launch_bounds__(256)
decode_flip_kernel(HmViews hv, HmViews hf, int N, int J, int H, int W, int vec,
                   const int32_t* __restrict__ joint_src, int shift, const double* __restrict__ affine,
                   int post_process,
                   float* __restrict__ out_avg, float* __restrict__ out_xy, float* __restrict__ out_maxval,
                   int32_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long total = (long long)N * J;
  const int HW = H * W;
  for (long long m = (long long)blockIdx.x * 8 + warp; m < total; m += (long long)gridDim.x * 8) {
    const int row = (int)(m / J), j = (int)(m % J);
    const float* a = map_base(hv, row, j, J, HW);
    const float* b = map_base(hf, row, joint_src[j], J, HW);
    float* o = out_avg + (size_t)m * HW;
    float best = -INFINITY;
    int bidx = 0x7fffffff;
#define PB_FLIP_UPD(v, e)                                              \
  {                                                                    \
    const bool best_nan = best != best;                                \
    if (bidx == 0x7fffffff || (!best_nan && (((v) > best) || ((v) != (v))))) { best = (v); bidx = (e); } \
  }
    if (vec) {  // W % 4 == 0, 16-byte aligned tensors: 128-bit loads and stores
      const float4* a4 = reinterpret_cast<const float4*>(a);
      float4* o4 = reinterpret_cast<float4*>(o);
      for (int i = lane; i < (HW >> 2); i += 32) {
        const int e = 4 * i, y = e / W, x0 = e - y * W;
        const float* brow = b + y * W;
        const float4 q = a4[i];
        // mirrored columns W-1-x; shifted right by one they become W-x (x >= 1) and W-1 for x = 0
        const float4 lo = *reinterpret_cast<const float4*>(brow + (W - x0 - 4));   // columns W-x0-4 .. W-x0-1
        float f0, f1, f2, f3;
        if (!shift) { f0 = lo.w; f1 = lo.z; f2 = lo.y; f3 = lo.x; }
        else {
          f1 = lo.w; f2 = lo.z; f3 = lo.y;
          f0 = x0 == 0 ? lo.w : __ldg(brow + (W - x0));
        }
        float4 r;
        r.x = (q.x + f0) * 0.5f; r.y = (q.y + f1) * 0.5f; r.z = (q.z + f2) * 0.5f; r.w = (q.w + f3) * 0.5f;
        o4[i] = r;
        PB_FLIP_UPD(r.x, e) PB_FLIP_UPD(r.y, e + 1) PB_FLIP_UPD(r.z, e + 2) PB_FLIP_UPD(r.w, e + 3)
      }
    } else {
      for (int e = lane; e < HW; e += 32) {
        const int y = e / W, x = e - y * W;
        const int xs = shift ? (x >= 1 ? W - x : W - 1) : W - 1 - x;
        const float v = (__ldg(a + e) + __ldg(b + y * W + xs)) * 0.5f;
        o[e] = v;
        PB_FLIP_UPD(v, e)
      }
    }
#undef PB_FLIP_UPD
    const ArgMax am = warp_argmax<true>(best, bidx);
    __syncwarp();  // the averaged map is read back for the quarter-pixel shift
    Affine6 aff;
    if (affine) aff = load_affine(affine + 6 * (size_t)row);
    const DecodeOut r = finish_map<false>(am, o, H, W, affine != nullptr, aff, post_process != 0);
    if (lane == 0) {
      reinterpret_cast<float2*>(out_xy)[m] = make_float2(r.x, r.y);
      out_maxval[m] = r.maxval;
      if (out_idx) out_idx[m] = r.idx;
    }
  }
}
}  // namespace pb200

extern "C" int pb200_decode_flip(const float* const* hm_views_host, const float* const* hm_flip_views_host,
                                 int n_ptr, int N, int J, int H, int W, const int32_t* joint_src,
                                 int shift_heatmap, const double* affine, int post_process,
                                 float* out_avg, float* out_xy, float* out_maxval, int32_t* out_idx,
                                 void* stream) {
  PB_REQUIRE(N >= 0 && J >= 1 && H >= 1 && W >= 1, "bad shape N=%d J=%d H=%d W=%d", N, J, H, W);
  PB_REQUIRE((long long)H * W < (1LL << 24), "map of %dx%d exceeds the float32-exact index range", H, W);
  if (N == 0) return PB200_OK;
  PB_REQUIRE(joint_src && out_avg && out_xy && out_maxval, "null pointer");
  HmViews hv, hf;
  int rc = fill_views(hm_views_host, n_ptr, N, hv);
  if (rc != PB200_OK) return rc;
  rc = fill_views(hm_flip_views_host, n_ptr, N, hf);
  if (rc != PB200_OK) return rc;
  const long long maps = (long long)N * J;
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  long long blocks = (maps + 7) / 8;
  const long long cap = (long long)sm * 8;
  if (blocks > cap) blocks = cap;
  const int vec = (W % 4 == 0) && views_vec_ok(hv, H * W) && views_vec_ok(hf, H * W) &&
                  ((reinterpret_cast<uintptr_t>(out_avg) & 15u) == 0);
  decode_flip_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      hv, hf, N, J, H, W, vec, joint_src, shift_heatmap, affine, post_process, out_avg, out_xy, out_maxval,
      out_idx);
  PB_LAUNCH_CHECK("decode_flip_kernel");
  return PB200_OK;
}

PB_DEFINE_DEBUG_READER(decode)
