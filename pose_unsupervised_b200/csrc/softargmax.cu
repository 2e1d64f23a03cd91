// softargmax.cu -- the differentiable side of the epipolar term ("next" row 2 of SURVEY.md section 8f):
// soft-argmax 2D joints from heatmaps (lib/utils/transforms.py:149-171, generate_integral_preds_2d_th)
// and the gradient of the fundamental loss (lib/core/loss.py:101-133) with respect to the 2D joints.
//
//   forward   p = softmax(beta * hm) over the map (beta = 100), x = sum_e p_e * col(e), y = sum_e p_e * row(e)
//   backward  d hm_e = beta * p_e * ((col(e) - x) * gx + (row(e) - y) * gy)
// One warp per map, 128-bit loads; the forward reads the map twice (max, then sums -- the second read is
// an L2 hit), the backward reads it once and writes the gradient map: both are HBM-bound passes.
#include "pb_common.cuh"

namespace pb200 {

// exp of a non-positive argument (logit minus the running maximum): the hardware ex2.approx path
// (2 ulp) is ample for a softmax whose result is compared at 1e-3 px, and it keeps both passes on
// the bandwidth side of the roofline instead of the SFU/FMA side.
__device__ __forceinline__ float pb_exp(float x) { return __expf(x); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
softargmax_fwd_kernel(const float* __restrict__ hm, long long maps, int H, int W, float beta, int vec,
                      float* __restrict__ out_xy, float* __restrict__ out_stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = H * W;
  for (long long m = (long long)blockIdx.x * 8 + warp; m < maps; m += (long long)gridDim.x * 8) {
    const float* base = hm + (size_t)m * HW;
    float mx = -INFINITY;
    float z = 0.f, sx = 0.f, sy = 0.f;
    if (vec) {  // W % 4 == 0 and 16-byte aligned maps: a float4 never straddles a row
      // ONE pass (online softmax): a running maximum per lane; the partial sums are rescaled
      // whenever it grows, and once more when the lanes are merged
      const float4* b4 = reinterpret_cast<const float4*>(base);
      for (int i = lane; i < (HW >> 2); i += 32) {
        const float4 q = b4[i];
        const float v0 = q.x * beta, v1 = q.y * beta, v2 = q.z * beta, v3 = q.w * beta;
        const float m4 = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
        if (m4 > mx) {
          const float r = pb_exp(mx - m4);   // 0 on the first chunk (mx = -inf)
          z *= r; sx *= r; sy *= r;
          mx = m4;
        }
        const int e = 4 * i, y = e / W, x = e - y * W;
        const float p0 = pb_exp(v0 - mx), p1 = pb_exp(v1 - mx), p2 = pb_exp(v2 - mx), p3 = pb_exp(v3 - mx);
        const float ps = (p0 + p1) + (p2 + p3);
        z += ps;
        sx += ps * (float)x + (p1 + 2.f * p2 + 3.f * p3);
        sy = fmaf(ps, (float)y, sy);
      }
      const float gm = warp_max(mx);
      const float r = (mx == -INFINITY) ? 0.f : pb_exp(mx - gm);   // lanes without elements contribute 0
      z *= r; sx *= r; sy *= r;
      mx = gm;
    } else {
      for (int e = lane; e < HW; e += 32) mx = fmaxf(mx, base[e] * beta);
      mx = warp_max(mx);
      for (int e = lane; e < HW; e += 32) {
        const float p = pb_exp(base[e] * beta - mx);
        const int y = e / W, x = e - y * W;
        z += p;
        sx = fmaf(p, (float)x, sx);
        sy = fmaf(p, (float)y, sy);
      }
    }
    z = warp_sum(z);
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    if (lane == 0) {
      out_xy[2 * m] = sx / z;
      out_xy[2 * m + 1] = sy / z;
      out_stats[2 * m] = mx;
      out_stats[2 * m + 1] = z;
    }
  }
}

__global__ void __launch_bounds__(256)
softargmax_bwd_kernel(const float* __restrict__ hm, const float* __restrict__ stats,
                      const float* __restrict__ xy, const float* __restrict__ grad_xy, long long maps,
                      int H, int W, float beta, int vec, float* __restrict__ grad_hm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = H * W;
  for (long long m = (long long)blockIdx.x * 8 + warp; m < maps; m += (long long)gridDim.x * 8) {
    const float* base = hm + (size_t)m * HW;
    float* g = grad_hm + (size_t)m * HW;
    const float mx = stats[2 * m], rz = 1.0f / stats[2 * m + 1];
    const float x0 = xy[2 * m], y0 = xy[2 * m + 1];
    const float gx = grad_xy[2 * m] * beta, gy = grad_xy[2 * m + 1] * beta;
    if (vec) {
      const float4* b4 = reinterpret_cast<const float4*>(base);
      float4* g4 = reinterpret_cast<float4*>(g);
      for (int i = lane; i < (HW >> 2); i += 32) {
        const float4 q = b4[i];
        const int e = 4 * i, y = e / W, x = e - y * W;
        const float ty = ((float)y - y0) * gy;
        float4 o;
        o.x = pb_exp(q.x * beta - mx) * rz * (((float)x - x0) * gx + ty);
        o.y = pb_exp(q.y * beta - mx) * rz * (((float)(x + 1) - x0) * gx + ty);
        o.z = pb_exp(q.z * beta - mx) * rz * (((float)(x + 2) - x0) * gx + ty);
        o.w = pb_exp(q.w * beta - mx) * rz * (((float)(x + 3) - x0) * gx + ty);
        g4[i] = o;
      }
    } else {
      for (int e = lane; e < HW; e += 32) {
        const float p = pb_exp(base[e] * beta - mx) * rz;
        const int y = e / W, x = e - y * W;
        g[e] = p * (((float)x - x0) * gx + ((float)y - y0) * gy);
      }
    }
  }
}

// thread per (frame, ordered pair, joint): d |x_b^T F x_a| * w_b * w_a * gscale, scattered with atomics
template <typename T, typename TW>
__global__ void __launch_bounds__(256)
epipolar_grad_kernel(const double* __restrict__ fmat, const int32_t* __restrict__ subj,
                     const T* __restrict__ xy, const TW* __restrict__ w, int B, int V, int J,
                     double gscale, double* __restrict__ grad_xy) {
  const int P = V * (V - 1);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * P * J) return;
  const int j = (int)(t % J);
  const int pr = (int)((t / J) % P);
  const int f = (int)(t / ((long long)J * P));
  const int a = pr / (V - 1);
  int b = pr % (V - 1);
  if (b >= a) ++b;
  const double* F = fmat + (((size_t)subj[f] * V + a) * V + b) * 9;
  const size_t ra = ((size_t)f * V + a) * J + j, rb = ((size_t)f * V + b) * J + j;
  const double xa = (double)xy[2 * ra], ya = (double)xy[2 * ra + 1];
  const double xb = (double)xy[2 * rb], yb = (double)xy[2 * rb + 1];
  const double t0 = fma(yb, F[3], xb * F[0]) + F[6];
  const double t1 = fma(yb, F[4], xb * F[1]) + F[7];
  const double t2 = fma(yb, F[5], xb * F[2]) + F[8];
  const double r = (t0 * xa + t1 * ya) + t2;
  double s = r > 0.0 ? gscale : (r < 0.0 ? -gscale : 0.0);   // d|r|/dr, 0 at r == 0 like torch.abs
  if (w != nullptr) s *= (double)w[rb] * (double)w[ra];
  if (s == 0.0) return;
  // d r / d x_a = (t0, t1);  d r / d x_b = (F00 xa + F01 ya + F02, F10 xa + F11 ya + F12)
  atomicAdd(grad_xy + 2 * ra, s * t0);
  atomicAdd(grad_xy + 2 * ra + 1, s * t1);
  atomicAdd(grad_xy + 2 * rb, s * ((F[0] * xa + F[1] * ya) + F[2]));
  atomicAdd(grad_xy + 2 * rb + 1, s * ((F[3] * xa + F[4] * ya) + F[5]));
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_softargmax_fwd(const float* hm, int N, int J, int H, int W, float beta, float* out_xy,
                                    float* out_stats, void* stream) {
  PB_REQUIRE(N >= 0 && J >= 1 && H >= 1 && W >= 1, "bad shape");
  if (N == 0) return PB200_OK;
  PB_REQUIRE(hm && out_xy && out_stats, "null pointer");
  const long long maps = (long long)N * J;
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  long long blocks = (maps + 7) / 8;
  if (blocks > (long long)sm * 8) blocks = (long long)sm * 8;
  const int vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(hm) & 15u) == 0);
  softargmax_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hm, maps, H, W, beta, vec, out_xy, out_stats);
  PB_LAUNCH_CHECK("softargmax_fwd_kernel");
  return PB200_OK;
}

extern "C" int pb200_softargmax_bwd(const float* hm, const float* stats, const float* xy, const float* grad_xy,
                                    int N, int J, int H, int W, float beta, float* grad_hm, void* stream) {
  PB_REQUIRE(N >= 0 && J >= 1 && H >= 1 && W >= 1, "bad shape");
  if (N == 0) return PB200_OK;
  PB_REQUIRE(hm && stats && xy && grad_xy && grad_hm, "null pointer");
  const long long maps = (long long)N * J;
  const int sm = cached_sm_count();
  if (sm <= 0) return PB200_ERR_CUDA;
  long long blocks = (maps + 7) / 8;
  if (blocks > (long long)sm * 8) blocks = (long long)sm * 8;
  const int vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(hm) & 15u) == 0) &&
                  ((reinterpret_cast<uintptr_t>(grad_hm) & 15u) == 0);
  softargmax_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hm, stats, xy, grad_xy, maps, H, W,
                                                                         beta, vec, grad_hm);
  PB_LAUNCH_CHECK("softargmax_bwd_kernel");
  return PB200_OK;
}

extern "C" int pb200_epipolar_grad(const double* fmat, const int32_t* subj_index, const void* xy, int xy_dtype,
                                   const void* weight, int w_dtype, int B, int V, int J, double gscale,
                                   double* grad_xy, void* stream) {
  PB_REQUIRE(fmat && subj_index && xy && grad_xy, "null pointer");
  PB_REQUIRE(B >= 0 && J >= 1 && V >= 2 && V <= PB200_MAX_VIEWS, "bad shape B=%d V=%d J=%d", B, V, J);
  PB_REQUIRE((xy_dtype | 1) == 1 && (w_dtype | 1) == 1, "dtype tags must be PB200_F32/PB200_F64");
  const long long n = (long long)B * V * (V - 1) * J;
  if (n == 0) return PB200_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
#define PB_EPG(T, TW) \
  epipolar_grad_kernel<T, TW><<<blocks, 256, 0, s>>>(fmat, subj_index, (const T*)xy, (const TW*)weight, B, V, J, gscale, grad_xy)
  if (xy_dtype == PB200_F32) { if (w_dtype == PB200_F32) PB_EPG(float, float); else PB_EPG(float, double); }
  else { if (w_dtype == PB200_F32) PB_EPG(double, float); else PB_EPG(double, double); }
#undef PB_EPG
  PB_LAUNCH_CHECK("epipolar_grad_kernel");
  return PB200_OK;
}
