// Projection-head L2 normalise over the channel dim of an NCHW tensor.
// Replaces lib/models/modules/projection.py:24  F.normalize(x, p=2, dim=1)  and its autograd backward.
// HBM-bound: forward reads x once and writes y once (2 * B*D*HW*4 bytes); backward reads x and gy once and
// writes gx once.  One CTA covers 32 consecutive pixels x all D channels: lanes run along the pixel
// dimension (coalesced 128 B rows), the 8 warps split the channels and keep their values in registers.
#include "pcl_common.cuh"

namespace pcl {

constexpr int NW = 8;            // warps per CTA (channel groups)
constexpr int VMAX = 32;         // register-resident channels per thread (D <= NW * VMAX = 256)

template <bool BWD>
__global__ void __launch_bounds__(NW * 32)
k_l2norm(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ out, int D, int64_t HW,
         int64_t tiles_per_image) {
  __shared__ float s_ss[NW][32];
  __shared__ float s_dot[NW][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t b = blockIdx.x / tiles_per_image;
  const int64_t p = (blockIdx.x - b * tiles_per_image) * 32 + lane;
  const bool ok = p < HW;
  const int64_t base = b * (int64_t)D * HW + p;
  float xv[VMAX], gv[VMAX];
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int d = g + i * NW;
    xv[i] = 0.f; gv[i] = 0.f;
    if (ok && d < D) {
      xv[i] = x[base + (int64_t)d * HW];
      if (BWD) gv[i] = gy[base + (int64_t)d * HW];
    }
    ss = fmaf(xv[i], xv[i], ss);
    if (BWD) dot = fmaf(xv[i], gv[i], dot);
  }
  // channels beyond NW*VMAX (D > 256): streamed, re-read in the second pass
  for (int d = g + VMAX * NW; d < D; d += NW) {
    if (ok) {
      float v = x[base + (int64_t)d * HW];
      ss = fmaf(v, v, ss);
      if (BWD) dot = fmaf(v, gy[base + (int64_t)d * HW], dot);
    }
  }
  s_ss[g][lane] = ss;
  if (BWD) s_dot[g][lane] = dot;
  __syncthreads();
  float tss = 0.f, tdot = 0.f;
#pragma unroll
  for (int k = 0; k < NW; ++k) { tss += s_ss[k][lane]; if (BWD) tdot += s_dot[k][lane]; }
  const float nrm = sqrtf(tss);
  const bool clamped = nrm < 1e-12f;
  const float inv = 1.f / fmaxf(nrm, 1e-12f);
  // d/dx [x / max(|x|, eps)]:  g/|x| - x (x.g)/|x|^3   (denominator constant when clamped)
  const float coef = (BWD && !clamped) ? tdot * inv * inv * inv : 0.f;
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int d = g + i * NW;
    if (d < D) out[base + (int64_t)d * HW] = BWD ? (gv[i] * inv - xv[i] * coef) : xv[i] * inv;
  }
  for (int d = g + VMAX * NW; d < D; d += NW) {
    float v = x[base + (int64_t)d * HW];
    out[base + (int64_t)d * HW] = BWD ? (gy[base + (int64_t)d * HW] * inv - v * coef) : v * inv;
  }
}

}  // namespace pcl

using namespace pcl;

static int launch_norm(bool bwd, const float* x, const float* gy, float* out, int32_t B, int32_t D, int64_t HW,
                       void* stream) {
  PCL_REQUIRE(x && out && B > 0 && D > 0 && HW > 0);
  if (bwd) PCL_REQUIRE(gy);
  const int64_t tiles = ceil_div64(HW, 32);
  const int64_t blocks = tiles * B;
  PCL_REQUIRE(blocks < (1ll << 31));
  cudaStream_t s = (cudaStream_t)stream;
  if (bwd) k_l2norm<true><<<(unsigned)blocks, NW * 32, 0, s>>>(x, gy, out, D, HW, tiles);
  else     k_l2norm<false><<<(unsigned)blocks, NW * 32, 0, s>>>(x, nullptr, out, D, HW, tiles);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

extern "C" int pcl_l2norm_fwd(const float* x, float* y, int32_t B, int32_t D, int64_t HW, void* stream) {
  return launch_norm(false, x, nullptr, y, B, D, HW, stream);
}

extern "C" int pcl_l2norm_bwd(const float* x, const float* gy, float* gx, int32_t B, int32_t D, int64_t HW,
                              void* stream) {
  return launch_norm(true, x, gy, gx, B, D, HW, stream);
}
