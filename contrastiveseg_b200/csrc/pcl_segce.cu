// Fused segmentation cross-entropy of ContrastCELoss.forward (the next row after the contrast path):
// replaces  F.interpolate(seg, (Himg,Wimg), bilinear, align_corners=True)  +  nn.CrossEntropyLoss(weight, ignore_index,
// reduction='mean')  (lib/loss/loss_contrast.py:180-181, lib/loss/loss_helper.py:169-212) without materialising the
// (B,K,Himg,Wimg) up-sampled logits (318 MB at B=8, K=19, 512x1024) or its log-softmax / gradient copies.
//   k_segce_fwd  one thread per label pixel: 4-neighbour bilinear logits for all K classes (two sweeps: max, then
//                sum of exp), NLL of the target class, log-sum-exp kept per pixel (B*Himg*Wimg floats) for backward;
//                per-CTA partial sums -> k_segce_finalize (fixed order: bit-reproducible)
//   k_segce_bwd  one thread per logit seg[b,c,y,x]: gathers over the label pixels in its bilinear footprint
//                (weight * (softmax - onehot)), no atomics: deterministic
// HBM: reads seg (B*K*h*w*4) + labels (B*Himg*Wimg*8), writes lse (B*Himg*Wimg*4); backward re-reads them and writes
// dseg.  Everything else stays in L1/L2 (a 4x4 block of label pixels shares its 4 source logits).
#include "pcl_common.cuh"
#include <math_constants.h>

namespace pcl {

struct SegCeArgs {
  const float* seg; const int64_t* target; const float* weight;
  int B, K, h, w, H, W, ignore_index;
  float ry, rx;                 // (h-1)/(H-1), (w-1)/(W-1)   (align_corners=True)
};

__device__ __forceinline__ void src_coord(int o, float r, int in_size, int& i0, int& i1, float& l1) {
  const float s = r * (float)o;                       // ATen area_pixel_compute_source_index, align_corners=True
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

constexpr int CE_THREADS = 256;

__global__ void __launch_bounds__(CE_THREADS)
k_segce_fwd(SegCeArgs a, float* __restrict__ lse_out, float* __restrict__ part_nll, float* __restrict__ part_w) {
  __shared__ float s_n[CE_THREADS], s_w[CE_THREADS];
  const int64_t HWo = (int64_t)a.H * a.W;
  const int64_t pix = (int64_t)blockIdx.x * CE_THREADS + threadIdx.x;
  const int b = blockIdx.y;
  float nll = 0.f, wt = 0.f;
  if (pix < HWo) {
    const int y = (int)(pix / a.W), x = (int)(pix - (int64_t)y * a.W);
    int y0, y1, x0, x1; float ly, lx;
    src_coord(y, a.ry, a.h, y0, y1, ly);
    src_coord(x, a.rx, a.w, x0, x1, lx);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const int64_t hw = (int64_t)a.h * a.w;
    const float* sb = a.seg + (int64_t)b * a.K * hw;
    const int64_t o00 = (int64_t)y0 * a.w + x0, o01 = (int64_t)y0 * a.w + x1, o10 = (int64_t)y1 * a.w + x0,
                  o11 = (int64_t)y1 * a.w + x1;
    float m = -CUDART_INF_F;
    for (int c = 0; c < a.K; ++c) {
      const float* p = sb + (int64_t)c * hw;
      const float v = w00 * p[o00] + w01 * p[o01] + w10 * p[o10] + w11 * p[o11];
      m = fmaxf(m, v);
    }
    float se = 0.f;
    const int64_t t = a.target[(int64_t)b * HWo + pix];
    float vt = 0.f;
    for (int c = 0; c < a.K; ++c) {
      const float* p = sb + (int64_t)c * hw;
      const float v = w00 * p[o00] + w01 * p[o01] + w10 * p[o10] + w11 * p[o11];
      se += expf(v - m);
      if (c == (int)t) vt = v;
    }
    const float lse = m + logf(se);
    lse_out[(int64_t)b * HWo + pix] = lse;
    if (t != (int64_t)a.ignore_index && t >= 0 && t < a.K) {
      const float wc = a.weight ? a.weight[t] : 1.f;
      nll = wc * (lse - vt);
      wt = wc;
    }
  }
  s_n[threadIdx.x] = nll; s_w[threadIdx.x] = wt;
  __syncthreads();
  for (int o = CE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s_n[threadIdx.x] += s_n[threadIdx.x + o]; s_w[threadIdx.x] += s_w[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    part_nll[blk] = s_n[0];
    part_w[blk] = s_w[0];
  }
}

__global__ void __launch_bounds__(1024)
k_segce_finalize(const float* __restrict__ part_nll, const float* __restrict__ part_w, int64_t n, float* __restrict__ out) {
  __shared__ double s_n[1024], s_w[1024];
  double an = 0.0, aw = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) { an += part_nll[i]; aw += part_w[i]; }
  s_n[threadIdx.x] = an; s_w[threadIdx.x] = aw;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s_n[threadIdx.x] += s_n[threadIdx.x + o]; s_w[threadIdx.x] += s_w[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = (float)(s_n[0] / s_w[0]);          // 'mean' reduction: sum(w_t * nll) / sum(w_t) over non-ignored pixels
    out[1] = (float)s_w[0];
  }
}

// Backward, tile formulation: one CTA = a 32 x 32 tile of label pixels of one image.
//   A  thread per label pixel: G[c] = coef * w_t * (softmax_c - [c == t]) for all K classes -> shared memory
//   B  separable transpose of the bilinear map, x direction:  Hx[c][row][xs] = sum_x wx(x, xs) * G[c][row][x]
//   C  y direction + flush: dseg[c][ys][xs] += sum_row wy(row, ys) * Hx[c][row][xs]   (global atomics only on the
//      <= 10 x 10 source cells of the tile; a cell is shared by at most 4 tiles)
// Each (pixel, class) softmax is evaluated once (the per-logit gather needs 16x as many).
constexpr int BT = 32;                 // tile edge (label pixels)
constexpr int SRC_MAX = 34;            // source cells per tile edge (scale >= 1: up to 32 + 2)

__global__ void __launch_bounds__(BT * BT)
k_segce_bwd_tile(SegCeArgs a, const float* __restrict__ lse, const float* __restrict__ fin,
                 const float* __restrict__ grad_out, float* __restrict__ dseg, int kc) {
  extern __shared__ float sm[];
  float* G = sm;                                   // [kc][BT][BT]
  float* Hx = G + (size_t)kc * BT * BT;            // [kc][BT][SRC_MAX]
  __shared__ int s_y0[BT], s_y1[BT], s_x0[BT], s_x1[BT];
  __shared__ float s_ly[BT], s_lx[BT];
  const int tx = threadIdx.x & (BT - 1), ty = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * BT, Y0 = blockIdx.y * BT;
  const int x = X0 + tx, y = Y0 + ty;
  const int64_t hw = (int64_t)a.h * a.w, HWo = (int64_t)a.H * a.W;
  if (threadIdx.x < BT) {
    int i0, i1; float l;
    src_coord(min(Y0 + (int)threadIdx.x, a.H - 1), a.ry, a.h, i0, i1, l);
    s_y0[threadIdx.x] = i0; s_y1[threadIdx.x] = i1; s_ly[threadIdx.x] = l;
    src_coord(min(X0 + (int)threadIdx.x, a.W - 1), a.rx, a.w, i0, i1, l);
    s_x0[threadIdx.x] = i0; s_x1[threadIdx.x] = i1; s_lx[threadIdx.x] = l;
  }
  __syncthreads();
  const int ys_base = s_y0[0], xs_base = s_x0[0];
  const int ny = min(Y0 + BT, a.H) - Y0, nx = min(X0 + BT, a.W) - X0;          // valid rows / cols of the tile
  const int ys_n = s_y1[ny - 1] - ys_base + 1, xs_n = s_x1[nx - 1] - xs_base + 1;
  const bool inside = x < a.W && y < a.H;
  const float scale = (grad_out ? grad_out[0] : 1.f) / fin[1];
  int64_t t = -1;
  float lse_p = 0.f, coef = 0.f;
  if (inside) {
    t = a.target[(int64_t)b * HWo + (int64_t)y * a.W + x];
    if (t != (int64_t)a.ignore_index && t >= 0 && t < a.K) {
      coef = scale * (a.weight ? a.weight[t] : 1.f);
      lse_p = lse[(int64_t)b * HWo + (int64_t)y * a.W + x];
    } else {
      t = -1;
    }
  }
  const int y0 = s_y0[ty], y1 = s_y1[ty], x0 = s_x0[tx], x1 = s_x1[tx];
  const float ly = s_ly[ty], lx = s_lx[tx];
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const int64_t o00 = (int64_t)y0 * a.w + x0, o01 = (int64_t)y0 * a.w + x1, o10 = (int64_t)y1 * a.w + x0,
                o11 = (int64_t)y1 * a.w + x1;
  for (int c0 = 0; c0 < a.K; c0 += kc) {
    const int cn = min(kc, a.K - c0);
    // ---- A: gradient w.r.t. the (virtual) up-sampled logits ----
    for (int cc = 0; cc < cn; ++cc) {
      float g = 0.f;
      if (t >= 0) {
        const float* p = a.seg + ((int64_t)b * a.K + c0 + cc) * hw;
        const float v = w00 * p[o00] + w01 * p[o01] + w10 * p[o10] + w11 * p[o11];
        g = coef * (expf(v - lse_p) - ((c0 + cc) == (int)t ? 1.f : 0.f));
      }
      G[((size_t)cc * BT + ty) * BT + tx] = g;
    }
    __syncthreads();
    // ---- B: x direction ----
    for (int i = threadIdx.x; i < cn * BT * xs_n; i += BT * BT) {
      const int xs = i % xs_n, row = (i / xs_n) % BT, cc = i / (xs_n * BT);
      const int xg = xs_base + xs;
      const float* gr = G + ((size_t)cc * BT + row) * BT;
      float acc = 0.f;
      for (int k = 0; k < nx; ++k) {
        float wgt = 0.f;
        if (s_x0[k] == xg) wgt += 1.f - s_lx[k];
        if (s_x1[k] == xg) wgt += s_lx[k];
        acc += wgt * gr[k];
      }
      Hx[((size_t)cc * BT + row) * SRC_MAX + xs] = acc;
    }
    __syncthreads();
    // ---- C: y direction, flush ----
    for (int i = threadIdx.x; i < cn * ys_n * xs_n; i += BT * BT) {
      const int xs = i % xs_n, ysl = (i / xs_n) % ys_n, cc = i / (xs_n * ys_n);
      const int yg = ys_base + ysl;
      float acc = 0.f;
      for (int k = 0; k < ny; ++k) {
        float wgt = 0.f;
        if (s_y0[k] == yg) wgt += 1.f - s_ly[k];
        if (s_y1[k] == yg) wgt += s_ly[k];
        acc += wgt * Hx[((size_t)cc * BT + k) * SRC_MAX + xs];
      }
      if (acc != 0.f) atomicAdd(&dseg[((int64_t)b * a.K + c0 + cc) * hw + (int64_t)yg * a.w + xs_base + xs], acc);
    }
    __syncthreads();
  }
}

// General backward (any scale, e.g. label grid smaller than the logits): one thread per logit gathers over the label
// pixels in its bilinear footprint.  Deterministic, but evaluates every (pixel, class) softmax up to 16 times; the tile
// kernel above is used whenever the label grid is at least as large as the logit grid (the reference's case).
__device__ __forceinline__ void footprint(int s, float r, int out_size, int& lo, int& hi) {
  if (r <= 0.f) { lo = 0; hi = out_size - 1; return; }
  lo = (int)floorf(((float)s - 1.f) / r) - 1;
  hi = (int)ceilf(((float)s + 1.f) / r) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

__global__ void __launch_bounds__(CE_THREADS)
k_segce_bwd_gather(SegCeArgs a, const float* __restrict__ lse, const float* __restrict__ fin,
                   const float* __restrict__ grad_out, float* __restrict__ dseg) {
  const int64_t hw = (int64_t)a.h * a.w;
  const int64_t e = (int64_t)blockIdx.x * CE_THREADS + threadIdx.x;
  const int bc = blockIdx.y;
  if (e >= hw) return;
  const int b = bc / a.K, c = bc - b * a.K;
  const int ys = (int)(e / a.w), xs = (int)(e - (int64_t)ys * a.w);
  const int64_t HWo = (int64_t)a.H * a.W;
  const float scale = (grad_out ? grad_out[0] : 1.f) / fin[1];
  const float* p = a.seg + ((int64_t)b * a.K + c) * hw;
  int ylo, yhi, xlo, xhi;
  footprint(ys, a.ry, a.H, ylo, yhi);
  footprint(xs, a.rx, a.W, xlo, xhi);
  float acc = 0.f;
  for (int y = ylo; y <= yhi; ++y) {
    int y0, y1; float ly;
    src_coord(y, a.ry, a.h, y0, y1, ly);
    float wy = 0.f;
    if (y0 == ys) wy += 1.f - ly;
    if (y1 == ys) wy += ly;
    if (wy == 0.f) continue;
    for (int x = xlo; x <= xhi; ++x) {
      int x0, x1; float lx;
      src_coord(x, a.rx, a.w, x0, x1, lx);
      float wx = 0.f;
      if (x0 == xs) wx += 1.f - lx;
      if (x1 == xs) wx += lx;
      if (wx == 0.f) continue;
      const int64_t pix = (int64_t)y * a.W + x;
      const int64_t t = a.target[(int64_t)b * HWo + pix];
      if (t == (int64_t)a.ignore_index || t < 0 || t >= a.K) continue;
      const float v = (1.f - ly) * (1.f - lx) * p[(int64_t)y0 * a.w + x0] + (1.f - ly) * lx * p[(int64_t)y0 * a.w + x1] +
                      ly * (1.f - lx) * p[(int64_t)y1 * a.w + x0] + ly * lx * p[(int64_t)y1 * a.w + x1];
      const float prob = expf(v - lse[(int64_t)b * HWo + pix]);
      const float wc = a.weight ? a.weight[t] : 1.f;
      acc += wy * wx * wc * (prob - (c == (int)t ? 1.f : 0.f));
    }
  }
  dseg[((int64_t)b * a.K + c) * hw + e] = acc * scale;
}

}  // namespace pcl

using namespace pcl;

static int make_ce(const float* seg, const int64_t* target, const float* weight, int B, int K, int h, int w, int H, int W,
                   int ignore_index, SegCeArgs* a) {
  if (!seg || !target || B <= 0 || K <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return PCL_ERR_ARG;
  a->seg = seg; a->target = target; a->weight = weight;
  a->B = B; a->K = K; a->h = h; a->w = w; a->H = H; a->W = W; a->ignore_index = ignore_index;
  a->ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  a->rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  return PCL_OK;
}

extern "C" int64_t pcl_seg_ce_scratch_floats(int32_t B, int32_t H, int32_t W) {
  const int64_t blocks = ceil_div64((int64_t)H * W, CE_THREADS) * B;
  return (int64_t)B * H * W + 2 * blocks + 2;            // lse | partial nll | partial weight | (loss, weight sum)
}

extern "C" int pcl_seg_ce_fwd(const float* seg, const int64_t* target, const float* class_weight, int32_t B, int32_t K,
                              int32_t h, int32_t w, int32_t H, int32_t W, int32_t ignore_index, float* scratch, float* loss,
                              void* stream) {
  SegCeArgs a;
  int st = make_ce(seg, target, class_weight, B, K, h, w, H, W, ignore_index, &a);
  if (st != PCL_OK) return st;
  PCL_REQUIRE(scratch && loss);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t HWo = (int64_t)H * W;
  const int64_t bx = ceil_div64(HWo, CE_THREADS), blocks = bx * B;
  float* lse = scratch;
  float* pn = scratch + (int64_t)B * HWo;
  float* pw = pn + blocks;
  float* fin = pw + blocks;
  k_segce_fwd<<<dim3((unsigned)bx, B), CE_THREADS, 0, s>>>(a, lse, pn, pw);
  PCL_LAUNCH_CHECK();
  k_segce_finalize<<<1, 1024, 0, s>>>(pn, pw, blocks, fin);
  PCL_LAUNCH_CHECK();
  PCL_CUDA(cudaMemcpyAsync(loss, fin, sizeof(float), cudaMemcpyDeviceToDevice, s));
  return PCL_OK;
}

extern "C" int pcl_seg_ce_bwd(const float* seg, const int64_t* target, const float* class_weight, int32_t B, int32_t K,
                              int32_t h, int32_t w, int32_t H, int32_t W, int32_t ignore_index, const float* scratch,
                              const float* grad_loss, float* dseg, void* stream) {
  SegCeArgs a;
  int st = make_ce(seg, target, class_weight, B, K, h, w, H, W, ignore_index, &a);
  if (st != PCL_OK) return st;
  PCL_REQUIRE(scratch && dseg);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t HWo = (int64_t)H * W;
  const int64_t blocks = ceil_div64(HWo, CE_THREADS) * B;
  const float* lse = scratch;
  const float* fin = scratch + (int64_t)B * HWo + 2 * blocks;
  const int64_t hw = (int64_t)h * w;
  if (H < h || W < w) {                                     // down-sampling: general gather kernel
    k_segce_bwd_gather<<<dim3((unsigned)ceil_div64(hw, CE_THREADS), B * K), CE_THREADS, 0, s>>>(a, lse, fin, grad_loss, dseg);
    PCL_LAUNCH_CHECK();
    return PCL_OK;
  }
  PCL_CUDA(cudaMemsetAsync(dseg, 0, (size_t)B * K * hw * sizeof(float), s));
  // classes per pass so that G (kc*32*32) + Hx (kc*32*34) floats fit 2 CTAs per SM
  int kc = K < 12 ? K : 12;
  const size_t smem = (size_t)kc * (BT * BT + BT * SRC_MAX) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    PCL_CUDA(cudaFuncSetAttribute(k_segce_bwd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * (BT * BT + BT * SRC_MAX) * 4));
    attr_done = true;
  }
  k_segce_bwd_tile<<<dim3((unsigned)ceil_div(W, BT), (unsigned)ceil_div(H, BT), B), BT * BT, smem, s>>>(a, lse, fin, grad_loss,
                                                                                                     dseg, kc);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}
