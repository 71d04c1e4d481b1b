// Fused segmentation cross-entropy of ContrastCELoss.forward (the next row after the contrast path):
// replaces  F.interpolate(seg, (Himg,Wimg), bilinear, align_corners=True)  +  nn.CrossEntropyLoss(weight, ignore_index,
// reduction='mean')  (lib/loss/loss_contrast.py:180-181, lib/loss/loss_helper.py:169-212) without materialising the
// (B,K,Himg,Wimg) up-sampled logits (318 MB at B=8, K=19, 512x1024) or its log-softmax / gradient copies.
//   k_segce_fwd  one thread per label pixel: 4-neighbour bilinear logits for all K classes (two sweeps: max, then
//                sum of exp), NLL of the target class, log-sum-exp kept per pixel (B*Himg*Wimg floats) for backward;
//                per-CTA partial sums -> k_segce_finalize (fixed order: bit-reproducible)
//   k_segce_bwd  one thread per logit seg[b,c,y,x]: gathers over the label pixels in its bilinear footprint
//                (weight * (softmax - onehot)), no atomics: deterministic — the gather kernel, used for down-sampling
//                geometries; the default up-sampling geometry runs k_segce_bwd_tile, which accumulates the source
//                cells shared by up to 4 neighbouring tiles with fp32 global atomics: its dseg is reproducible only to
//                the last ulp or two (summation order of <= 4 addends), which the parity tolerance (1e-5 max|g|) covers
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
constexpr int BT = 32;                 // tile width  (label pixels)
constexpr int BTY = 8;                 // tile height: 32 x 8 = 256 threads per CTA -> several CTAs per SM hide the
                                       // load latency of the bulk-synchronous phases (a 32 x 32 tile ran 1 CTA/SM)
constexpr int KC_FWD_MAX = 32;         // classes staged per pass in the forward (one pass for K <= 32: one exposed load latency)

// Source patch geometry of a tile (identical for every thread of the CTA).
struct TileGeo {
  int ys_base, xs_base, ys_n, xs_n, ny, nx;
};

// Forward, tile formulation: the <= (32*ry+2) x (32*rx+2) source logits of the tile are staged in shared memory per class
// chunk; every thread (one label pixel) runs an online log-sum-exp over the classes from shared memory.
__global__ void __launch_bounds__(BT * BTY)
k_segce_fwd(SegCeArgs a, float* __restrict__ lse_out, float* __restrict__ part_nll, float* __restrict__ part_w, int pitch,
            int kc) {
  extern __shared__ float sm[];                    // [kc][pitch] source patch
  __shared__ float s_n[BT * BTY / 32], s_w[BT * BTY / 32];
  __shared__ int s_y0[BTY], s_y1[BTY], s_x0[BT], s_x1[BT];
  __shared__ float s_ly[BTY], s_lx[BT];
  const int tx = threadIdx.x & (BT - 1), ty = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * BT, Y0 = blockIdx.y * BTY;
  const int x = X0 + tx, y = Y0 + ty;
  const int64_t hw = (int64_t)a.h * a.w, HWo = (int64_t)a.H * a.W;
  if (threadIdx.x < BT) {
    int i0, i1; float l;
    if (threadIdx.x < BTY) {
      src_coord(min(Y0 + (int)threadIdx.x, a.H - 1), a.ry, a.h, i0, i1, l);
      s_y0[threadIdx.x] = i0; s_y1[threadIdx.x] = i1; s_ly[threadIdx.x] = l;
    }
    src_coord(min(X0 + (int)threadIdx.x, a.W - 1), a.rx, a.w, i0, i1, l);
    s_x0[threadIdx.x] = i0; s_x1[threadIdx.x] = i1; s_lx[threadIdx.x] = l;
  }
  __syncthreads();
  const int ys_base = s_y0[0], xs_base = s_x0[0];
  const int ny = min(Y0 + BTY, a.H) - Y0, nx = min(X0 + BT, a.W) - X0;
  const int ys_n = s_y1[ny - 1] - ys_base + 1, xs_n = s_x1[nx - 1] - xs_base + 1;
  const bool inside = x < a.W && y < a.H;
  const int ly0 = s_y0[ty] - ys_base, ly1 = s_y1[ty] - ys_base, lx0 = s_x0[tx] - xs_base, lx1 = s_x1[tx] - xs_base;
  const float ly = s_ly[ty], lx = s_lx[tx];
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const int i00 = ly0 * xs_n + lx0, i01 = ly0 * xs_n + lx1, i10 = ly1 * xs_n + lx0, i11 = ly1 * xs_n + lx1;
  int64_t t = -1;
  if (inside) t = a.target[(int64_t)b * HWo + (int64_t)y * a.W + x];
  float m = -CUDART_INF_F, se = 0.f, vt = 0.f;
  for (int c0 = 0; c0 < a.K; c0 += kc) {
    const int cn = min(kc, a.K - c0);
    const int patch = ys_n * xs_n;
    for (int i = threadIdx.x; i < cn * patch; i += BT * BTY) {
      const int cc = i / patch, r = i - cc * patch;
      const int yy = r / xs_n, xx = r - yy * xs_n;
      sm[cc * pitch + r] = a.seg[((int64_t)b * a.K + c0 + cc) * hw + (int64_t)(ys_base + yy) * a.w + xs_base + xx];
    }
    __syncthreads();
    for (int cc = 0; cc < cn; ++cc) {
      const float* p = sm + cc * pitch;
      const float v = w00 * p[i00] + w01 * p[i01] + w10 * p[i10] + w11 * p[i11];
      if (v > m) { se = se * __expf(m - v) + 1.f; m = v; }     // online log-sum-exp
      else       { se += __expf(v - m); }
      if (c0 + cc == (int)t) vt = v;
    }
    __syncthreads();
  }
  float nll = 0.f, wt = 0.f;
  if (inside) {
    const float lse = m + __logf(se);
    lse_out[(int64_t)b * HWo + (int64_t)y * a.W + x] = lse;
    if (t != (int64_t)a.ignore_index && t >= 0 && t < a.K) {
      const float wc = a.weight ? a.weight[t] : 1.f;
      nll = wc * (lse - vt);
      wt = wc;
    }
  }
  nll = warp_sum(nll); wt = warp_sum(wt);
  if ((threadIdx.x & 31) == 0) { s_n[threadIdx.x >> 5] = nll; s_w[threadIdx.x >> 5] = wt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float an = 0.f, aw = 0.f;
    for (int i = 0; i < BT * BTY / 32; ++i) { an += s_n[i]; aw += s_w[i]; }      // fixed order
    const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    part_nll[blk] = an;
    part_w[blk] = aw;
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

// Backward, tile formulation: one CTA = a 32 x 32 tile of label pixels of one image, classes in chunks of kc.
//   A  thread per label pixel: G[c] = coef * w_t * (softmax_c - [c == t]) from the staged source patch -> shared memory
//   B  separable transpose of the bilinear map, x direction:  Hx[c][row][xs] = sum_x wx(x, xs) * G[c][row][x]
//   C  y direction + flush: dseg[c][ys][xs] += sum_row wy(row, ys) * Hx[c][row][xs]   (global atomics only on the
//      source cells of the tile; a cell is shared by at most 4 tiles)
// Each (pixel, class) softmax is evaluated once (the per-logit gather below needs 16x as many); B and C only visit the
// label columns / rows inside the footprint of their source cell.
constexpr int SRC_MAX = 34;            // source cells per tile edge (scale >= 1: up to 32 + 2)

__global__ void __launch_bounds__(BT * BTY)
k_segce_bwd_tile(SegCeArgs a, const float* __restrict__ lse, const float* __restrict__ fin,
                 const float* __restrict__ grad_out, float* __restrict__ dseg, int kc, int pitch) {
  extern __shared__ float sm[];
  float* G = sm;                                   // [kc][BTY][BT]
  float* Hx = G + (size_t)kc * BTY * BT;           // [kc][BTY][SRC_MAX]
  float* S = Hx + (size_t)kc * BTY * SRC_MAX;      // [kc][pitch] source patch
  __shared__ int s_y0[BTY], s_y1[BTY], s_x0[BT], s_x1[BT];
  __shared__ float s_ly[BTY], s_lx[BT];
  const int tx = threadIdx.x & (BT - 1), ty = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * BT, Y0 = blockIdx.y * BTY;
  const int x = X0 + tx, y = Y0 + ty;
  const int64_t hw = (int64_t)a.h * a.w, HWo = (int64_t)a.H * a.W;
  if (threadIdx.x < BT) {
    int i0, i1; float l;
    if (threadIdx.x < BTY) {
      src_coord(min(Y0 + (int)threadIdx.x, a.H - 1), a.ry, a.h, i0, i1, l);
      s_y0[threadIdx.x] = i0; s_y1[threadIdx.x] = i1; s_ly[threadIdx.x] = l;
    }
    src_coord(min(X0 + (int)threadIdx.x, a.W - 1), a.rx, a.w, i0, i1, l);
    s_x0[threadIdx.x] = i0; s_x1[threadIdx.x] = i1; s_lx[threadIdx.x] = l;
  }
  __syncthreads();
  const int ys_base = s_y0[0], xs_base = s_x0[0];
  const int ny = min(Y0 + BTY, a.H) - Y0, nx = min(X0 + BT, a.W) - X0;         // valid rows / cols of the tile
  const int ys_n = s_y1[ny - 1] - ys_base + 1, xs_n = s_x1[nx - 1] - xs_base + 1;
  const int patch = ys_n * xs_n;
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
  const int ly0 = s_y0[ty] - ys_base, ly1 = s_y1[ty] - ys_base, lx0 = s_x0[tx] - xs_base, lx1 = s_x1[tx] - xs_base;
  const float ly = s_ly[ty], lx = s_lx[tx];
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const int i00 = ly0 * xs_n + lx0, i01 = ly0 * xs_n + lx1, i10 = ly1 * xs_n + lx0, i11 = ly1 * xs_n + lx1;
  const float inv_rx = a.rx > 0.f ? 1.f / a.rx : 0.f, inv_ry = a.ry > 0.f ? 1.f / a.ry : 0.f;
  // The transposed bilinear map of the tile, once per CTA (it is the same for every class and row): weight of label
  // column k in source column xs (s_wx), of label row k in source row ys (s_wy), and the label range [klo, khi] inside
  // a source cell's footprint.  Phases B and C below were evaluating these (two compares, floorf / ceilf) per class, per
  // row and per k: 152 times each.  Row stride 33 / 9: different source cells of one k fall into different banks.
  __shared__ float s_wx[SRC_MAX * (BT + 1)], s_wy[SRC_MAX * (BTY + 1)];
  __shared__ short s_xlo[SRC_MAX], s_xhi[SRC_MAX], s_ylo[SRC_MAX], s_yhi[SRC_MAX];
  for (int i = threadIdx.x; i < xs_n * BT; i += BT * BTY) {
    const int xs = i / BT, k = i - xs * BT, xg = xs_base + xs;
    float wgt = 0.f;
    if (k < nx) {
      if (s_x0[k] == xg) wgt += 1.f - s_lx[k];
      if (s_x1[k] == xg) wgt += s_lx[k];
    }
    s_wx[xs * (BT + 1) + k] = wgt;
  }
  for (int i = threadIdx.x; i < ys_n * BTY; i += BT * BTY) {
    const int ysl = i / BTY, k = i - ysl * BTY, yg = ys_base + ysl;
    float wgt = 0.f;
    if (k < ny) {
      if (s_y0[k] == yg) wgt += 1.f - s_ly[k];
      if (s_y1[k] == yg) wgt += s_ly[k];
    }
    s_wy[ysl * (BTY + 1) + k] = wgt;
  }
  if (threadIdx.x < xs_n) {
    const int xg = xs_base + threadIdx.x;
    int klo = 0, khi = nx - 1;
    if (a.rx > 0.f) {
      klo = max(0, (int)floorf(((float)xg - 1.f) * inv_rx) - 1 - X0);
      khi = min(nx - 1, (int)ceilf(((float)xg + 1.f) * inv_rx) + 1 - X0);
    }
    s_xlo[threadIdx.x] = (short)klo; s_xhi[threadIdx.x] = (short)khi;
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + ys_n) {
    const int ysl = threadIdx.x - 64, yg = ys_base + ysl;
    int klo = 0, khi = ny - 1;
    if (a.ry > 0.f) {
      klo = max(0, (int)floorf(((float)yg - 1.f) * inv_ry) - 1 - Y0);
      khi = min(ny - 1, (int)ceilf(((float)yg + 1.f) * inv_ry) + 1 - Y0);
    }
    s_ylo[ysl] = (short)klo; s_yhi[ysl] = (short)khi;
  }
  __syncthreads();
  for (int c0 = 0; c0 < a.K; c0 += kc) {
    const int cn = min(kc, a.K - c0);
    for (int i = threadIdx.x; i < cn * patch; i += BT * BTY) {
      const int cc = i / patch, r = i - cc * patch;
      const int yy = r / xs_n, xx = r - yy * xs_n;
      S[cc * pitch + r] = a.seg[((int64_t)b * a.K + c0 + cc) * hw + (int64_t)(ys_base + yy) * a.w + xs_base + xx];
    }
    __syncthreads();
    // ---- A: gradient w.r.t. the (virtual) up-sampled logits ----
    for (int cc = 0; cc < cn; ++cc) {
      float g = 0.f;
      if (t >= 0) {
        const float* p = S + cc * pitch;
        const float v = w00 * p[i00] + w01 * p[i01] + w10 * p[i10] + w11 * p[i11];
        g = coef * (__expf(v - lse_p) - ((c0 + cc) == (int)t ? 1.f : 0.f));
      }
      G[((size_t)cc * BTY + ty) * BT + tx] = g;
    }
    __syncthreads();
    // ---- B: x direction (only the label columns whose footprint touches source column xg) ----
    for (int i = threadIdx.x; i < cn * BTY * xs_n; i += BT * BTY) {
      const int xs = i % xs_n, row = (i / xs_n) % BTY, cc = i / (xs_n * BTY);
      const int klo = s_xlo[xs], khi = s_xhi[xs];
      const float* gr = G + ((size_t)cc * BTY + row) * BT;
      const float* wr = s_wx + xs * (BT + 1);
      float acc = 0.f;
      for (int k = klo; k <= khi; ++k) acc += wr[k] * gr[k];
      Hx[((size_t)cc * BTY + row) * SRC_MAX + xs] = acc;
    }
    __syncthreads();
    // ---- C: y direction, flush ----
    for (int i = threadIdx.x; i < cn * ys_n * xs_n; i += BT * BTY) {
      const int xs = i % xs_n, ysl = (i / xs_n) % ys_n, cc = i / (xs_n * ys_n);
      const int yg = ys_base + ysl;
      const int klo = s_ylo[ysl], khi = s_yhi[ysl];
      const float* wr = s_wy + ysl * (BTY + 1);
      float acc = 0.f;
      for (int k = klo; k <= khi; ++k) acc += wr[k] * Hx[((size_t)cc * BTY + k) * SRC_MAX + xs];
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

static int patch_pitch(const SegCeArgs& a) {
  // upper bound of the source cells a 32 x BTY label tile touches (ys_n * xs_n), + 1 to de-phase the class planes
  const int py = (int)((float)(BTY - 1) * a.ry) + 3, px = (int)(31.f * a.rx) + 3;
  return (py < a.h + 1 ? py : a.h + 1) * (px < a.w + 1 ? px : a.w + 1) + 1;
}

extern "C" int64_t pcl_seg_ce_scratch_floats(int32_t B, int32_t H, int32_t W) {
  const int64_t blocks = (int64_t)ceil_div(H, BTY) * ceil_div(W, BT) * B;
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
  const int64_t blocks = (int64_t)ceil_div(H, BTY) * ceil_div(W, BT) * B;
  float* lse = scratch;
  float* pn = scratch + (int64_t)B * HWo;
  float* pw = pn + blocks;
  float* fin = pw + blocks;
  const int pitch = patch_pitch(a);
  int kc = K < KC_FWD_MAX ? K : KC_FWD_MAX;
  while (kc > 1 && (size_t)kc * pitch * sizeof(float) > 96 * 1024) --kc;
  const size_t smem = (size_t)kc * pitch * sizeof(float);
  if (smem > 200 * 1024) return PCL_ERR_UNSUPPORTED;
  PCL_SMEM_OPT_IN(k_segce_fwd, 200 * 1024);
  k_segce_fwd<<<dim3((unsigned)ceil_div(W, BT), (unsigned)ceil_div(H, BTY), B), BT * BTY, smem, s>>>(a, lse, pn, pw, pitch, kc);
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
  const int64_t blocks = (int64_t)ceil_div(H, BTY) * ceil_div(W, BT) * B;
  const float* lse = scratch;
  const float* fin = scratch + (int64_t)B * HWo + 2 * blocks;
  const int64_t hw = (int64_t)h * w;
  if (H < h || W < w) {                                     // down-sampling: general gather kernel
    k_segce_bwd_gather<<<dim3((unsigned)ceil_div64(hw, CE_THREADS), B * K), CE_THREADS, 0, s>>>(a, lse, fin, grad_loss, dseg);
    PCL_LAUNCH_CHECK();
    return PCL_OK;
  }
  PCL_CUDA(cudaMemsetAsync(dseg, 0, (size_t)B * K * hw * sizeof(float), s));
  // classes per pass: G (kc*32*32) + Hx (kc*32*34) + source patch (kc*pitch) floats, two CTAs per SM when possible
  const int pitch = patch_pitch(a);
  int kc = K < 24 ? K : 24;                  // one pass for K <= 24 (fewer bulk-synchronous phases per CTA)
  while (kc > 1 && (size_t)kc * (BTY * BT + BTY * SRC_MAX + pitch) * sizeof(float) > 200 * 1024) --kc;
  const size_t smem = (size_t)kc * (BTY * BT + BTY * SRC_MAX + pitch) * sizeof(float);
  if (smem > 200 * 1024) return PCL_ERR_UNSUPPORTED;
  PCL_SMEM_OPT_IN(k_segce_bwd_tile, 200 * 1024);
  k_segce_bwd_tile<<<dim3((unsigned)ceil_div(W, BT), (unsigned)ceil_div(H, BTY), B), BT * BTY, smem, s>>>(a, lse, fin, grad_loss,
                                                                                                     dseg, kc, pitch);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}
