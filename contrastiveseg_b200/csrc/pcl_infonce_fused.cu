// Self-contrast InfoNCE, forward AND backward in ONE launch, for the small-anchor regime of the reference configs
// (A = N <= 1024 anchors, D = 256: BASELINE configs[1], SURVEY §8 row a6 "S1: launch/latency-bound").
//
// Replaces lib/loss/loss_contrast.py:91-128 and its autograd backward.  The streaming sweeps of pcl_infonce_tc.cu need
// six launches for this shape (prep/fill, NEG, POS, finalize, backward, reduce) that together move < 30 MB and take
// ~85 us of pure latency; here the whole A x A logit matrix (<= 8 x 4 tiles of 128 x 256) lives in tensor memory:
//
//   CTA (r, c) = row tile r (128 anchors) x column tile c (256 anchors), <= 32 CTAs, all co-resident:
//     TMA      anchor tile r (64 KB) + anchor tile c (128 KB), SWIZZLE_128B
//     MMA1     S = A_r . A_c^T  (16 x tcgen05.mma M128 N256 K16) -> TMEM columns [0, 256), computed ONCE
//     phase 1  NEG: per row, sum over negatives of exp2(s k1 - m2)                    -> partial (c, row)
//     -- inter-CTA barrier (release/acquire counter in global memory; every live CTA is resident) --
//     phase 2  POS: S re-read from TMEM; log-prob sum, sum 1/(e + Neg), count         -> partial (c, row)
//     -- inter-CTA barrier --
//     phase 3  gradient tile H = G + G^T (closed form) from the SAME S in TMEM, bf16, written
//              K-major/128B-swizzled over the (no longer needed) anchor tile r
//     MMA2     dA_r (partial over c) = H . A_c   (A_c re-read MN-major from shared memory) -> TMEM columns [256, 512)
//              -> global partial (c, row, 256); the dense-gradient writer sums the <= 4 partials in fixed order
//   The last CTA to finish reduces the per-row-tile loss sums in fixed order (bit-reproducible) and re-arms the counters.
//
// No similarity recompute (the streaming backward spends half its MMAs on it), no A x N temporary, one launch.
// Row stabiliser: Cauchy-Schwarz bound m_i = |a_i| max|a| / T as in pcl_infonce_tc.cu (exact in real arithmetic).
#include "pcl_common.cuh"
#include "pcl_sweep.cuh"
#include "ptx_sm100.cuh"

namespace pcl {
namespace fused {

constexpr int BM = 128, BN = 256, BK = 64, NKB = 4, DDIM = 256;
constexpr int A_KB_BYTES = BM * BK * 2;          // 16 KB
constexpr int C_KB_BYTES = BN * BK * 2;          // 32 KB
constexpr int NUM_THREADS = 320;                 // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int EPI_THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr float LN2 = 0.6931471805599453f;

struct Args {
  const int32_t* acls; const int32_t* plan; const float* row_m2;
  int a_rows, a_pad, s_max;          // s_max: column tiles of a_pad (stride of the partial arrays)
  float k1, rs_scale;                // log2(e)/T, T/bT
  int nan_safe;
  float* partials;                   // [5][s_max][a_pad]: k = 0 tile-loss scratch, 1 neg, 2 possum, 3 s, 4 count
  float* rowstats;                   // [6][a_rows]
  float* loss;
  float* dpartials;                  // [s_max][a_pad][256]
  unsigned int* sync;                // [0] barrier 1, [1] barrier 2, [2] finished CTAs  (zero before the first launch)
  int phase_mask;                    // 7 = all phases in one launch; single bits: one phase per launch (emulation only)
};

struct Smem {
  uint8_t a[NKB * A_KB_BYTES];       // anchor tile r; re-used for the gradient tile H (same size: 128 x 256 bf16)
  uint8_t c[NKB * C_KB_BYTES];       // anchor tile c
  uint64_t ac_full, s_full, g_full, da_full;
  uint32_t tmem_base;
  float comb[3][2][BM];
  float red[BM];
  float4 colstat[BN];                // phase 3: per column j of tile c: {m2_j, Neg_j, c_j S_j, -c_j Neg_j}
  float2 colneg[BN];                 // {m2_j, c_j S_j}: all an all-negative chunk needs
  int collab[BN];                    // label of column j (-2: beyond the live anchors)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float ld_cg(const float* p) {
#ifdef PCL_EMULATION
  return *p;
#else
  return __ldcg(p);                  // written by other SMs during this launch: bypass the (non-coherent) L1
#endif
}

// Barrier among the `n_live` CTAs of this launch (all resident: <= 32 CTAs of one per SM).  Called by the 256 epilogue
// threads; the TMA / MMA warps are parked on their mbarriers meanwhile.
__device__ __forceinline__ void cta_group_barrier(unsigned int* ctr, unsigned int n_live) {
  __threadfence();                                        // this thread's partials are visible device-wide
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
  if (threadIdx.x == 64) {                                // first epilogue thread
#ifndef PCL_EMULATION
    atomicAdd(ctr, 1u);
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(ctr) < n_live) {
      __nanosleep(64);
      if (++spins > (1u << 22)) { printf("pcl: fused InfoNCE inter-CTA barrier timeout (%d,%d)\n", blockIdx.x, blockIdx.y); __trap(); }
    }
    __threadfence();
#endif
  }
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
}

// Scheduling note (measured with the in-graph timeline, profiles/r2_1*_timeline_*): next to the 268 MB zero-fill of the
// dense gradient every global access of this kernel (labels, TMA loads, the two partial exchanges, the barrier atomics)
// takes 4-8x longer — co-resident with the fill (144 registers: 3 warps + 2 fill warps fit a 16 K sub-partition) it needs
// 70 us instead of 35.  At 168 registers its CTAs cannot be placed next to fill CTAs, start when the first fill CTAs
// retire and run at full speed: the step is 5-10 us shorter that way, so the register count is left alone.
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_self_fused(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC, Args a) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int A = a.plan ? min(a.plan[PCL_PLAN_A], a.a_rows) : a.a_rows;
  const int c = blockIdx.x, r = blockIdx.y;
  const int R_live = (A + BM - 1) / BM, S_live = (A + BN - 1) / BN;
  const int64_t pstride = (int64_t)a.s_max * a.a_pad;
  if (c == 0 && r == 0 && (a.phase_mask & 4)) {
    // rows that no CTA owns (beyond the live row tiles) read as empty; no anchors at all: zero loss
    for (int i = R_live * BM + threadIdx.x; i < a.a_rows; i += blockDim.x)
      for (int k = 0; k < 6; ++k) a.rowstats[k * a.a_rows + i] = 0.f;
    if (A <= 0 && threadIdx.x == 0) *a.loss = 0.f;
  }
  if (A <= 0 || r >= R_live || c >= S_live) return;
  const unsigned int n_live = (unsigned int)(R_live * S_live);
  const int row0 = r * BM, col0 = c * BN;
  tl_begin(a.sync, PCL_TL_FUSED);
  // per-CTA phase stamps (diagnostics, when the timeline is on): words [64 + 16 * cta, +16) = 8 x uint64
  unsigned long long* stamps = (a.sync[7] != 0u) ? reinterpret_cast<unsigned long long*>(a.sync + 64) + 8 * (r * a.s_max + c) : nullptr;
#define PCL_STAMP(i) do { if (stamps != nullptr && threadIdx.x == 64) stamps[i] = pcl_globaltimer(); } while (0)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmC);
    ptx::mbar_init(&sm.ac_full, 1);
    ptx::mbar_init(&sm.s_full, 1);
    ptx::mbar_init(&sm.g_full, EPI_THREADS / 32);
    ptx::mbar_init(&sm.da_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(&sm.tmem_base);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  const uint32_t tmem_dA = tmem_base + BN;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&sm.ac_full, NKB * (A_KB_BYTES + C_KB_BYTES));
      for (int kb = 0; kb < NKB; ++kb) ptx::tma_load_2d(sm.a + kb * A_KB_BYTES, &tmA, &sm.ac_full, kb * BK, row0);
      for (int kb = 0; kb < NKB; ++kb) ptx::tma_load_2d(sm.c + kb * C_KB_BYTES, &tmC, &sm.ac_full, kb * BK, col0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = ptx::make_idesc_bf16(BM, BN, 0, 0);      // S = A_r . A_c^T   (both K-major)
      constexpr uint32_t idesc2 = ptx::make_idesc_bf16(BM, DDIM, 0, 1);    // dA += H . A_c     (B operand MN-major)
      const uint32_t a_base = ptx::smem_u32(sm.a), c_base = ptx::smem_u32(sm.c);
      ptx::mbar_wait(&sm.ac_full, 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = ptx::make_desc_kmajor_sw128(a_base + kb * A_KB_BYTES + k * 32);
          const uint64_t dc = ptx::make_desc_kmajor_sw128(c_base + kb * C_KB_BYTES + k * 32);
          ptx::mma_f16_ss(tmem_base, da, dc, idesc1, (kb | k) != 0 ? 1u : 0u);
        }
      }
      ptx::mma_commit(&sm.s_full);
      if (a.phase_mask & 4) {
        ptx::mbar_wait(&sm.g_full, 0);                    // gradient tile H is in shared memory (written over tile r)
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          // A operand: H, K-major with K = column index j: 4 blocks of 64 j (16 KB each), 32 B per 16-element K step
          const uint64_t dh = ptx::make_desc_kmajor_sw128(a_base + (k >> 2) * A_KB_BYTES + (k & 3) * 32);
          // B operand: anchor tile c read MN-major: N = feature d (4 chunks of 64 at C_KB_BYTES), K = j (16 rows = 2 KB)
          const uint64_t dc = ptx::make_desc_mnmajor_sw128(c_base + k * 2048, C_KB_BYTES, 1024);
          ptx::mma_f16_ss(tmem_dA, dh, dc, idesc2, k != 0 ? 1u : 0u);
        }
        ptx::mma_commit(&sm.da_full);
      }
    }
  } else {
    const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                     // which 128 columns of the 256-column tile
    const int r_in = quarter * 32 + lane;
    const int row = row0 + r_in;
    const bool valid = row < A;
    const int rcls = valid ? a.acls[row] : -1;
    const float m2 = valid ? a.row_m2[row] : 0.f;
    const int cbase = col0 + half * (BN / 2);
    int clab[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int cj = cbase + ch * 32 + lane;
      clab[ch] = cj < A ? a.acls[cj] : -2;
    }
    PCL_STAMP(0);
    ptx::mbar_wait(&sm.s_full, 0);
    ptx::tc_fence_after();
    PCL_STAMP(1);
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * (BN / 2);
    float* pn = a.partials + 1 * pstride;
    float* pp0 = a.partials + 2 * pstride;
    float* pp1 = a.partials + 3 * pstride;
    float* pp2 = a.partials + 4 * pstride;
    const int64_t po = (int64_t)c * a.a_pad + row;

    // ---------------- phase 1: negatives ----------------
    if (a.phase_mask & 1) {
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_row + ch * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int l0 = __shfl_sync(0xffffffffu, clab[ch], j), l1 = __shfl_sync(0xffffffffu, clab[ch], j + 1);
          const float e0 = ptx::ex2_approx(fmaf(__uint_as_float(v[j]), a.k1, -m2));
          const float e1 = ptx::ex2_approx(fmaf(__uint_as_float(v[j + 1]), a.k1, -m2));
          acc0 += (valid && l0 != -2 && l0 != rcls) ? e0 : 0.f;
          acc1 += (valid && l1 != -2 && l1 != rcls) ? e1 : 0.f;
        }
      }
      sm.comb[0][half][r_in] = acc0 + acc1;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (half == 0) pn[po] = sm.comb[0][0][r_in] + sm.comb[0][1][r_in];
      PCL_STAMP(2);
      if (a.phase_mask == 7) cta_group_barrier(&a.sync[0], n_live);
      PCL_STAMP(3);
    }

    // ---------------- phase 2: positives ----------------
    float neg_i = 1.f;
    if (a.phase_mask & 6) {
      float n = 0.f;
      if (valid)
        for (int cc = 0; cc < S_live; ++cc) n += ld_cg(pn + (int64_t)cc * a.a_pad + row);      // fixed order
      neg_i = valid ? n : 1.f;
    }
    if (a.phase_mask & 2) {
      float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        // Anchors are grouped by class, so a 32-column chunk holds a few classes and most (warp, chunk) pairs contain no
        // positive at all (19 classes: ~1 in 10 does): find out with a walk over the chunk's distinct labels (ballots)
        // and skip the tensor-memory load and the element loop otherwise.
        bool mine = false;
        {
          const int L = clab[ch];
          unsigned remaining = 0xffffffffu;
          while (remaining) {
            const int lab = __shfl_sync(0xffffffffu, L, __ffs(remaining) - 1);
            remaining &= ~__ballot_sync(0xffffffffu, L == lab);
            mine |= (lab == rcls);
          }
        }
        if (!__any_sync(0xffffffffu, mine && valid)) continue;
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_row + ch * 32, v);
        ptx::tmem_ld_wait();
        const int cb = cbase + ch * 32;
        // branch-free (a divergent branch around the three MUFU ops serialised their latencies: 9 us per CTA)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int lj = __shfl_sync(0xffffffffu, clab[ch], j);
          const float keep = (valid && lj == rcls && cb + j != row) ? 1.f : 0.f;
          const float x = fmaf(__uint_as_float(v[j]), a.k1, -m2);
          const float t = ptx::ex2_approx(x) + neg_i;
          q0 = fmaf(keep, x - ptx::lg2_approx(t), q0);
          q1 = fmaf(keep, ptx::rcp_approx(t), q1);
          q2 += keep;
        }
      }
      sm.comb[0][half][r_in] = q0 * LN2;
      sm.comb[1][half][r_in] = q1;
      sm.comb[2][half][r_in] = q2;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (half == 0) {
        pp0[po] = sm.comb[0][0][r_in] + sm.comb[0][1][r_in];
        pp1[po] = sm.comb[1][0][r_in] + sm.comb[1][1][r_in];
        pp2[po] = sm.comb[2][0][r_in] + sm.comb[2][1][r_in];
      }
      PCL_STAMP(4);
      if (a.phase_mask == 7) cta_group_barrier(&a.sync[1], n_live);
      PCL_STAMP(5);
    }

    // ---------------- phase 3: gradient tile, MMA2, dA partial ----------------
    if (a.phase_mask & 4) {
      // totals of this thread's row (fixed order over the column tiles)
      float ps = 0.f, s_i = 0.f, cnt = 0.f;
      if (valid)
        for (int cc = 0; cc < S_live; ++cc) {
          const int64_t o = (int64_t)cc * a.a_pad + row;
          ps += ld_cg(pp0 + o); s_i += ld_cg(pp1 + o); cnt += ld_cg(pp2 + o);
        }
      float c_i = valid ? a.rs_scale / ((float)A * cnt) : 0.f;
      if (a.nan_safe && !(cnt > 0.f)) c_i = 0.f;
      const float cs_i = c_i * s_i, cn_i = -c_i * neg_i;
      if (c == 0) {
        // row statistics (same layout as the streaming sweeps) and the row tile's loss sum, by the column-tile-0 CTA
        float rl = 0.f;
        if (valid) {
          rl = -a.rs_scale * ps / cnt;
          if (a.nan_safe && !(cnt > 0.f)) rl = 0.f;
        }
        if (half == 0) {
          if (row < a.a_rows) {
            a.rowstats[row] = valid ? m2 * LN2 : 0.f;
            a.rowstats[a.a_rows + row] = valid ? neg_i : 0.f;
            a.rowstats[2 * a.a_rows + row] = ps;
            a.rowstats[3 * a.a_rows + row] = s_i;
            a.rowstats[4 * a.a_rows + row] = cnt;
            a.rowstats[5 * a.a_rows + row] = rl;
          }
          sm.red[r_in] = rl;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        if (threadIdx.x == 64) {
          float t = 0.f;
          for (int i = 0; i < BM; ++i) t += sm.red[i];                    // fixed order
          a.partials[r] = t;                                              // k = 0 region: per-row-tile loss sums
        }
      }
      // statistics of the 256 columns of tile c (every column is an anchor too: H = G + G^T): one column per epilogue
      // thread, into shared memory (the element loop reads them as broadcasts: no register copies, no shuffles)
      {
        const int t = threadIdx.x - 64;                                    // 0..255
        const int cj = col0 + t;
        const bool cok = cj < A;
        float nj = 1.f, sj = 0.f, cntj = 1.f;
        if (cok) {
          nj = 0.f; cntj = 0.f;
          for (int cc = 0; cc < S_live; ++cc) {
            const int64_t o = (int64_t)cc * a.a_pad + cj;
            nj += ld_cg(pn + o); sj += ld_cg(pp1 + o); cntj += ld_cg(pp2 + o);
          }
        }
        float c_j = cok ? a.rs_scale / ((float)A * cntj) : 0.f;
        if (a.nan_safe && !(cntj > 0.f)) c_j = 0.f;
        const float m2j = cok ? a.row_m2[cj] : 0.f;
        sm.colstat[t] = make_float4(m2j, nj, c_j * sj, -c_j * nj);
        sm.colneg[t] = make_float2(m2j, c_j * sj);
        sm.collab[t] = cok ? a.acls[cj] : -2;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      uint8_t* h_row = sm.a + r_in * 128;                                  // + kblock * 16 KB + swizzled 16-byte chunk
      const float rv = valid ? 1.f : 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_row + ch * 32, v);
        const int cb = cbase + ch * 32;
        const int sb = half * (BN / 2) + ch * 32;                          // column index inside the tile
        // One vote per CHUNK: does any row of this warp share its class with any (live) column of the chunk?  (A vote per
        // element is a convergence point that keeps the compiler from interleaving the elements' LDS -> MUFU chains:
        // 6 us per CTA.)  Anchors are grouped by class, so ~9 of 10 (warp, chunk) pairs are all-negative.
        bool mine = !(cb + 32 <= A);                                       // ragged chunks take the general path
        {
          const int L = clab[ch];
          unsigned remaining = 0xffffffffu;
          while (remaining) {
            const int lab = __shfl_sync(0xffffffffu, L, __ffs(remaining) - 1);
            remaining &= ~__ballot_sync(0xffffffffu, L == lab);
            mine |= (lab == rcls);
          }
        }
        const bool general = __any_sync(0xffffffffu, mine);
        ptx::tmem_ld_wait();
        uint32_t packed[16];
        if (!general) {
          // all-negative chunk: H_ij = c_i S_i e_ij + c_j S_j e_ji, no class test, no reciprocal
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 c0 = sm.colneg[sb + j], c1 = sm.colneg[sb + j + 1];      // {m2_j, c_j S_j}: warp-uniform loads
            const float x0 = __uint_as_float(v[j]) * a.k1, x1 = __uint_as_float(v[j + 1]) * a.k1;
            const float h0 = fmaf(cs_i, ptx::ex2_approx(x0 - m2), rv * c0.y * ptx::ex2_approx(x0 - c0.x));
            const float h1 = fmaf(cs_i, ptx::ex2_approx(x1 - m2), rv * c1.y * ptx::ex2_approx(x1 - c1.x));
            packed[j >> 1] = pack_bf16x2(h0, h1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float hv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int jj = j + u;
              const int lj = sm.collab[sb + jj];                           // uniform over the warp
              const float4 cs = sm.colstat[sb + jj];
              const float x = __uint_as_float(v[jj]) * a.k1;
              const float e = ptx::ex2_approx(x - m2);
              const float e2 = ptx::ex2_approx(x - cs.x);
              const bool same = lj == rcls;
              const bool diag = (cb + jj == row);
              const float g_ij = same ? (diag ? 0.f : cn_i * ptx::rcp_approx(e + neg_i)) : cs_i * e;
              const float g_ji = same ? (diag ? 0.f : cs.w * ptx::rcp_approx(e2 + cs.y)) : cs.z * e2;
              hv[u] = (!valid || lj == -2) ? 0.f : (g_ij + g_ji);
            }
            packed[j >> 1] = pack_bf16x2(hv[0], hv[1]);
          }
        }
        // this thread's 32 columns = half a K-block (64 j) of the H tile: 4 chunks of 16 B, 128B-swizzled by row
        const int kblock = half * 2 + (ch >> 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (ch & 1) * 4 + q;
          uint4 val = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
          *reinterpret_cast<uint4*>(h_row + kblock * A_KB_BYTES + ((chunk ^ (r_in & 7)) << 4)) = val;
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();                   // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.g_full);
      PCL_STAMP(6);
      // dA tile: TMEM -> the partial of this column tile
      ptx::mbar_wait(&sm.da_full, 0);
      ptx::tc_fence_after();
      float* dst = a.dpartials + ((int64_t)c * a.a_pad + row) * DDIM + half * 128;
      const uint32_t t_da = tmem_dA + ((uint32_t)(quarter * 32) << 16) + half * 128;
      uint32_t dbuf[2][32];
      ptx::tmem_ld_32x32b_x32(t_da, dbuf[0]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        ptx::tmem_ld_wait();
        if (ch < 3) ptx::tmem_ld_32x32b_x32(t_da + (ch + 1) * 32, dbuf[(ch + 1) & 1]);       // next chunk in flight
        uint32_t(&v)[32] = dbuf[ch & 1];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dst + ch * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      PCL_STAMP(7);
      // ---- last CTA: loss = fixed-order sum of the row tiles' sums / A; counters re-armed for the next launch ----
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (threadIdx.x == 64) {
        bool last = true;
#ifndef PCL_EMULATION
        last = atomicAdd(&a.sync[2], 1u) == n_live - 1;
#else
        last = (a.sync[2] += 1u) == n_live;
#endif
        if (last) {
          __threadfence();
          float t = 0.f;
          for (int i = 0; i < R_live; ++i) t += ld_cg(a.partials + i);
          *a.loss = t / (float)A;
          a.sync[0] = 0u; a.sync[1] = 0u; a.sync[2] = 0u;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  tl_end(a.sync, PCL_TL_FUSED);
#undef PCL_STAMP
}

}  // namespace fused
}  // namespace pcl

using namespace pcl;

// defined in pcl_infonce_tc.cu
namespace pcl { int tc_make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint32_t box_rows); }

int pcl::self_fused_supported(const pcl_tc_desc* d) {
  return d && d->mode == 0 && d->D == fused::DDIM && d->a_rows >= 1 && d->a_rows <= 1024;
}

int pcl::self_fused(const pcl_tc_desc* d, const float* row_m2, float* partials, float* rowstats, float* loss, float* dpartials,
                    unsigned int* sync, void* stream) {
  if (!self_fused_supported(d)) return PCL_ERR_UNSUPPORTED;
  PCL_REQUIRE(d->anchors_bf16 && d->anchor_cls && row_m2 && partials && rowstats && loss && dpartials && sync);
  if (!(d->temperature > 0.f) || !(d->base_temperature > 0.f)) return PCL_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  fused::Args a;
  memset(&a, 0, sizeof(a));
  a.acls = d->anchor_cls; a.plan = d->plan; a.row_m2 = row_m2;
  a.a_rows = d->a_rows;
  const int row_tiles = ceil_div(d->a_rows, fused::BM);
  a.a_pad = row_tiles * fused::BM;
  a.s_max = ceil_div(a.a_pad, fused::BN);
  a.k1 = 1.4426950408889634f / d->temperature;
  a.rs_scale = d->temperature / d->base_temperature;
  a.nan_safe = d->nan_safe;
  a.partials = partials; a.rowstats = rowstats; a.loss = loss; a.dpartials = dpartials; a.sync = sync;
  CUtensorMap tmA, tmC;
  int st = tc_make_tmap(&tmA, d->anchors_bf16, (uint64_t)a.a_pad, fused::BM);
  if (st != PCL_OK) return st;
  st = tc_make_tmap(&tmC, d->anchors_bf16, (uint64_t)a.a_pad, fused::BN);
  if (st != PCL_OK) return st;
  const size_t smem = sizeof(fused::Smem) + 1024;
  PCL_SMEM_OPT_IN(fused::k_self_fused, smem);
  dim3 grid(a.s_max, row_tiles);
#ifdef PCL_EMULATION
  // the host-fiber emulator runs the blocks of a grid one after another: no inter-CTA barrier; one phase per launch
  // (the similarity tile is recomputed in every launch — same values)
  for (int ph = 1; ph <= 4; ph <<= 1) {
    a.phase_mask = ph;
    fused::k_self_fused<<<grid, fused::NUM_THREADS, smem, s>>>(tmA, tmC, a);
    PCL_LAUNCH_CHECK();
  }
#else
  a.phase_mask = 7;
  fused::k_self_fused<<<grid, fused::NUM_THREADS, smem, s>>>(tmA, tmC, a);
  PCL_LAUNCH_CHECK();
#endif
  return PCL_OK;
}
