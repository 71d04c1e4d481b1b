// Shared helpers for the pixel-contrast engine kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pcl.h"

namespace pcl {

// ---- error plumbing -----------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* file, int line);
void set_error_text(const char* text);            // driver-API failures (tensor-map encoding): same per-thread slot

#define PCL_CUDA(expr)                                                  \
  do {                                                                  \
    cudaError_t e__ = (expr);                                           \
    if (e__ != cudaSuccess) {                                           \
      ::pcl::set_cuda_error(e__, __FILE__, __LINE__);                   \
      return PCL_ERR_CUDA;                                              \
    }                                                                   \
  } while (0)

void count_launch();
#define PCL_LAUNCH_CHECK()          \
  do {                              \
    ::pcl::count_launch();          \
    PCL_CUDA(cudaGetLastError());   \
  } while (0)

#define PCL_REQUIRE(cond)                 \
  do {                                    \
    if (!(cond)) return PCL_ERR_ARG;      \
  } while (0)

// internal (not part of the C ABI): fused variants used by pcl_step_forward / pcl_step_backward
int select_gather_ex(const pcl_geom* g, const float* embed, const uint16_t* keys, const int32_t* chunk_pref,
                     const int32_t* plan, const int32_t* ranks, uint64_t seed, int normalize, int32_t* anchor_meta,
                     float* anchors_f32, void* anchors_bf16, float* inv_norm, float* norm_max, float* row_m2,
                     float m2_scale, float* partials, int64_t n_slot_rows, void* stream,
                     const unsigned long long* seed_ctr = nullptr, unsigned int* dbg = nullptr,
                     const int32_t* prev_rows = nullptr, float* grad_clear = nullptr);
int tc_fwd_ex(const pcl_tc_desc* d, float* row_m2, float* partials, float* rowstats, float* loss, void* stream,
              bool skip_prep, unsigned long long* step_counter = nullptr);
int simt_fwd_ex(const pcl_sweep_desc* d, float* partials, float* rowstats, float* loss, void* stream,
                unsigned long long* step_counter);
int tc_fwd_topk_ex(const pcl_tc_desc* d, int k, uint32_t* scratch, float* row_m2, float* partials, float* rowstats,
                   float* loss, void* stream, bool skip_prep, unsigned long long* step_counter);
int tc_query(const pcl_tc_desc* d, int64_t* n_slot_rows, float* m2_scale);
int tc_bwd_ex(const pcl_tc_desc* d, const float* row_m2, const float* rowstats, const float* grad_loss, float* dpartials,
              float* dA, void* stream, int* splits_out, int* a_pad_out, const uint32_t* topk_scratch = nullptr);
int zero_scatter_reduce(const pcl_geom* g, const int32_t* plan, const int32_t* anchor_meta, const float* dpartials,
                        int splits, int a_pad, float inv_T, const float* grad_loss, float* grad_embed, void* stream);

int self_fused_supported(const pcl_tc_desc* d);
int self_fused(const pcl_tc_desc* d, const float* row_m2, float* partials, float* rowstats, float* loss, float* dpartials,
               unsigned int* sync, void* stream);
int class_stats_plan(const pcl_geom* g, const int64_t* labels, const float* seg, const int64_t* predict, uint16_t* keys,
                     int32_t* chunk_pref, int32_t* counts, int32_t* plan, unsigned int* done_ctr, void* stream);
int scatter_reduce_rows(const pcl_geom* g, const int32_t* plan, const int32_t* anchor_meta, const float* dpartials, int splits,
                        int split_cols, int a_pad, float inv_T, const float* grad_scale, float* grad_embed, unsigned long long* step_counter,
                        void* stream, unsigned int* dbg = nullptr, int32_t* prev_rows = nullptr);
int fill_zero(void* ptr, uint64_t bytes, void* stream, unsigned int* dbg, int reserve_sms);
int scatter_rows(const pcl_geom* g, const int32_t* plan, const int32_t* anchor_meta, const float* dA,
                 const float* anchors_f32, const float* inv_norm, int normalize, float* grad_embed, void* stream);

// ---- per-device kernel attributes ----------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (function, device) and SETS the limit (a later call with
// a smaller size lowers it).  Every kernel therefore opts in ONCE per device to the largest size it can ever need
// (known at compile time: D <= 256); the bit mask is indexed by the current device, so a process that drives several
// GPUs gets every device configured.
struct SmemOptIn { unsigned long long mask; };
template <typename F>
static inline int smem_opt_in(SmemOptIn& st, F func, size_t max_bytes) {
  int dev = 0;
  PCL_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(&st.mask, __ATOMIC_ACQUIRE) & bit) return PCL_OK;
  PCL_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_bytes));
  __atomic_fetch_or(&st.mask, bit, __ATOMIC_RELEASE);
  return PCL_OK;
}
#define PCL_SMEM_OPT_IN(func, max_bytes)                                   \
  do {                                                                      \
    static ::pcl::SmemOptIn st__ = {0ull};                                  \
    int rc__ = ::pcl::smem_opt_in(st__, func, (max_bytes));                 \
    if (rc__ != PCL_OK) return rc__;                                        \
  } while (0)

// SM count of the current device (cached per device; 148 when sizing on a host without one)
int num_sms();

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- optional in-graph timeline (diagnostics) ----------------------------------------------------------
// The step's `sync` buffer (pcl_step_desc.sync, >= 1024 words) doubles as a timeline when word 7 is non-zero: words
// [16, 16 + 4*16) hold, per kernel slot k, {min block start, max block end} of %globaltimer (ns) — the only way to see
// how the kernels of a captured graph actually overlap (ncu serialises them).  Costs one predicated load per block when off.
enum { PCL_TL_KEYS = 0, PCL_TL_SELECT = 1, PCL_TL_FUSED = 2, PCL_TL_SCATTER = 3, PCL_TL_FILL = 4 };
__device__ __forceinline__ unsigned long long pcl_globaltimer() {
#ifdef PCL_EMULATION
  return 0ull;
#else
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}
__device__ __forceinline__ void tl_begin(unsigned int* dbg, int slot) {
  if (dbg != nullptr && dbg[7] != 0u && threadIdx.x == 0)
    atomicMin(reinterpret_cast<unsigned long long*>(dbg + 16) + 2 * slot, pcl_globaltimer());
}
__device__ __forceinline__ void tl_end(unsigned int* dbg, int slot) {
  if (dbg != nullptr && dbg[7] != 0u && threadIdx.x == 0)
    atomicMax(reinterpret_cast<unsigned long long*>(dbg + 16) + 2 * slot + 1, pcl_globaltimer());
}

// ---- device helpers -------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Class rank used for the sorted anchor layout: 1,2,...,K-1,0  (class 0 last: in bank mode its
// positives are the analytic zero tail, Q3).
__device__ __host__ __forceinline__ int class_rank(int c, int K) { return c == 0 ? K - 1 : c - 1; }

// ATen nearest-neighbour source index (UpSampleNearest: min(int(floorf(dst*scale)), in-1), scale=in/out in fp32).
__device__ __host__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Keyed bijection on [0, n): an 8-round balanced Feistel network on the next even number of bits, with cycle walking.
// Used when no permutation table is injected: view j of a group takes element perm(j), so distinct j give distinct
// elements without any sequential state (replaces torch.randperm(n)[:k]).  The round function hashes (key, round, half)
// with the full 64-bit key, so small groups are sampled without the pairwise structure of multiply/xorshift rounds
// (measured on n = 60, k = 10: pair co-occurrence chi2/dof ~1.0, was 3.1; n = 7, k = 3: ~1.0, was 26;
// tests/test_host_logic.py checks the host model of this function, tests/test_emu_kernels.py that the kernels follow it).
__device__ __forceinline__ uint32_t keyed_perm(uint32_t j, uint32_t n, uint64_t key) {
  if (n <= 1) return 0;
  const int bits = 32 - __clz(n - 1);                  // ceil(log2 n) >= 1
  const int hb = (bits + 1) >> 1;                      // half width (<= 16): the network permutes [0, 2^(2 hb)) >= [0, n)
  const uint32_t hmask = (1u << hb) - 1u;
  uint32_t x = j;
  do {
    uint32_t L = x >> hb, R = x & hmask;
#pragma unroll
    for (int r = 0; r < 8; ++r) {                       // 4 rounds leave visible structure on domains of 16 elements
      const uint32_t f = (uint32_t)mix64(key + (uint64_t)r * 0x9E3779B97F4A7C15ull + R) & hmask;
      const uint32_t t = R;
      R = L ^ f;
      L = t;
    }
    x = (L << hb) | R;
  } while (x >= n);                                    // domain <= 4n: a few iterations at most on average
  return x;
}

}  // namespace pcl
