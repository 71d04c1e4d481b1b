// CUDA-graph support of the loss step: everything a captured step needs that a kernel parameter
// cannot carry.  A captured launch sequence replays with frozen parameters, so the per-step sampling seed of
// pcl_select_gather (a by-value kernel argument) would repeat; instead the anchor ranks are drawn by this kernel from
// a step counter that lives in device memory and are handed to the selection as an injected rank table
// (pcl_step_desc.ranks), which the selection kernel already understands (the parity tests use the same input).
//
// The draw is the selection kernel's own device RNG (keyed bijection on [0, n), pcl_common.cuh): view j of an
// (image, class, hard|easy) group takes element perm_key(j) — replaces torch.randperm(n)[:k] of
// lib/loss/loss_contrast.py:79-82 — so a graph replay samples exactly like an eager step with seed (base, counter).
#include "pcl_common.cuh"

namespace pcl {

__global__ void __launch_bounds__(1024)
k_step_ranks(pcl_geom g, const int32_t* __restrict__ plan, uint64_t base_seed, unsigned long long* __restrict__ counter,
             int32_t* __restrict__ ranks) {
  __shared__ unsigned long long s_step;
  if (threadIdx.x == 0) {
    s_step = *counter;
    *counter = s_step + 1ull;                 // the next replay draws a fresh set
  }
  __syncthreads();
  const uint64_t seed = (base_seed * 0x9E3779B97F4A7C15ull + (uint64_t)s_step + 1ull);
  const int TC = plan[PCL_PLAN_TC], V = plan[PCL_PLAN_V];
  const int ms = g.max_samples, K = g.K;
  for (int i = threadIdx.x; i < ms; i += blockDim.x) {
    int rank = 0;
    if (V > 0 && i < TC * V) {
      const int t = i / V, v = i - t * V;
      const int32_t* q = plan + PCL_PLAN_HEADER + (int64_t)t * 8;
      const int b = q[0], c = q[1], nh = q[2], ne = q[3], kh = q[4];
      const bool easy = v >= kh;
      const int j = easy ? v - kh : v;
      const int n = easy ? ne : nh;
      rank = (int)keyed_perm((uint32_t)j, (uint32_t)n, mix64(seed ^ ((uint64_t)(b * K + c) << 1 | (easy ? 1u : 0u))));
    }
    ranks[i] = rank;
  }
}

}  // namespace pcl

extern "C" int pcl_step_ranks(const pcl_step_desc* d, uint64_t* step_counter, int32_t* ranks, void* stream) {
  if (!d || !d->plan || !step_counter || !ranks) return PCL_ERR_ARG;
  if (d->g.max_samples <= 0 || d->g.K <= 0) return PCL_ERR_ARG;
  pcl::k_step_ranks<<<1, 1024, 0, (cudaStream_t)stream>>>(d->g, d->plan, d->seed, (unsigned long long*)step_counter,
                                                           ranks);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}
