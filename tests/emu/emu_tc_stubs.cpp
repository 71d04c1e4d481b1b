// TEST INFRASTRUCTURE: entry points of the tcgen05/TMA tensor path (inline PTX, not emulated) for the host-thread build
// of the SIMT kernels — every one reports PCL_ERR_UNSUPPORTED — plus the marker that keeps this build out of the product.
#include "pcl_common.cuh"

extern "C" int pcl_emulated(void) { return 1; }      // contrastiveseg_b200._abi.load refuses a library that exports this

namespace pcl {
int tc_fwd_ex(const pcl_tc_desc*, float*, float*, float*, float*, void*, bool) { return PCL_ERR_UNSUPPORTED; }
int tc_query(const pcl_tc_desc*, int64_t*, float*) { return PCL_ERR_UNSUPPORTED; }
int tc_bwd_ex(const pcl_tc_desc*, const float*, const float*, const float*, float*, float*, void*, int*, int*) {
  return PCL_ERR_UNSUPPORTED;
}
}  // namespace pcl

// sizing only: reports "no extra scratch" so that D == 256 workspaces can be built and used on the exact fp32 path
extern "C" int pcl_tc_sizes(const pcl_tc_desc* d, pcl_sweep_sizes_t* out) {
  if (!d || !out) return PCL_ERR_ARG;
  memset(out, 0, sizeof(*out));
  out->rowstat_f32 = d->a_rows;
  return PCL_OK;
}
extern "C" int pcl_to_bf16(const float*, void*, int64_t, int64_t, void*) { return PCL_ERR_UNSUPPORTED; }
extern "C" int pcl_infonce_tc_fwd(const pcl_tc_desc*, float*, float*, float*, float*, void*) { return PCL_ERR_UNSUPPORTED; }
extern "C" int pcl_infonce_tc_bwd(const pcl_tc_desc*, const float*, const float*, const float*, float*, float*, void*) {
  return PCL_ERR_UNSUPPORTED;
}
extern "C" int pcl_tc_dump_logits(const pcl_tc_desc*, float*, float*, void*) { return PCL_ERR_UNSUPPORTED; }
