// Library-level entry points: version, status text, CUDA error capture.
#include "pcl_common.cuh"

namespace pcl {
static thread_local char g_last_error[512] = "";

void set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file,
           line);
}
void set_error_text(const char* text) { snprintf(g_last_error, sizeof(g_last_error), "%s", text); }
static unsigned long long g_launches = 0ull;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
}  // namespace pcl

extern "C" uint64_t pcl_launch_count(void) { return __atomic_load_n(&pcl::g_launches, __ATOMIC_RELAXED); }

extern "C" int pcl_version(void) { return PCL_VERSION; }

extern "C" const char* pcl_strerror(int status) {
  switch (status) {
    case PCL_OK: return "ok";
    case PCL_ERR_ARG: return "invalid argument";
    case PCL_ERR_CUDA: return "CUDA error (see pcl_last_cuda_error)";
    case PCL_ERR_UNSUPPORTED: return "unsupported configuration";
    case PCL_ERR_SHAPE: return "shape mismatch (the reference would raise an index error)";
    default: return "unknown status";
  }
}

extern "C" const char* pcl_last_cuda_error(void) { return pcl::g_last_error; }

extern "C" int64_t pcl_abi_sizeof(int struct_id) {
  switch (struct_id) {
    case 0: return (int64_t)sizeof(pcl_geom);
    case 1: return (int64_t)sizeof(pcl_select_sizes_t);
    case 2: return (int64_t)sizeof(pcl_sweep_desc);
    case 3: return (int64_t)sizeof(pcl_sweep_sizes_t);
    case 4: return (int64_t)sizeof(pcl_bank_geom);
    case 5: return (int64_t)sizeof(pcl_tc_desc);
    case 6: return (int64_t)sizeof(pcl_step_desc);
    default: return -1;
  }
}

extern "C" int pcl_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    pcl::set_cuda_error(e, __FILE__, __LINE__);
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}
