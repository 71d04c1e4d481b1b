/* Minimal C client of the engine's C ABI (include/pcl.h): what a non-Python binding does before its first compute call —
 * check the ABI version and struct layouts, size the caller-owned scratch for a geometry.  Host-only entry points, so it
 * also runs on a machine without a GPU (compute calls would return PCL_ERR_CUDA there).
 *
 *   gcc -std=c99 -I include examples/c_abi_sizes.c -L contrastiveseg_b200/lib -lpcl_b200 \
 *       -Wl,-rpath,$PWD/contrastiveseg_b200/lib -o /tmp/c_abi_sizes && /tmp/c_abi_sizes
 */
#include <stdio.h>
#include <string.h>

#include "pcl.h"

int main(void) {
  if (pcl_version() / 100 != PCL_VERSION / 100) { fprintf(stderr, "ABI major version mismatch\n"); return 1; }
  if (pcl_abi_sizeof(0) != (int64_t)sizeof(pcl_geom) || pcl_abi_sizeof(6) != (int64_t)sizeof(pcl_step_desc)) {
    fprintf(stderr, "struct layout mismatch\n");
    return 1;
  }
  /* BASELINE configs[1]: HRNet-W48 on 1024x512 crops, batch 8, stride-4 embedding (256 x 128 x 256), 19 classes */
  pcl_geom g;
  memset(&g, 0, sizeof g);
  g.B = 8; g.D = 256; g.h = 128; g.w = 256; g.Himg = 512; g.Wimg = 1024; g.K = 19;
  g.max_samples = 1024; g.max_views = 100; g.ignore_label = -1;
  pcl_select_sizes_t sel;
  int st = pcl_select_sizes(&g, &sel);
  if (st != PCL_OK) { fprintf(stderr, "pcl_select_sizes: %s\n", pcl_strerror(st)); return 1; }
  pcl_sweep_desc sw;
  memset(&sw, 0, sizeof sw);
  sw.a_rows = g.max_samples; sw.D = g.D; sw.mode = 0; sw.temperature = 0.1f; sw.base_temperature = 0.07f;
  pcl_sweep_sizes_t ss;
  st = pcl_sweep_sizes(&sw, &ss);
  if (st != PCL_OK) { fprintf(stderr, "pcl_sweep_sizes: %s\n", pcl_strerror(st)); return 1; }
  printf("pcl %d  devices %d\n", pcl_version(), pcl_device_count());
  printf("selection scratch: keys %lld u16, chunk histograms %lld i32 (%d chunks), plan %lld i32, anchor meta %lld i32\n",
         (long long)sel.keys_u16, (long long)sel.chunk_pref_i32, (int)sel.nchunk, (long long)sel.plan_i32,
         (long long)sel.anchor_meta_i32);
  printf("exact sweep: %d row tiles x %d column splits, partials %lld f32, backward partials %lld f32\n", (int)ss.row_tiles,
         (int)ss.splits, (long long)ss.partial_f32, (long long)ss.dpartial_f32);
  /* an invalid geometry is refused with a status code, never a crash */
  g.K = 1000;
  if (pcl_select_sizes(&g, &sel) != PCL_ERR_ARG) { fprintf(stderr, "invalid geometry accepted\n"); return 1; }
  return 0;
}
