/*
 * pcl.h — C ABI of the B200 pixel-contrast loss engine (libpcl_b200.so).
 *
 * The reference (tfzhou/ContrastiveSeg) has no FFI on this path: it is Python/ATen behind three
 * nn.Module classes and one trainer method.  This header is the boundary a
 * maintainer binds instead; each entry point names the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name starts with h_;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant and never throws; the
 *     library keeps no state between calls (apart from a mutex-protected cache of encoded TMA descriptors, which
 *     describe addresses/shapes only); the return value is 0 or a negative pcl_status;
 *   - scratch comes from caller-provided buffers (sizes: pcl_select_sizes / pcl_sweep_sizes), so the
 *     caller's allocator owns all memory and the sequence is CUDA-graph capturable;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns PCL_ERR_CUDA;
 *   - one device per process (the one-process-per-GPU model of the reference's DDP launcher,
 *     lib/utils/distributed.py): opt-in shared-memory limits are raised once per process with cudaFuncSetAttribute,
 *     which is a per-device setting, so driving a second device from the same process is not supported.
 */
#ifndef PCL_H_
#define PCL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCL_VERSION 100          /* major*100 + minor */
#define PCL_MAX_CLASSES 256      /* class ids tracked: [0, K), K <= 256 */
#define PCL_CHUNK 1024           /* pixels per selection chunk */
#define PCL_ROW_TILE 64          /* anchor rows per SIMT tile */
#define PCL_PLAN_HEADER 16       /* int32 words at the start of the plan buffer */

typedef enum {
  PCL_OK = 0,
  PCL_ERR_ARG = -1,        /* bad argument (null pointer, size out of range, unsupported D ...) */
  PCL_ERR_CUDA = -2,       /* CUDA runtime error (see pcl_last_cuda_error) or no device */
  PCL_ERR_UNSUPPORTED = -3,/* valid request the engine does not implement (e.g. tensor path with D != 256) */
  PCL_ERR_SHAPE = -4       /* reference would raise an index error (e.g. Q6 label grid larger than feature map) */
} pcl_status;

/* plan buffer header words (int32), written by pcl_plan_anchors */
enum { PCL_PLAN_TC = 0, PCL_PLAN_V = 1, PCL_PLAN_A = 2, PCL_PLAN_FLAGS = 3, PCL_PLAN_NPAIR_MAX = 4 };
/* PCL_PLAN_FLAGS bits */
enum { PCL_FLAG_NO_CLASS = 1, PCL_FLAG_ZERO_VIEWS = 2, PCL_FLAG_SPLIT_ERROR = 4 };

int         pcl_version(void);
const char* pcl_strerror(int status);
const char* pcl_last_cuda_error(void);          /* text of the last CUDA error seen by this thread */
int         pcl_device_count(void);             /* 0 without a usable CUDA device */
/* FFI layout self-check: sizeof() of a host struct as this library was compiled.  struct_id: 0 pcl_geom,
 * 1 pcl_select_sizes_t, 2 pcl_sweep_desc, 3 pcl_sweep_sizes_t, 4 pcl_bank_geom, 5 pcl_tc_desc, 6 pcl_step_desc;
 * -1 for an unknown id.  A binding compares these with its own struct definitions before the first call. */
int64_t     pcl_abi_sizeof(int struct_id);
/* Diagnostics: number of kernels this library has launched (or recorded into a stream capture) since it was loaded,
 * process-wide.  The difference around one eager pass of a call sequence = kernels per step (bench.py: gpu_launches). */
uint64_t    pcl_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Geometry of one loss call (host struct, passed by pointer).
 * ----------------------------------------------------------------------------------------------*/
typedef struct {
  int32_t B, D, h, w;          /* embedding (B,D,h,w) fp32 NCHW                                  */
  int32_t Himg, Wimg;          /* label map (B,Himg,Wimg) int64                                  */
  int32_t K;                   /* number of class ids, labels outside [0,K) count as ignored     */
  int32_t max_samples;         /* contrast.max_samples                                            */
  int32_t max_views;           /* contrast.max_views                                              */
  int32_t ignore_label;        /* loss.params.ce_ignore_index (default -1)                        */
} pcl_geom;

/* Element counts of the selection scratch buffers for a geometry. */
typedef struct {
  int64_t keys_u16;            /* B*h*w             uint16 pixel keys (2*label + easy, 2K = ignored) */
  int64_t chunk_pref_i32;      /* B*2K*nchunk       per-chunk histogram of the pixel keys           */
  int64_t counts_i32;          /* B*2K              per-image key totals                            */
  int64_t plan_i32;            /* header + 8 words per (image,class) pair                           */
  int64_t anchor_meta_i32;     /* 4*max_samples     pixel, image, class, reference row              */
  int32_t nchunk;
  int32_t max_pairs;           /* B*K */
} pcl_select_sizes_t;

int pcl_select_sizes(const pcl_geom* g, pcl_select_sizes_t* out);

/* ------------------------------------------------------------------------------------------------
 * a3 + a4 (first half): label down-sampling, argmax, hard/easy key per pixel, per-chunk histograms.
 * Replaces lib/loss/loss_contrast.py:131-134 (nearest interpolation of the label map),
 * :183 (torch.max(seg,1)) and the unique/nonzero scans of :37-39,:60-61.
 *   labels   (B,Himg,Wimg) int64
 *   seg      (B,K,h,w) fp32 or NULL      — if non-NULL, predict = argmax over the K planes
 *   predict  (B,h,w) int64 or NULL       — used when seg is NULL
 * ----------------------------------------------------------------------------------------------*/
int pcl_class_stats(const pcl_geom* g, const int64_t* labels, const float* seg, const int64_t* predict,
                    uint16_t* keys, int32_t* chunk_pref, int32_t* counts, void* stream);

/* a4: class filter (count > max_views), TC, V = min(max_samples / TC, max_views), hard/easy split
 * rule and the class-sorted anchor row layout.  Replaces lib/loss/loss_contrast.py:37-48,63-77.
 * Reads the per-image key totals `counts` written by pcl_class_stats; writes the plan. */
int pcl_plan_anchors(const pcl_geom* g, const int32_t* counts, int32_t* plan, void* stream);

/* a4 (second half) + gather: choose the pixels of every anchor row and gather their embeddings.
 * Replaces lib/loss/loss_contrast.py:79-86 (randperm, index, X_[ptr] = X[ii, indices]) and the
 * 268 MB permute+contiguous of :141-142 (never materialised here).
 *   ranks     NULL -> device RNG (keyed Feistel permutation, `seed`), no host round trip;
 *             else (TC,V) int32 table: ranks[t*V+v] = perm value used for view v of pair t
 *             (hard ranks first, then easy ranks) — replay of the reference's torch.randperm draws.
 *   normalize 0: embed is already L2-normalised (projection.py:24 ran upstream);
 *             1: embed is the raw projection, only the gathered columns are normalised here.
 * Outputs (row s = class-sorted anchor index, rows >= A are zero-filled):
 *   anchor_meta (4,max_samples) int32: pixel, image, class, reference row r = v*TC + t
 *   anchors_f32 (max_samples,D), anchors_bf16 (max_samples rounded up to 128, D) or NULL,
 *   inv_norm (max_samples) fp32 (1/||x|| of the raw column; 1 when normalize == 0),
 *   norm_max  optional (may be NULL): 1 float, max over anchors of ||a||. */
int pcl_select_gather(const pcl_geom* g, const float* embed, const uint16_t* keys, const int32_t* chunk_pref,
                      const int32_t* plan, const int32_t* ranks, uint64_t seed, int normalize,
                      int32_t* anchor_meta, float* anchors_f32, void* anchors_bf16, float* inv_norm,
                      float* norm_max, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a5 + a6: the contrast set and the InfoNCE sweep.
 * Replaces lib/loss/loss_contrast.py:91-128 and lib/loss/loss_contrast_mem.py:91-152.
 * ----------------------------------------------------------------------------------------------*/
typedef struct {
  /* anchors (rows); device scalar plan[PCL_PLAN_A] gives the live row count, a_rows its upper bound */
  const float*   anchors;        /* (a_rows, D) fp32, class-sorted                                   */
  const int32_t* anchor_cls;     /* (a_rows) int32                                                   */
  const int32_t* diag_col;       /* (a_rows) int32 contrast column removed from the positives (Q1)   */
  const int32_t* plan;           /* plan buffer or NULL (then a_live = a_rows)                       */
  int32_t a_rows;
  int32_t D;
  /* contrast set.  mode 0: self-contrast (columns == anchors, loss_contrast.py:91-128)
   *                mode 1: memory bank, read in place from the two queues (mem:91-105,221):
   *                        column n = (c-1)*R + q, R = M0+M1, c = 1..K-1 (class 0 skipped, Q2), q < M0 -> segment
   *                        row, else pixel row; plus `R` all-zero columns of label 0 handled analytically (Q3)
   *                mode 2: explicit matrix `contrast` (n_cols, D) with labels `contrast_cls`           */
  int32_t mode;
  const float*   segment_queue;  /* (K, M0, D) fp32                                                  */
  const float*   pixel_queue;    /* (K, M1, D) fp32, or NULL with bank_M1 == 0 (pre-concatenated queue
                                    (K, R, D) passed as segment_queue with bank_M0 == R)             */
  int32_t bank_K, bank_M0, bank_M1;
  const float*   contrast;       /* mode 2                                                           */
  const int32_t* contrast_cls;   /* mode 2                                                           */
  int32_t n_cols;                /* mode 2                                                           */
  float temperature, base_temperature;
  int32_t nan_safe;              /* 0: rows without positives give NaN like the reference (Q8); 1: they give 0 */
} pcl_sweep_desc;

typedef struct {
  int64_t n_real_cols;         /* streamed columns                                                  */
  int32_t row_tiles, splits;   /* launch grid of the SIMT sweep                                     */
  int64_t partial_f32;         /* splits * a_rows_padded floats, per partial statistic (5 of them)  */
  int64_t rowstat_f32;         /* a_rows floats, per final statistic (6: m, neg, possum, s, npos, row_loss) */
  int64_t dpartial_f32;        /* splits * a_rows_padded * D floats (backward partial dA)           */
} pcl_sweep_sizes_t;

int pcl_sweep_sizes(const pcl_sweep_desc* d, pcl_sweep_sizes_t* out);

/* Forward: loss (1 float) and per-row statistics.  partials: 5 * partial_f32 floats of scratch;
 * rowstats: 6 * rowstat_f32 floats (m, neg, possum, s, npos, row_loss), kept for the backward. */
int pcl_infonce_fwd(const pcl_sweep_desc* d, float* partials, float* rowstats, float* loss, void* stream);

/* Backward: dA (a_rows, D) = d loss / d anchors (closed form), multiplied by the
 * upstream gradient *grad_loss (device scalar, may be NULL = 1).  dpartials: scratch. */
int pcl_infonce_bwd(const pcl_sweep_desc* d, const float* rowstats, const float* grad_loss, float* dpartials,
                    float* dA, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a10: per-anchor top-k hard-negative selection on the exact fp32 sweep — an extension the
 * reference code does not have (lib/loss/loss_contrast.py:116-117 and loss_contrast_mem.py:140-141 sum ALL
 * negatives), default off.  Neg_i = sum of exp(l - m_i) over the k negatives of anchor i with the largest
 * logits; ties at the k-th value share the remaining slots equally (order independent); rows with <= k
 * negatives keep all of them (== pcl_infonce_fwd).  Exact 3-level radix select on the fp32 logits, streamed:
 * three histogram sweeps + the weighted NEG sweep + the stock POS sweep; the A x N logits are never stored.
 *   scratch  uint32 words, pcl_topk_scratch_u32(d) of them (negative = pcl_status): per-row radix bins followed by
 *            the selection result sel[4][a_rows] = tau key (order-preserving integer image of the k-th largest
 *            negative logit; 0 = all kept), tie weight (float bits), G = #negatives above tau, E = #ties at tau.
 *            The SAME scratch must be passed to the backward.
 *   partials / rowstats / loss / dpartials / dA: as pcl_infonce_fwd / pcl_infonce_bwd.
 * ----------------------------------------------------------------------------------------------*/
int64_t pcl_topk_scratch_u32(const pcl_sweep_desc* d);
int pcl_infonce_topk_fwd(const pcl_sweep_desc* d, int32_t k, uint32_t* scratch, float* partials, float* rowstats,
                         float* loss, void* stream);
int pcl_infonce_topk_bwd(const pcl_sweep_desc* d, int32_t k, uint32_t* scratch, const float* rowstats,
                         const float* grad_loss, float* dpartials, float* dA, void* stream);

/* dA -> dense NCHW gradient of the embedding (zero fill + scatter; with normalize == 1 the
 * projection-normalise backward of the touched columns is fused in: dx = (g - y (y.g)) / ||x||).
 * Replaces the autograd backward of loss_contrast.py:83-85,:141-142 (152 SelectBackward0 zero fills). */
int pcl_scatter_grad(const pcl_geom* g, const int32_t* plan, const int32_t* anchor_meta, const float* dA,
                     const float* anchors_f32, const float* inv_norm, int normalize, float* grad_embed,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * a1: projection-head normalise, lib/models/modules/projection.py:24  (F.normalize(x, p=2, dim=1)).
 * ----------------------------------------------------------------------------------------------*/
int pcl_l2norm_fwd(const float* x, float* y, int32_t B, int32_t D, int64_t HW, void* stream);
int pcl_l2norm_bwd(const float* x, const float* gy, float* gx, int32_t B, int32_t D, int64_t HW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * §8f row 1 (next row): the segmentation CE of ContrastCELoss.forward, fused.  Replaces
 * lib/loss/loss_contrast.py:180-181 (F.interpolate(seg, (Himg,Wimg), bilinear, align_corners=True)) +
 * lib/loss/loss_helper.py:169-212 (nn.CrossEntropyLoss(weight, ignore_index, reduction mean)) and their autograd
 * backward; the (B,K,Himg,Wimg) up-sampled logits are never materialised.
 *   seg (B,K,h,w) fp32, target (B,Himg,Wimg) int64, class_weight (K) fp32 or NULL, scratch: pcl_seg_ce_scratch_floats.
 * ----------------------------------------------------------------------------------------------*/
int64_t pcl_seg_ce_scratch_floats(int32_t B, int32_t Himg, int32_t Wimg);
int pcl_seg_ce_fwd(const float* seg, const int64_t* target, const float* class_weight, int32_t B, int32_t K, int32_t h,
                   int32_t w, int32_t Himg, int32_t Wimg, int32_t ignore_index, float* scratch, float* loss, void* stream);
int pcl_seg_ce_bwd(const float* seg, const int64_t* target, const float* class_weight, int32_t B, int32_t K, int32_t h,
                   int32_t w, int32_t Himg, int32_t Wimg, int32_t ignore_index, const float* scratch,
                   const float* grad_loss, float* dseg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a7 + a8: memory bank.  Replaces segmentor/trainer_contrastive.py:102-138 (_dequeue_and_enqueue).
 * Step 1 (parallel): build this rank's enqueue packet.  Step 2 (ordered): apply the packets of all
 * ranks in rank order — world 1 reproduces the reference exactly, world > 1 is the allgather merge
 * (SURVEY §8e).  Packet layout per (image b, class c) slot, slot = b*K + c, `slot_f32` floats:
 *   [0] n pixels of the class in the sub-sampled label map (0 = slot empty)   [1] K' = min(n, F)
 *   [2 .. 2+D)            normalised segment mean          (trainer_contrastive.py:120-122)
 *   [2+D .. 2+D+F*D)      K' normalised pixel rows         (trainer_contrastive.py:126-130)
 * ----------------------------------------------------------------------------------------------*/
typedef struct {
  int32_t B, D, h, w;          /* keys (B,D,h,w) fp32                                               */
  int32_t Himg, Wimg;          /* labels (B,Himg,Wimg) int64                                        */
  int32_t K, M;                /* queues (K,M,D)                                                    */
  int32_t network_stride;      /* labels[:, ::s, ::s]  (Q6: flat index of that grid == feature column) */
  int32_t pixel_update_freq;   /* F                                                                 */
} pcl_bank_geom;

int64_t pcl_bank_packet_floats(const pcl_bank_geom* g);    /* floats in one rank's packet           */
int64_t pcl_bank_scratch_floats(const pcl_bank_geom* g);   /* floats of scratch for pcl_bank_packet */

/* ranks: NULL -> device RNG with `seed`; else (B*K, F) int32 table of perm values per slot. */
int pcl_bank_packet(const pcl_bank_geom* g, const float* keys, const int64_t* labels, const int32_t* ranks,
                    uint64_t seed, float* scratch, float* packet, void* stream);

/* Same for a captured launch sequence (CUDA graph): kernel arguments are frozen at capture, so the per-call part of the
 * seed is read from device memory: effective seed = seed + *seed_offset (uint64 on the device, e.g. the step counter
 * pcl_step_ranks advances; NULL -> 0).  Device RNG only. */
int pcl_bank_packet_dev(const pcl_bank_geom* g, const float* keys, const int64_t* labels, uint64_t seed,
                        const uint64_t* seed_offset, float* scratch, float* packet, void* stream);

/* Apply `world` packets (contiguous, rank-major) to the queues in place, with the result of the reference's sequential
 * per-image loop over the gathered batch (trainer_contrastive.py:96-139): every queue row is written once, by the LAST
 * of the world * B slots that would have written it.  world * B <= 1024 (PCL_ERR_UNSUPPORTED otherwise).
 * shadow_bf16 (optional, ((K-1)*2M rounded up to 256, D) bf16) is the engine's class-blocked bf16 copy used by the
 * tensor path. */
int pcl_bank_apply(const pcl_bank_geom* g, const float* packets, int32_t world, float* segment_queue,
                   int64_t* segment_queue_ptr, float* pixel_queue, int64_t* pixel_queue_ptr,
                   void* shadow_bf16, void* stream);

/* Same; additionally advances *enqueue_counter (device uint64, may be NULL) when done: the seed offset of the NEXT
 * pcl_bank_packet_dev of a captured sequence. */
int pcl_bank_apply_ctr(const pcl_bank_geom* g, const float* packets, int32_t world, float* segment_queue,
                       int64_t* segment_queue_ptr, float* pixel_queue, int64_t* pixel_queue_ptr,
                       void* shadow_bf16, uint64_t* enqueue_counter, void* stream);

/* Rebuild the bf16 shadow from the fp32 queues (after a checkpoint load / external write). */
int pcl_bank_shadow_rebuild(const float* segment_queue, const float* pixel_queue, int32_t K, int32_t M,
                            int32_t D, void* shadow_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a6 on the tensor cores (tcgen05 + TMA, bf16 operands, fp32 accumulate in TMEM), D must be 256.
 * Same semantics and row-statistic layout as pcl_infonce_fwd; operands are bf16 copies:
 *   anchors_bf16   (a_rows rounded up to 128, 256) bf16 — written here from anchors_f32 when that is non-NULL
 *   contrast_bf16  mode 1: the bank shadow ((K-1)*R rounded up to 256 rows, class-blocked, see pcl_bank_apply);
 *                  mode 2: (contrast_rows_alloc, 256) rows with labels contrast_cls (sorted = labels are
 *                  non-decreasing, enables the no-compare fast path); mode 0: ignored (columns == anchors)
 *   contrast_norm_bound  upper bound of the contrast rows' L2 norm (1 for a normalised bank); only has to keep
 *                  exp() in range — the loss does not depend on it
 * ----------------------------------------------------------------------------------------------*/
typedef struct {
  const float*   anchors_f32;    /* (a_rows, 256) or NULL when anchors_bf16 is already filled                */
  void*          anchors_bf16;
  const int32_t* anchor_cls; const int32_t* diag_col; const int32_t* plan;
  int32_t a_rows, D;
  int32_t mode;
  const void*    contrast_bf16;
  const int32_t* contrast_cls;
  int64_t n_cols; int64_t contrast_rows_alloc;
  int32_t bank_K, bank_R, sorted;
  float contrast_norm_bound;
  float temperature, base_temperature;
  int32_t nan_safe;
  int32_t neg_only;              /* 1: stop after the similarity + negative-sum sweep (the dense contraction alone,
                                    used to report its roofline fraction); loss / rowstats are then not final */
} pcl_tc_desc;

int pcl_tc_sizes(const pcl_tc_desc* d, pcl_sweep_sizes_t* out);
/* a10 on the tensor path (same semantics as pcl_infonce_topk_fwd/bwd, selection on the fp32-accumulated logits of the bf16
 * operands): scratch = pcl_tc_topk_scratch_u32(d) uint32 (per-row radix bins + the selection state, kept for the
 * backward); three extra similarity sweeps whose epilogue bins the logit keys (red.global), then the weighted sweep. */
int64_t pcl_tc_topk_scratch_u32(const pcl_tc_desc* d);
int pcl_infonce_tc_topk_fwd(const pcl_tc_desc* d, int32_t k, uint32_t* scratch, float* row_m2, float* partials,
                            float* rowstats, float* loss, void* stream);
int pcl_infonce_tc_topk_bwd(const pcl_tc_desc* d, int32_t k, const uint32_t* scratch, const float* row_m2,
                            const float* rowstats, const float* grad_loss, float* dpartials, float* dA, void* stream);
int pcl_to_bf16(const float* src, void* dst_bf16, int64_t n_real, int64_t n_total, void* stream);
/* row_m2: (a_rows rounded up to 256) + 512 floats of scratch (per-row stabiliser, kept for the backward,
 * followed by the per-label column bounds of a sorted explicit contrast set). */
int pcl_infonce_tc_fwd(const pcl_tc_desc* d, float* row_m2, float* partials, float* rowstats, float* loss,
                       void* stream);
/* Backward on the tensor cores: dA (a_rows, 256) fp32 = d loss / d anchors * (*grad_loss or 1).  row_m2 and rowstats
 * come from pcl_infonce_tc_fwd; dpartials: pcl_tc_sizes().dpartial_f32 floats of scratch. */
int pcl_infonce_tc_bwd(const pcl_tc_desc* d, const float* row_m2, const float* rowstats, const float* grad_loss,
                       float* dpartials, float* dA, void* stream);
/* Pipeline self-test: raw similarity tiles S = A.C^T into dump[(a_rows up to 256) x (n_cols up to 256)] fp32. */
int pcl_tc_dump_logits(const pcl_tc_desc* d, float* row_m2, float* dump, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One loss step = the calls above chained on one stream behind a single descriptor (what
 * PixelContrastLoss.forward / its autograd backward bind; lib/loss/loss_contrast.py:130-147,
 * lib/loss/loss_contrast_mem.py:154-171).  All scratch is caller-owned.
 * ----------------------------------------------------------------------------------------------*/
typedef struct {
  pcl_geom g;
  /* inputs */
  const float*   embed;          /* (B,D,h,w)                                                        */
  const int64_t* labels;         /* (B,Himg,Wimg)                                                    */
  const float*   seg;            /* (B,K,h,w) or NULL                                                */
  const int64_t* predict;        /* (B,h,w) or NULL                                                  */
  const int32_t* ranks;          /* injected permutation values or NULL (device RNG)                 */
  uint64_t seed;
  int32_t normalize;
  /* contrast set: mode 0 self, 1 bank */
  int32_t mode;
  const float* segment_queue; const float* pixel_queue;
  int32_t bank_K, bank_M0, bank_M1;
  float temperature, base_temperature;
  int32_t nan_safe;
  /* scratch */
  uint16_t* keys; int32_t* chunk_pref; int32_t* counts; int32_t* plan; int32_t* anchor_meta;
  float* anchors_f32; void* anchors_bf16; float* inv_norm; float* norm_max;
  float* partials; float* rowstats; float* dpartials; float* dA;
  /* sweep selection: 0 = exact fp32 SIMT sweep (any D % 32 == 0, D <= 256); 1 = bf16 tcgen05 sweep (D == 256).
   * The tensor path reads the bank through its bf16 shadow (pcl_bank_apply / pcl_bank_shadow_rebuild). */
  int32_t precision;
  const void* shadow_bf16; int64_t shadow_rows;
  float contrast_norm_bound;
  float* row_m2;                 /* (max_samples rounded up to 256) + 512 floats (tensor path only)  */
  /* outputs */
  float* loss;                   /* 1 float                                                          */
  float* grad_embed;             /* (B,D,h,w), written by pcl_step_backward                          */
  uint32_t* sync;                /* optional: 1024 words of device memory, ZERO before the first call, owned by this step:
                                  * inter-CTA counters of the fused kernels (re-armed by the kernels themselves).
                                  * Non-NULL: pcl_step_stats runs as ONE launch (plan folded into the scan's last block).
                                  * Words [0,8): counters; word 7 != 0 turns on the diagnostic timeline in words [16,1024)
                                  * (csrc/pcl_common.cuh: per-kernel min start / max end of %globaltimer) */
} pcl_step_desc;

int pcl_step_stats(const pcl_step_desc* d, void* stream);      /* pcl_class_stats + pcl_plan_anchors      */
int pcl_step_forward(const pcl_step_desc* d, void* stream);    /* pcl_select_gather + pcl_infonce_fwd     */
int pcl_step_backward(const pcl_step_desc* d, const float* grad_loss, void* stream); /* bwd + scatter     */
/* Same backward for a grad_embed buffer the caller has already zero-filled (e.g. on a second stream, overlapped with
 * the forward: the 268 MB fill is the HBM floor of the step, the selection and sweep kernels are latency-bound): runs
 * the backward sweep and scatters only the sampled columns.  Nothing else of grad_embed is written. */
int pcl_step_backward_prezeroed(const pcl_step_desc* d, const float* grad_loss, void* stream);

/* CUDA-graph capture of a step.  The three calls above neither allocate nor synchronise, so a
 * stats -> forward -> backward sequence can be captured with cudaStreamBeginCapture / replayed with cudaGraphLaunch.
 * Kernel arguments are frozen in a captured graph, so the by-value sampling seed (pcl_step_desc.seed) would repeat:
 * call pcl_step_ranks between pcl_step_stats and pcl_step_forward and pass its output as pcl_step_desc.ranks.  It
 * draws the anchor ranks with the selection's own device RNG from (d->seed, *step_counter) and increments the
 * counter in device memory, so every replay samples a fresh anchor set (replaces torch.randperm of
 * lib/loss/loss_contrast.py:79-82 exactly like the eager device RNG does).
 *   step_counter  1 uint64 in device memory (caller-initialised, e.g. 0), read and incremented on the stream
 *   ranks         max_samples int32 */
int pcl_step_ranks(const pcl_step_desc* d, uint64_t* step_counter, int32_t* ranks, void* stream);
/* pcl_step_forward for captured sequences without the separate rank draw: the selection kernel draws the anchors from
 * *step_counter itself (same bijection and seed formula as pcl_step_ranks) and the forward's last kernel advances it. */
int pcl_step_forward_ctr(const pcl_step_desc* d, uint64_t* step_counter, void* stream);

/* Fused small-anchor step (self-contrast, no bank, tensor path, D = 256, max_samples <= 1024, no in-kernel normalise:
 * the shape of BASELINE configs[1]).  Four launches for the whole loss step instead of eleven:
 *   pcl_step_stats          label/argmax scan + totals + plan in one launch (d->sync != NULL)
 *   pcl_step_fused_select   selection + gather; anchor ranks drawn from *step_counter (captured sequences), or from
 *                           d->seed / d->ranks when step_counter is NULL
 *   pcl_step_fused_loss     ONE kernel for the InfoNCE forward and backward (logits stay in tensor memory;
 *                           csrc/pcl_infonce_fused.cu)
 *   pcl_step_fused_scatter  sum of the per-column-tile gradient partials + scatter of the A sampled rows into
 *                           d->grad_embed; advances *step_counter.
 * The dense gradient d->grad_embed must be zero outside the A sampled columns.  Two ways:
 *   full fill     pcl_step_fused_fill before the scatter, e.g. on a second stream forked AFTER the selection: the 268 MB fill
 *                 saturates HBM and multiplies the latency of every DRAM read issued next to it (scan 22 -> 60 us, selection
 *                 13 -> 54 us, profiles/r2_07_*, r2_08_*), while the L2-resident InfoNCE kernel hides behind it;
 *   sparse reset  prev_rows != NULL (1 + 2 * max_samples int32, zero before the first call, owned by the step): the buffer
 *                 persists between steps, zero-filled ONCE by the caller; the selection kernel clears exactly the entries
 *                 the previous step scattered and the scatter records its own.  The caller must not write the buffer.
 * d->sync is required.  pcl_step_fused_supported: 1 if the descriptor qualifies, else 0. */
int pcl_step_fused_supported(const pcl_step_desc* d);
int pcl_step_fused_select(const pcl_step_desc* d, const uint64_t* step_counter, const int32_t* prev_rows, void* stream);
int pcl_step_fused_loss(const pcl_step_desc* d, void* stream);
int pcl_step_fused_scatter(const pcl_step_desc* d, const float* grad_scale, uint64_t* step_counter, int32_t* prev_rows,
                           void* stream);
int pcl_step_fused_fill(const pcl_step_desc* d, void* stream);   /* zero-fill of d->grad_embed (B*D*h*w floats) */
/* Zero-fill `bytes` (multiple of 16, 16-byte aligned) with 16-byte stores on the engine's own fill kernel. */
int pcl_fill_zero(void* ptr, uint64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCL_H_ */
