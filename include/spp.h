/* spp.h — C ABI of the B200-native selective-pose glue path (libspp.so).
 *
 * The reference (Person-Recognition-for-Pose-Estimation) has no plugin / FFI interface: its hot path
 * is a set of Python functions.  Each entry point below replaces one of them and says
 * which (paths relative to the reference root; HF = transformers/models/vitpose, the un-vendored
 * third-party processor the reference imports at training/modify_models.py:335).
 *
 * Conventions
 *   - every pointer marked DEVICE is a CUDA device pointer on the current device; HOST pointers are
 *     small parameter arrays read before the launch.  The caller owns all buffers; nothing is
 *     allocated here except through the caller-provided workspace.
 *   - all work is enqueued on `stream` (a cudaStream_t); no entry point synchronises, all of them
 *     are CUDA-graph capturable.
 *   - return value 0 = success, negative = error; spp_last_error() returns the thread-local message.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SPP_H_
#define SPP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *spp_stream_t; /* cudaStream_t */

#define SPP_OK 0
#define SPP_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define SPP_ERR_CUDA (-2)      /* a CUDA runtime call failed */
#define SPP_ERR_WORKSPACE (-3) /* workspace too small */

#define SPP_MAX_LEVELS 4

/* ABI version (bumped when a signature changes). */
int spp_abi_version(void);
const char *spp_last_error(void);
/* Number of SMs of the current device (grid sizing is a multiple of it); <0 on error. */
int spp_device_sm_count(void);

/* Process-wide CTA budgets for the two kernels that otherwise take every SM (one persistent CTA per SM): a caller that
 * runs them side by side (the pipeline: heatmap decode on 148 - R SMs, match GEMM on R) sets both before it captures its
 * CUDA graph.  max_ctas = 0 removes the limit, < 0 only queries; returns the previous value (or -1 for a bad `which`).
 * spp_match_workspace_bytes depends on the match limit: query it after setting the limit. */
#define SPP_LIMIT_HEATMAP_CTAS 0
#define SPP_LIMIT_MATCH_CTAS 1
/* Third budget, a different kind: the persistent crop kernel (spp_crop_affine*_ws / _run) normally fills every CTA slot of the
 * machine for its whole run, so a small kernel enqueued beside it (the match re-score, a late detection kernel) would wait for
 * it to end.  SPP_LIMIT_CROP_FREE_CTAS = n launches it with n CTAs fewer: n slots of 224 threads / 12.5 k registers / 45 KB stay
 * free across the SMs.  The crop is HBM-bound, a few per cent fewer CTAs do not slow it. */
#define SPP_LIMIT_CROP_FREE_CTAS 2
int spp_set_launch_limit(int which, int max_ctas);

/* ------------------------------------------------------------------ detection head + NMS ----- */

/* Replaces Head.forward, eval branch — training/yolopt/nets/nn.py:255-270 (after the per-level conv
 * stacks) incl. make_anchors (training/yolopt/util.py:85-96) and DFL (nn.py:222-225).
 *   levels[l]  DEVICE  raw map [batch, 64+nc, level_h[l], level_w[l]] fp32, NCHW contiguous
 *   out        DEVICE  [batch, 4+nc, A] fp32: (cx, cy, w, h) in pixels, sigmoid class scores
 */
int spp_head_decode(const float *const *levels, const int *level_h, const int *level_w, const float *strides,
                    int num_levels, int batch, int nc, float *out, spp_stream_t stream);

/* Workspace for the two NMS entry points below (bytes).  max_candidates bounds the per-image
 * candidate list: min(num_anchors * nc, max_candidates) candidates are kept per image (pass <= 0 for
 * "all of them").  Results are exact whenever the number of candidates of every image fits; beyond
 * that the surplus (arbitrary) candidates are dropped and out_count[b] is returned as ~kept = -(kept + 1).  With
 * nc == 1 and max_candidates <= 0 this cannot happen.  A bound of <= 512 also selects the small-footprint NMS kernel
 * (512 threads, ~15 KB of shared memory instead of 1 024 threads / 112 KB): same results, but its CTAs fit beside the
 * resident CTAs of the bandwidth-bound kernels, so a caller that knows its scene density gets the overlap. */
size_t spp_nms_workspace_bytes(int batch, int num_anchors, int nc, int max_candidates);

/* Replaces non_max_suppression(outputs, conf, iou) — training/yolopt/util.py:123-169, without its
 * wall-clock bail-out (:133-134,:166-167).
 *   pred       DEVICE  [batch, 4+nc, A] fp32 (the tensor spp_head_decode produces)
 *   out_dets   DEVICE  [batch, max_det, 6] fp32 rows (x1, y1, x2, y2, conf, cls), conf-descending;
 *                      rows >= out_count[b] are zero
 *   out_count  DEVICE  [batch] int32
 *   out_keys   DEVICE  [batch, max_det] int32 candidate key anchor*nc + cls of each kept row (may be NULL)
 * nc == 1: best class only; nc > 1: multi-label, one candidate per (anchor, class) above conf.
 * Suppression: class-aware through the reference's +cls*max_wh box offset; IoU > iou (strict), fp32,
 * no FMA contraction; equal scores: lower candidate key first.
 */
int spp_nms_decoded(const float *pred, int batch, int nc, int num_anchors, float conf_thres, float iou_thres,
                    int max_det, int max_nms, float max_wh, int max_candidates, float *out_dets, int *out_count,
                    int *out_keys, void *workspace, size_t workspace_bytes, spp_stream_t stream);

/* Fused Head.forward(eval) + non_max_suppression straight from the raw per-level maps: the decoded
 * [batch, 4+nc, A] tensor is never materialised and box channels are only read for candidates.
 * Same outputs as spp_nms_decoded. */
int spp_decode_nms(const float *const *levels, const int *level_h, const int *level_w, const float *strides,
                   int num_levels, int batch, int nc, float conf_thres, float iou_thres, int max_det, int max_nms,
                   float max_wh, int max_candidates, float *out_dets, int *out_count, int *out_keys,
                   void *workspace, size_t workspace_bytes, spp_stream_t stream);

/* Implementation choice for spp_decode_nms / spp_decode_nms_split, process-wide: 0 = three launches per call (candidate
 * scan, candidate decode, sort + NMS; default), 1 = one fused kernel (a CTA per image does all of it; fewer launches,
 * e.g. for tiny batches).  Results are identical bit for bit.  Returns the previous mode; any other argument only queries. */
int spp_decode_nms_mode(int mode);

/* The same two entry points on the head's conv outputs BEFORE the per-level torch.cat of nn.py:257
 * (SURVEY.md 8f-1): box_levels[l] DEVICE [batch, 64, level_h[l], level_w[l]] (self.box[i](x[i])) and
 * cls_levels[l] DEVICE [batch, nc, level_h[l], level_w[l]] (self.cls[i](x[i])), NCHW contiguous fp32.
 * Identical results; the concatenated copy (5 MB per 720p frame) is never made. */
int spp_head_decode_split(const float *const *box_levels, const float *const *cls_levels, const int *level_h,
                          const int *level_w, const float *strides, int num_levels, int batch, int nc, float *out,
                          spp_stream_t stream);
int spp_decode_nms_split(const float *const *box_levels, const float *const *cls_levels, const int *level_h,
                         const int *level_w, const float *strides, int num_levels, int batch, int nc, float conf_thres,
                         float iou_thres, int max_det, int max_nms, float max_wh, int max_candidates, float *out_dets,
                         int *out_count, int *out_keys, void *workspace, size_t workspace_bytes, spp_stream_t stream);

/* ------------------------------------------------------------------ embedding + gallery match - */

/* Replaces the tail of Backbone.forward — libs/net_adaface.py:334-337 (mode 0: x / ||x||, no eps)
 * and F.normalize — training/lightning/face_recognition/module.py:138 (mode 1: x / max(||x||, eps)).
 *   x DEVICE [m, dim] fp32; out DEVICE [m, dim] fp32 (may be NULL); norm DEVICE [m] fp32 (may be NULL);
 *   out_bf16 DEVICE [m, dim] bf16 bits (may be NULL) — the GEMM operand. */
int spp_l2_normalize(const float *x, int m, int dim, int mode, float eps, float *out, float *norm,
                     uint16_t *out_bf16, spp_stream_t stream);

/* fp32 [n, dim] -> bf16 [n, dim] (gallery enrolment; off the hot path). */
int spp_f32_to_bf16(const float *x, size_t count, uint16_t *out, spp_stream_t stream);

size_t spp_match_workspace_bytes(int m, int n, int dim);

/* Replaces the cosine match + top-1 — training/lightning/face_recognition/module.py:136-145
 * (F.linear(F.normalize(emb), gallery) ; max(1)) and libs/head_adaface.py:79-81, plus the north-star's
 * threshold gate (not in the reference).
 *   emb        DEVICE  [m, dim] fp32 raw embeddings (normalised here, F.normalize semantics)
 *   gallery    DEVICE  [n, dim] bf16 bits, rows unit-norm (enrolment-time normalisation), dim == 512
 *   threshold  gate: id = -1 where sim < threshold; pass NaN for "no gate"
 *   id_offset  added to every returned id (gallery shard base on a multi-GPU run)
 *   out_id     DEVICE  [m] int32;  out_sim DEVICE [m] fp32 (cosine of the winner, fp32 re-scored)
 *   out_key    DEVICE  [m] uint64 (may be NULL): (orderable(sim) << 32) | (0xFFFFFFFF - global id),
 *              so that an integer MAX all-reduce over gallery shards yields the global top-1 with the
 *              lowest index winning ties (what torch.max does).
 * The dense contraction runs as a tcgen05/TMEM bf16 GEMM with a fused per-row top-2 epilogue; the
 * surviving candidates are re-scored in fp32 so the returned id is the fp32 arg-max.
 */
int spp_match_top1(const float *emb, const uint16_t *gallery, int m, int n, int dim, float threshold,
                   int id_offset, int *out_id, float *out_sim, unsigned long long *out_key, void *workspace,
                   size_t workspace_bytes, spp_stream_t stream);

/* spp_match_top1 with the two enrolment-time facts the exactness guarantee depends on:
 *   gallery_f32   DEVICE [n, dim] fp32 rows the bf16 gallery was rounded from, or NULL.  When given, the surviving
 *                 candidates are re-scored against THESE rows: ids and similarities are those of the reference's fp32
 *                 F.linear(F.normalize(emb), gallery).max(1) (face_recognition/module.py:136-145), the bf16 copy only
 *                 steers the candidate search.  NULL: re-score against the bf16 rows (ids / sims exact for that gallery).
 *   max_row_norm  largest L2 norm of a gallery row (1 for a normalised gallery; > 1 e.g. for quirk Q3 enrolment).  It
 *                 scales the band of candidate-search scores that are re-scored: |bf16 score - fp32 score| <= 2^-8 * norm.
 * The returned id is the exact fp32 arg-max (lowest id on ties) for ANY gallery content: the GEMM epilogue keeps the best
 * two scores per (probe, gallery chunk) plus the best score it dropped, and a chunk whose dropped scores reach the
 * re-score band (near-duplicate enrolments, the crowded top of a 1M-id gallery) is re-scanned in exact fp32. */
int spp_match_top1_ex(const float *emb, const uint16_t *gallery, const float *gallery_f32, float max_row_norm, int m, int n,
                      int dim, float threshold, int id_offset, int *out_id, float *out_sim, unsigned long long *out_key,
                      void *workspace, size_t workspace_bytes, spp_stream_t stream);

/* Unpack all-reduced keys: id = -1 where sim < threshold (NaN: no gate). */
int spp_match_unpack_keys(const unsigned long long *keys, int m, float threshold, int *out_id, float *out_sim,
                          spp_stream_t stream);

/* ------------------------------------------------------------------ gallery sharded over GPUs - */

/* North-star config 3: the gallery is sharded by rows over the GPUs of one box, one process per GPU.  The
 * exchange steps that NCCL would do (all-gather of the probes, top-1 (value, index) all-reduce) are done by the match
 * kernels themselves over NVLink peer memory: every rank owns one exchange buffer, mapped by all ranks through CUDA IPC.
 * The reference has no multi-GPU inference path (its only collectives are training-time DDP, training/yolopt/main.py:57-60).
 *
 *   spp_peer_alloc   cudaMalloc + zero an exchange buffer on the current device, return its IPC handle (64 bytes, HOST)
 *   spp_peer_open    map another process's buffer from its handle (peer access enabled lazily);  spp_peer_close unmaps
 *   spp_peer_free    release a buffer obtained from spp_peer_alloc
 * The host exchanges the 64-byte handles with any transport it has (torch.distributed all_gather_object in dist.py). */
#define SPP_MAX_PEERS 16
#define SPP_IPC_HANDLE_BYTES 64
typedef struct spp_peer_group {
    int world, rank;
    int m_local;                   /* probes per rank and step — the same on every rank */
    void *buffers[SPP_MAX_PEERS];  /* DEVICE pointers valid in THIS process; buffers[rank] is this rank's own buffer */
} spp_peer_group;

size_t spp_peer_buffer_bytes(int world, int m_local, int dim);
int spp_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle_out);
int spp_peer_open(const unsigned char *handle, void **dev_ptr);
int spp_peer_close(void *dev_ptr);
int spp_peer_free(void *dev_ptr);
/* 1 if the current device can map peer memory of device `other` (cudaDeviceCanAccessPeer), 0 if not, <0 on error. */
int spp_peer_can_access(int other_device);

#define SPP_SHARDED_STAGE_PUSH 1      /* F.normalize this rank's probes, store fp32 + bf16 into every rank's buffer */
#define SPP_SHARDED_STAGE_WAIT 2      /* wait until every rank's probes of this step have landed here */
#define SPP_SHARDED_STAGE_SEARCH 4    /* tcgen05 GEMM + top-2 of ALL world * m_local probes against the local shard */
#define SPP_SHARDED_STAGE_FINALIZE 8  /* fp32 re-score; every probe's packed (sim, id) key is stored into its owner's buffer */
#define SPP_SHARDED_STAGE_REDUCE 16   /* wait for all ranks' keys of MY probes, integer max, unpack + gate; advances the step */
#define SPP_SHARDED_STAGE_ALL 31

size_t spp_sharded_match_workspace_bytes(int world, int m_local, int n_shard, int dim);

/* One step of the sharded match = the five kernels above, enqueued on `stream` (graph-capturable; no host
 * synchronisation, no NCCL).  EVERY rank of the group must enqueue the same number of steps; a rank that waits more than
 * 60 s for a peer traps (CUDA error) instead of hanging.  `stages` = SPP_SHARDED_STAGE_ALL; a subset runs only those
 * kernels (per-stage timing in bench.py — over one step all five must still run, in order, on every rank).
 *   emb        DEVICE [m_local, dim] fp32 raw embeddings of THIS rank's probes
 *   shard      DEVICE [n_shard, dim] bf16 rows [id_offset, id_offset + n_shard) of the gallery; shard_f32 / max_row_norm
 *              as in spp_match_top1_ex
 *   out_id / out_sim / out_key   DEVICE [m_local]: global top-1 of this rank's probes over ALL shards (out_key may be NULL)
 */
int spp_sharded_match_top1(const spp_peer_group *group, const float *emb, const uint16_t *shard, const float *shard_f32,
                           float max_row_norm, int n_shard, int dim, int id_offset, float threshold, int stages, int *out_id,
                           float *out_sim, unsigned long long *out_key, void *workspace, size_t workspace_bytes,
                           spp_stream_t stream);

/* ------------------------------------------------------------------ face -> person association */

/* The "selective" step between match and crop.  NOT in the reference (scripts/modify_models.py:71-76 is a
 * TODO): builder-defined, parity unpinned (restated in oracle/assoc.py).  Per frame: every
 * face detection with a matched identity picks the person detection that contains its centre and covers most
 * of it; persons picked by at least one face are emitted, in row order, as COCO boxes for the crop.
 *   face_dets    DEVICE [batch, face_cap, 6]   rows of spp_decode_nms (x1, y1, x2, y2, conf, cls)
 *   face_count   DEVICE [batch] int32;  face_ids DEVICE [batch, face_cap] int32 (identity per row, -1 = none)
 *   person_dets  DEVICE [batch, person_cap, 6]; person_count DEVICE [batch] int32
 *   out_boxes    DEVICE [batch, cap, 4] fp32 COCO (x, y, w, h), zero padding
 *   out_ident    DEVICE [batch, cap] int32 identity of each selected person (-1 padding)
 *   out_row      DEVICE [batch, cap] int32 person row of each selected person (may be NULL)
 *   out_count    DEVICE [batch] int32 selected persons per frame (<= cap)
 */
int spp_associate(const float *face_dets, const int *face_count, const int *face_ids, int face_cap,
                  const float *person_dets, const int *person_count, int person_cap, int batch, int cap,
                  float *out_boxes, int *out_ident, int *out_row, int *out_count, spp_stream_t stream);

/* ------------------------------------------------------------------ crop --------------------- */

#define SPP_CROP_HF_UDP 0   /* HF VitPoseImageProcessor (box_to_center_and_scale + get_warp_matrix) */
#define SPP_CROP_GLUONCV 1  /* training/lightning/pose_estimation/datamodule_v2.py:119-129,213-226 */

/* Replaces VitPoseImageProcessor.preprocess(images, boxes) — HF image_processing_vitpose.py:68-172,
 * 386-448 (variant 0), or the dataset crop of datamodule_v2.py (variant 1; exact bilinear, no cv2
 * fixed-point quantisation).
 *   frames     DEVICE  [num_frames, 3, frame_h, frame_w] fp32
 *   boxes      DEVICE  [p, 4] fp32 COCO (x, y, w, h) in frame pixels
 *   frame_idx  DEVICE  [p] int32
 *   mean, std  HOST    [3] fp32: out = (sample - mean[c]) / std[c]
 *   out        DEVICE  [p, 3, out_h, out_w] fp32
 * Sample = exact bilinear at the UDP source coordinate, 0 when it falls outside
 * [0, frame_w-1] x [0, frame_h-1] (scipy order=1, mode='constant').
 */
int spp_crop_affine(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                    const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                    int variant, float *out, spp_stream_t stream);

/* Same for uint8 frames [num_frames, 3, frame_h, frame_w] (what HF does when handed uint8 images: the
 * interpolated value is rounded half-up to uint8 by scipy before rescale / normalise; here the
 * interpolation runs in fp64 so the rounding agrees).  mean/std are in 0..255 units (HF folds the 1/255
 * rescale into them, image_processing_backends.py:301-305). */
int spp_crop_affine_u8(const uint8_t *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                       const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                       int variant, float *out, spp_stream_t stream);

/* The same two ops through the persistent kernels: a plan kernel computes the fp64 source map, both coordinate tables and the
 * band layout once per crop into `workspace`, and a resident stream kernel pulls (crop, channel, 64-row slab) items from a
 * ticket counter in it, fed by bulk-TMA copies of the planned tables (no per-CTA set-up, short drain).  Results are identical
 * bit for bit.  workspace: DEVICE, 16-byte aligned, spp_crop_workspace_bytes(p, out_h, out_w, frames_u8) bytes, private to
 * the call until it has finished on `stream`; NULL selects the kernels of spp_crop_affine / spp_crop_affine_u8.
 * Replaces the same reference code as spp_crop_affine (HF image_processing_vitpose.py:68-172, 386-448). */
size_t spp_crop_workspace_bytes(int p, int out_h, int out_w, int frames_u8);
/* Which implementation a call WITH a workspace runs: 0 = always one CTA per work item, 1 = automatic (the persistent kernels for
 * fp32 frames up to 4 096 crops; the per-item kernel beyond, and for uint8 frames, which are bound by their instruction stream), 2 = always the persistent kernels.  Process-wide; mode < 0 only queries; returns the previous value. */
int spp_crop_policy(int mode);
int spp_crop_affine_ws(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                       const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                       int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream);
int spp_crop_affine_u8_ws(const uint8_t *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                          const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                          int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream);

/* Every combination in one call: frames fp32 (frames_u8 = 0) or uint8, output fp32 or bf16 (out_bf16 = 1: the fp32 result rounded
 * to nearest even, `[p, 3, out_h, out_w]` of 2-byte elements — for a pose backbone under bf16 autocast it halves the bytes this
 * HBM-bound op writes), workspace NULL (one CTA per work item) or from spp_crop_workspace_bytes (persistent kernels), planned = 1
 * when spp_crop_plan has already run on that workspace for these boxes. */
int spp_crop_affine_ex(const void *frames, int frames_u8, int num_frames, int frame_h, int frame_w, const float *boxes,
                       const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std, int variant,
                       void *out, int out_bf16, void *workspace, size_t workspace_bytes, int planned, spp_stream_t stream);

/* The two halves of spp_crop_affine*_ws as separate calls, same arguments and workspace: the plan reads only the boxes, so it
 * can be enqueued early (SelectivePosePipeline runs it beside the heatmap decode); the run must follow ITS plan on the device
 * (every run consumes the ticket counter its plan reset; a run on a workspace without a plan for this crop count traps, and
 * spp_crop_policy must not change between a plan and its run).  frames_u8 selects the table format of the uint8 kernels. */
int spp_crop_plan(int frames_u8, int num_frames, int frame_h, int frame_w, const float *boxes, const int *frame_idx, int p,
                  int out_h, int out_w, int variant, void *workspace, size_t workspace_bytes, spp_stream_t stream);
int spp_crop_affine_run(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                        const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                        int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream);
int spp_crop_affine_u8_run(const uint8_t *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                           const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                           int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream);

/* ------------------------------------------------------------------ heatmap decode ----------- */

#define SPP_DECODE_DARK 0        /* HF post_process_pose_estimation: argmax + DARK + UDP back-projection */
#define SPP_DECODE_SOFTARGMAX 1  /* live module: training/lightning/pose_estimation/module.py:237-296,534-546 */
#define SPP_DECODE_QUARTER 2     /* gluoncv get_final_preds as called at pose_estimation/module_v2.py:214-222 */

#define SPP_DECODE_FLAG_SCALE_SCORE 1 /* softargmax: score *= clamp(sqrt(box area)/96, .5, 2) (module.py:287-294) */
#define SPP_DECODE_FLAG_BACKPROJECT 2 /* softargmax: x = kx*(x2-x1)+x1 ... (module.py:543-544); else normalised */
#define SPP_DECODE_FLAG_CENTER_SCALE 4 /* DARK / QUARTER: `boxes` rows are (center_x, center_y, scale_x, scale_y) as in
                                          HF keypoints_from_heatmaps(heatmaps, center, scale) (scale = size/200*1.25) and
                                          gluoncv get_final_preds(heatmaps, center, scale) (scale in pixels) */

#define SPP_DECODE_FLAG_HF_F32_INDEX 8 /* DARK: reproduce HF's float32 flat tap index (image_processing_vitpose.py:248-257:
                                          `index += stride * arange` adds in place into a float32 array).  Exact for the first
                                          2^24 / ((w+2)(h+2)) maps of a call (5 084 maps = 299 crops x 17 joints at 64x48);
                                          beyond that HF reads its 7 taps 1-2 padded columns off.  Default (flag clear):
                                          the taps at their true positions for every map.  Reference quirk Q6. */

/* Replaces flip-test averaging (module.py:473-484 with the correct channel swap of
 * "module copy.py":465-472 / HF modeling_vitpose.py:80-117) + the three decode variants, in one pass
 * over the heatmaps.
 *   hm          DEVICE  [p, k, h, w] fp32 heatmaps of the crop
 *   hm_flipped  DEVICE  [p, k, h, w] fp32 heatmaps of the MIRRORED crop, not flipped back (NULL: no flip test)
 *   perm        DEVICE  [k] int32 left/right channel permutation (NULL: identity — module_v2.py:201)
 *   boxes       DEVICE  [p, 4] fp32: COCO (x, y, w, h) for DARK / QUARTER, (x1, y1, x2, y2) for SOFTARGMAX;
 *                       NULL: coordinates stay in heatmap space (DARK/QUARTER) or normalised (SOFTARGMAX)
 *   kernel      DARK Gaussian modulation kernel size (HF default 11); odd, 3..17
 *   crop_h/w    model input size the boxes were cropped to (256 x 192): fixes the aspect ratio
 *   keypoints   DEVICE  [p, k, 2] fp32; scores DEVICE [p, k] fp32; argmax DEVICE [p, k] int32 (flat y*w+x)
 */
int spp_heatmap_decode(const float *hm, const float *hm_flipped, const int *perm, int p, int k, int h, int w,
                       const float *boxes, int mode, int flags, int kernel, int crop_h, int crop_w,
                       float *keypoints, float *scores, int *argmax, spp_stream_t stream);

/* The same decode on bf16 heatmaps (a ViTPose decoder that emits bf16 halves the bytes of this HBM-bound
 * pass — HF modeling_vitpose.py:120-145 is where the heatmaps are produced).  Elements are widened to fp32 on load (exact)
 * and every operation after that is the fp32 kernel's: results equal spp_heatmap_decode applied to the widened maps, bit
 * for bit.  Against the reference run on the ORIGINAL fp32 maps the bf16 rounding itself moves ~1.5 % of the arg-max
 * indices and keypoints by up to 0.1 px (tools/study_bf16_heatmaps.py), so fp32 stays the default input. */
int spp_heatmap_decode_bf16(const uint16_t *hm, const uint16_t *hm_flipped, const int *perm, int p, int k, int h, int w,
                            const float *boxes, int mode, int flags, int kernel, int crop_h, int crop_w,
                            float *keypoints, float *scores, int *argmax, spp_stream_t stream);

/* ------------------------------------------------------------------ pose results ------------- */

/* Replaces the result loop of PoseEstimationModule.validation_step, training/lightning/pose_estimation/
 * module.py:534-549: COCO keypoint rows (x, y, v) and the instance score.
 *   keypoints   DEVICE [p, k, 2] fp32: normalised (kx, ky) when boxes_xyxy is given, image pixels otherwise
 *   scores      DEVICE [p, k] fp32
 *   boxes_xyxy  DEVICE [p, 4] fp32 (x1, y1, x2, y2) or NULL:  x = kx * (x2 - x1) + x1,  y = ky * (y2 - y1) + y1
 *   keypoint_thresh  v = 2 if score > thresh else 1 (module.py:172 keypoint_thresh, default 0.3)
 *   out_keypoints        DEVICE [p, k, 3] fp32 (x, y, v) — the "keypoints" list of a COCO result, row-major
 *   out_instance_score   DEVICE [p] fp32 mean of the keypoint scores (may be NULL)
 */
int spp_pose_results(const float *keypoints, const float *scores, const float *boxes_xyxy, int p, int k,
                     float keypoint_thresh, float *out_keypoints, float *out_instance_score, spp_stream_t stream);

/* Object keypoint similarity of explicit (prediction, ground truth) pairs: pycocotools COCOeval.computeOks
 * (third-party, not in the reference tree; driven from module.py:598-615).  PARITY UNPINNED.
 *   pred           DEVICE [p, k, pred_stride] fp32, pred_stride 2 or 3 (x, y[, v])
 *   gt             DEVICE [p, k, 3] fp32 (x, y, v);  gt_boxes_xywh DEVICE [p, 4] or NULL (used when no joint is labelled)
 *   gt_area        DEVICE [p] fp32;  sigmas DEVICE [k] fp32 (COCO_SIGMAS, datamodule.py:36-40)
 *   out_oks        DEVICE [p] fp32
 */
int spp_pose_oks(const float *pred, int pred_stride, const float *gt, const float *gt_boxes_xywh, const float *gt_area,
                 const float *sigmas, int p, int k, float *out_oks, spp_stream_t stream);

/* ------------------------------------------------------------------ detection evaluation ----- */

/* Replaces compute_metric(output, target, iou_v) — training/yolopt/util.py:99-120 — for a whole batch: which detections are
 * true positives at each IoU threshold (the rows the reference's test loop collects, training/yolopt/main.py:210-229).
 *   dets          DEVICE [batch, det_cap, 6] rows of spp_nms_decoded / spp_decode_nms;  det_count DEVICE [batch] int32
 *   targets       DEVICE [batch, target_cap, 5] fp32 (cls, x1, y1, x2, y2) in the detections' pixel frame, zero padded
 *   target_count  DEVICE [batch] int32
 *   iou_v         HOST   [n_iou] fp32 thresholds (the reference: linspace(0.5, 0.95, 10)); n_iou <= 16
 *   correct       DEVICE [batch, det_cap, n_iou] uint8 (0 / 1); rows >= det_count[b] are 0
 * IoU = inter / (area_label + area_det - inter + 1e-7) in fp32, operations in the reference's order: bit-exact. */
int spp_det_match_targets(const float *dets, const int *det_count, int det_cap, const float *targets, const int *target_count,
                          int target_cap, const float *iou_v, int n_iou, int batch, unsigned char *correct, spp_stream_t stream);

size_t spp_det_ap_workspace_bytes(int n, int n_iou, int nc_max);

/* Replaces compute_ap(tp, conf, output, target) — training/yolopt/util.py:225-300 (plots excluded; `smooth` :172-177).
 *   tp          DEVICE [n, n_iou] uint8   true-positive matrix of all detections of the evaluation set
 *   conf        DEVICE [n] fp32;  pred_cls DEVICE [n] fp32 (class ids as the detection rows carry them)
 *   target_cls  DEVICE [nt] fp32 class id of every label;  classes are integers in [0, nc_max)
 *   eps         the reference's 1e-16
 *   out_classes      DEVICE [nc_max] int32   numpy.unique(target) (ascending; -1 padding);  out_num_classes DEVICE [1]
 *   out_ap           DEVICE [nc_max, n_iou] fp64   AP per (class, threshold), rows in out_classes order
 *   out_class_stats  DEVICE [nc_max, 4] fp64       tp, fp, precision, recall at the max-F1 confidence
 *   out_summary      DEVICE [6] fp64               m_pre, m_rec, map50, mean_ap, max-F1 index (0..999), number of classes
 * fp64 throughout, numpy's summation orders and numpy.interp's branches reproduced (see det_metrics.cu). */
int spp_det_average_precision(const unsigned char *tp, const float *conf, const float *pred_cls, int n, const float *target_cls,
                              int nt, int n_iou, int nc_max, double eps, int *out_classes, int *out_num_classes, double *out_ap,
                              double *out_class_stats, double *out_summary, void *workspace, size_t workspace_bytes,
                              spp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPP_H_ */
