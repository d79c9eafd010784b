/*
 * vosd_b200.h -- C ABI of libvosd_b200.so: the B200 (sm_100a) region pipeline of
 * VOSDetectron-style Mask R-CNN (proposals -> NMS -> FPN level assignment ->
 * multi-level RoIAlign fwd/bwd -> mask paste-back).
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless the parameter is documented "host";
 *   - plain pointers + explicit sizes + cudaStream_t, no framework types;
 *   - the caller allocates outputs and the opaque workspace (size from the matching
 *     *_workspace_bytes()); no hidden cudaMalloc, no hidden synchronisation, no global
 *     mutable state (the two vosd_debug_* test hooks only act in processes that export
 *     VOSD_B200_TEST_HOOKS=1): calls are re-entrant and only enqueue work on `stream`;
 *   - return value: VOSD_OK (0) or a negative vosd_status -- never exit()
 *     (the reference launchers print and exit(-1), roi_align_kernel.cu:135-139,283-287);
 *   - there is no CPU implementation behind any entry point.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef VOSD_B200_H_
#define VOSD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VOSD_API __attribute__((visibility("default")))
#else
#define VOSD_API
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;
#endif

typedef enum vosd_status {
    VOSD_OK = 0,
    VOSD_ERR_BAD_SHAPE = -1,     /* negative / zero / inconsistent extents                 */
    VOSD_ERR_BAD_ARG = -2,       /* NULL pointer, misaligned pointer, bad enum             */
    VOSD_ERR_UNSUPPORTED = -3,   /* legal in the reference but beyond a compiled-in limit  */
    VOSD_ERR_WORKSPACE = -4,     /* workspace NULL or smaller than *_workspace_bytes()     */
    VOSD_ERR_LAUNCH = -5         /* cudaGetLastError() != cudaSuccess after a launch       */
} vosd_status;

#define VOSD_MAX_LEVELS 8        /* FPN levels in one multi-level call                     */
#define VOSD_MAX_ANCHORS 16      /* anchors per location (A); FPN uses 3                    */
#define VOSD_MAX_TOPK 16384      /* pre-NMS top-k / NMS segment length handled in one CTA  */

VOSD_API const char* vosd_version(void);
VOSD_API const char* vosd_status_string(int status);
/* Number of kernels this library has enqueued since load (diagnostic counter). */
VOSD_API unsigned long long vosd_launch_count(void);
/* Binds the library's CUDA runtime to `device` for the calling thread (cudaSetDevice). */
VOSD_API int vosd_set_device(int device);

/* ------------------------------------------------------------------------------------ */
/* RoIAlign, single feature level.                                                        */
/* Replaces ROIAlignForwardLaucher / ROIAlignBackwardLaucher                              */
/* (lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.h:13-27; kernels               */
/* roi_align_kernel.cu:65-121 and :195-270) with the SAME argument order, so the          */
/* reference launcher can be swapped in for an A/B comparison.  Forward: within 1e-5      */
/* relative of the reference kernel built for sm_100a (the default separable kernel sums  */
/* the same products in a different order; measured max |err| 7e-7 for N(0,1) features;   */
/* every configuration outside sampling_ratio == 2, pooled width 7/14/28, C % 32 == 0 and */
/* the "staged" family are BIT-identical).  Backward: identical addends, different order. */
/*   bottom_data (N,C,H,W) fp32; bottom_rois (R,5) fp32 [batch,x1,y1,x2,y2] in input-     */
/*   image pixels; top_data (R,C,ph,pw) fp32.  sampling_ratio <= 0: adaptive grid.        */
/* ------------------------------------------------------------------------------------ */
VOSD_API int vosd_roialign_fwd(const float* bottom_data, float spatial_scale, int num_rois,
                      int height, int width, int channels,
                      int aligned_height, int aligned_width, int sampling_ratio,
                      const float* bottom_rois, float* top_data, cudaStream_t stream);

/* bottom_diff (N,C,H,W) is ACCUMULATED into: the caller zero-fills it, exactly as
 * RoIAlignFunction.backward does (functions/roi_align.py:39-40).  Pass zero_init != 0 to
 * have the library clear it on `stream` first (no pre-zeroed buffer needed). */
VOSD_API int vosd_roialign_bwd(const float* top_diff, float spatial_scale, int batch_size, int num_rois,
                      int height, int width, int channels,
                      int aligned_height, int aligned_width, int sampling_ratio,
                      const float* bottom_rois, float* bottom_diff, int zero_init,
                      cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* RoIAlign over all FPN levels in ONE launch.                                            */
/* Replaces the per-level loop + torch.cat + xform_shuffled[restore] of                   */
/* Generalized_RCNN.roi_feature_transform (lib/modeling/model_builder.py:262-303; mirrors */
/* lib_vos/vos_modeling/vos_model_builder.py:449-521).                                    */
/*   level_data/level_h/level_w/level_scale : HOST arrays of num_levels entries           */
/*       (device pointer of the (N,C,H_l,W_l) map, its extents, its spatial scale);       */
/*   rois (R,5) device; roi_level (R) device int32, index into the level arrays;          */
/*   out_index (R) device int32 or NULL: row of top_data that RoI i is written to         */
/*       (NULL = i).  With rois = concat(rois_fpn2..5) and out_index = the inverse of     */
/*       rois_idx_restore_int32 the result equals the reference's un-shuffled blob.       */
/* ------------------------------------------------------------------------------------ */
VOSD_API int vosd_roialign_ml_fwd(const float* const* level_data, const int* level_h, const int* level_w,
                         const float* level_scale, int num_levels, int channels,
                         int aligned_height, int aligned_width, int sampling_ratio,
                         int num_rois, const float* rois, const int* roi_level,
                         const int* out_index, float* top_data, cudaStream_t stream);

/* The same forward with a caller-provided workspace: the DEFAULT forward of the Python wrappers.  For the heads of
 * the reference (sampling_ratio == 2, aligned_width in {7, 14, 28}, channels % 32 == 0, <= 4 levels, 16-byte aligned
 * maps) a plan kernel derives every RoI's sample taps once into the workspace and a persistent, TMA-fed kernel
 * (csrc/roialign_rw.cuh) streams the footprint rows; every other configuration runs vosd_roialign_ml_fwd's kernels
 * (workspace unused).  batch_size = N of the (N,C,H_l,W_l) maps (bounds of the tensor maps).  Same results contract as
 * vosd_roialign_ml_fwd (within 1e-5 of the reference kernel); non-finite texels propagate as in the reference.
 * The workspace must be 256-byte aligned and at least vosd_roialign_fwd_workspace_bytes(...) (level_h / level_w: HOST
 * arrays); it is scratch, owned by the caller, and may be reused by the next call on the same stream. */
VOSD_API size_t vosd_roialign_fwd_workspace_bytes(const int* level_h, const int* level_w, int num_levels, int batch_size,
                                                  int channels, int aligned_height, int aligned_width, int num_rois);
VOSD_API int vosd_roialign_ml_fwd_ws(const float* const* level_data, const int* level_h, const int* level_w,
                                     const float* level_scale, int num_levels, int batch_size, int channels,
                                     int aligned_height, int aligned_width, int sampling_ratio,
                                     int num_rois, const float* rois, const int* roi_level,
                                     const int* out_index, float* top_data,
                                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* The same forward for CHANNELS-LAST maps: level_data[l] points to a (N, H_l, W_l, C) fp32 array (what a
 * torch.channels_last tensor of logical shape (N,C,H_l,W_l) holds), 16-byte aligned; top_data stays (R,C,ph,pw).
 * Tensor-mode TMA loads (32 channels x <= 32 texels x 1 row per box) land in shared memory in the layout the
 * consumer reads, so no warp stages or transposes anything.  Same summation order and
 * parity gate as the default NCHW kernel.  Supported: sampling_ratio == 2, aligned_width in {7, 14, 28},
 * channels % 32 == 0 (every head of the reference); anything else returns VOSD_ERR_UNSUPPORTED and the caller uses
 * vosd_roialign_ml_fwd on NCHW maps.  batch_size = N (bounds of the tensor maps). */
VOSD_API int vosd_roialign_ml_fwd_nhwc(const float* const* level_data, const int* level_h, const int* level_w,
                                       const float* level_scale, int num_levels, int batch_size, int channels,
                                       int aligned_height, int aligned_width, int sampling_ratio,
                                       int num_rois, const float* rois, const int* roi_level,
                                       const int* out_index, float* top_data, cudaStream_t stream);

/* Test hook selecting the RoIAlign kernel family, so every path stays covered by the parity tests:
 *   0 = default (separable warp-specialised forward where sampling_ratio == 2, pooled width 7/14/28 and
 *       C % 32 == 0, else the staged forward; record-based atomic-scatter backward),
 *   1 = generic un-staged kernels everywhere (also the in-kernel fallback for oversize RoIs),
 *   2 = staged kernels everywhere (bit-exact forward; adds the staged backward).
 * Returns the previous setting.  Takes effect ONLY in processes that export VOSD_B200_TEST_HOOKS=1 (the test suite
 * does); elsewhere it is a no-op, so a production process has no mutable library state. */
VOSD_API int vosd_debug_force_generic(int on);

/* level_diff[l] (N,C,H_l,W_l) accumulated into (zero_init as above, needs batch_size). */
VOSD_API int vosd_roialign_ml_bwd(const float* top_diff, float* const* level_diff, const int* level_h,
                         const int* level_w, const float* level_scale, int num_levels,
                         int batch_size, int channels,
                         int aligned_height, int aligned_width, int sampling_ratio,
                         int num_rois, const float* rois, const int* roi_level,
                         const int* out_index, int zero_init, cudaStream_t stream);

/* Channels-last twin of vosd_roialign_ml_bwd: level_diff[l] is the gradient of a torch.channels_last map, i.e.  */
/* (N, H_l, W_l, C) in memory; everything else as above (top_diff stays (R, C, ph, pw)).  One reduction per       */
/* (texel, channel) of the footprint, 128-byte transactions (lanes = channels).  Heads of the reference only:     */
/* sampling_ratio 2, pooled sizes multiples of 7, channels % 32 == 0; anything else VOSD_ERR_UNSUPPORTED (the      */
/* caller converts and uses vosd_roialign_ml_bwd).  Replaces the same ROIAlignBackwardLaucher calls                */
/* (roi_align_kernel.cu:195-290) for a channels-last backbone.                                                     */
VOSD_API int vosd_roialign_ml_bwd_nhwc(const float* top_diff, float* const* level_diff, const int* level_h,
                                       const int* level_w, const float* level_scale, int num_levels,
                                       int batch_size, int channels,
                                       int aligned_height, int aligned_width, int sampling_ratio,
                                       int num_rois, const float* rois, const int* roi_level,
                                       const int* out_index, int zero_init, cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* RPN proposal generation for all (level, image) segments in one call.                   */
/* Replaces GenerateProposalsOp.forward + proposals_for_one_image                         */
/* (lib/modeling/generate_proposals.py:20-168): per-segment top-k by score (:131-139),    */
/* bbox_transform (lib/utils/boxes.py:156-205), clip_tiled_boxes (:138-153),              */
/* _filter_boxes (generate_proposals.py:171-182), cython_nms.nms                          */
/* (lib/utils/cython_nms.pyx:37-87) and keep[:post_nms_topN] (:163-166).                  */
/* ------------------------------------------------------------------------------------ */
typedef struct vosd_rpn_level {
    const float* scores;      /* (N, A, H, W) fp32, rpn_cls_prob                         */
    const float* deltas;      /* (N, 4A, H, W) fp32, rpn_bbox_pred                       */
    int height, width;        /* H, W of this level                                       */
    int num_anchors;          /* A <= VOSD_MAX_ANCHORS                                    */
    double feat_stride;       /* 1 / spatial_scale (generate_proposals.py:18)             */
    double anchors[4 * VOSD_MAX_ANCHORS]; /* (A,4) fp64 base anchors of generate_anchors.py;*/
                                          /* shifted in fp64, rounded to fp32 (boxes.py:164)*/
} vosd_rpn_level;

/* Row capacity of every segment in the outputs below:
 *   cap = post_nms_topN > 0 ? min(post_nms_topN, m) : m,  m = min(pre_nms_topN>0 ? pre : inf, max_l A*H_l*W_l) */
VOSD_API int vosd_proposals_capacity(const vosd_rpn_level* levels /*host*/, int num_levels,
                            int pre_nms_topN, int post_nms_topN);
VOSD_API size_t vosd_generate_proposals_workspace_bytes(const vosd_rpn_level* levels /*host*/, int num_levels,
                                               int num_images, int pre_nms_topN, int post_nms_topN);
/*   levels: HOST array; im_info (N,3) device fp32 [height, width, scale];
 *   out_rois  (num_levels, N, cap, 5) fp32 [image, x1,y1,x2,y2] (rows >= count untouched = caller's fill);
 *   out_probs (num_levels, N, cap) fp32; out_count (num_levels, N) int32.
 *   nms_thresh <= 0 skips NMS (generate_proposals.py:159).  Ties between equal scores are
 *   ordered by lower flat (h,w,a) index first (the reference's order is unspecified).
 *   Any pre_nms_topN: while at most VOSD_MAX_TOPK boxes of every level enter NMS the call is three launches (cluster
 *   top-k + decode, IoU bitmask, greedy reduce); beyond that (pre_nms_topN <= 0 or > 16384 on a large level: the full
 *   argsort branch, generate_proposals.py:131-132) one streamed kernel sorts lazily in batches and keeps the greedy
 *   state in shared memory and in the outputs -- same results, O(kept * n) work, no n x n bitmask.                */
VOSD_API int vosd_generate_proposals(const vosd_rpn_level* levels, int num_levels, int num_images,
                            const float* im_info, int pre_nms_topN, int post_nms_topN,
                            float nms_thresh, float min_size,
                            float* out_rois, float* out_probs, int* out_count,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Streaming decode of EVERY anchor of one level (the `pre_nms_topN <= 0 or >= len(scores)`
 * branch, generate_proposals.py:131-132, and a standalone bbox_transform+clip):
 *   deltas (N,4A,H,W) -> boxes (N, H*W*A, 4) fp32 in (H,W,A) order, clipped to im_info.   */
VOSD_API int vosd_decode_anchors(const vosd_rpn_level* level /*host, scores unused*/, int num_images,
                        const float* im_info, float* boxes, cudaStream_t stream);

/* Optional NaN guard of GenerateProposalsOp (generate_proposals.py:62-63):
 * sets *flag (device int32, caller-zeroed) to 1 if any of the n floats is NaN.             */
VOSD_API int vosd_any_nan(const float* data, size_t n, int* flag, cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* Greedy box NMS.  Replaces box_utils.nms -> cython_nms.nms                              */
/* (lib/utils/boxes.py:329-333, lib/utils/cython_nms.pyx:37-87): fp32 IEEE arithmetic,    */
/* +1 widths, `>=` threshold, dets in any order.                                          */
/*   dets (n,5) fp32 [x1,y1,x2,y2,score]; keep (n) int64 ascending INDEX order (np.where  */
/*   semantics, :87); num_keep device int32.  n <= VOSD_MAX_TOPK.                         */
/* ------------------------------------------------------------------------------------ */
VOSD_API size_t vosd_nms_workspace_bytes(int n);
VOSD_API int vosd_nms(const float* dets, int n, float thresh, int64_t* keep, int* num_keep,
             void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* collect + distribute.  Replaces collect(), map_rois_to_fpn_levels() and distribute()   */
/* (lib/modeling/collect_and_distribute_fpn_rpn_proposals.py:91-138, lib/utils/fpn.py:    */
/* 11-28) on the outputs of vosd_generate_proposals.                                      */
/*   Groups: images are collected `images_per_group` at a time (the reference collects    */
/*   over its whole minibatch, :102-105; per-frame inference = 1).  G = N/images_per_group*/
/*   out_rois (G, post, 5)   top-`post` RoIs of the group by score, best first;           */
/*   out_count (G)           rows valid in each group;                                     */
/*   out_level (G, post)     int32 FPN level of each RoI in [k_min, k_max];                */
/*   level_count (G, k_max-k_min+1)                                                        */
/*   order (G, post)         concat over levels of the indices assigned to each level      */
/*                           (rois_fpnL = out_rois[order[level segment]]);                 */
/*   restore (G, post)       rois_idx_restore_int32 = inverse permutation of order.        */
/*   Rows >= out_count[g] of out_rois / out_level / order / restore are written as zeros   */
/*   (the outputs need no clearing by the caller).                                         */
/* ------------------------------------------------------------------------------------ */
VOSD_API size_t vosd_collect_distribute_workspace_bytes(int num_levels, int num_images, int cap,
                                               int images_per_group, int post_nms_topN);
VOSD_API int vosd_collect_distribute(const float* rois, const float* probs, const int* count,
                            int num_levels, int num_images, int cap, int images_per_group,
                            int post_nms_topN, int k_min, int k_max,
                            float canonical_scale, int canonical_level,
                            float* out_rois, int* out_count, int* out_level, int* level_count,
                            int* order, int* restore,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Level assignment + per-level split only (distribute() on caller-supplied rois; also the
 * test-time twin _add_multilevel_rois_for_test, lib/core/test.py:909-927).
 *   rois (R,5); outputs as above with G = 1, post = R.                                    */
VOSD_API int vosd_distribute(const float* rois, int num_rois, int k_min, int k_max,
                    float canonical_scale, int canonical_level,
                    int* out_level, int* level_count, int* order, int* restore,
                    cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* Mask paste-back.  Replaces the per-detection body of segm_results                      */
/* (lib/core/test.py:801-855; copy lib_vos/tools/vos_test.py:867-921) up to, not          */
/* including, the RLE encode: expand_boxes (lib/utils/boxes.py:242-258) -> int32          */
/* truncation -> cv2.resize(INTER_LINEAR) of the zero-padded (M+2)^2 mask -> > thresh ->  */
/* paste into a zero (im_h, im_w) uint8 canvas per detection.                             */
/*   masks (R,K,M,M) fp32; cls (R) int32 mask channel per detection or NULL (channel 0,   */
/*   CLS_SPECIFIC_MASK False); ref_boxes (R,4) fp32 original-image coords;                 */
/*   out (R, im_h, im_w) uint8 (every byte written); out_prob optional (R,im_h,im_w) fp32  */
/*   = the resized probabilities before thresholding (0 outside the box), may be NULL.    */
/* ------------------------------------------------------------------------------------ */
VOSD_API int vosd_paste_masks(const float* masks, const int* cls, const float* ref_boxes,
                     int num_dets, int num_classes, int mask_size, int im_h, int im_w,
                     float thresh, uint8_t* out, float* out_prob, cudaStream_t stream);

/* Same paste, with the result (also) as 1 bit per pixel: out_packed (R, ceil(im_h*im_w/8)) uint8, pixel 8j+k of a
 * detection's frame -> bit k of byte j (LSB first; the layout of vosd_pack_mask_bits and the payload of the
 * frame-sharded all-gather).  `out` (dense uint8, reference layout) may be NULL when im_h*im_w % 16 == 0: the
 * kernel then writes 1/8 of the bytes.  Written by the paste kernel itself, no second pass over the dense masks. */
VOSD_API int vosd_paste_masks_packed(const float* masks, const int* cls, const float* ref_boxes,
                                     int num_dets, int num_classes, int mask_size, int im_h, int im_w,
                                     float thresh, uint8_t* out, uint8_t* out_packed, cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* Fused paste -> COCO RLE ("next" row).  Replaces the whole per-detection */
/* body of segm_results INCLUDING `mask_util.encode(np.array(im_mask[:, :, np.newaxis],      */
/* order='F'))[0]` (lib/core/test.py:814-848; copy lib_vos/tools/vos_test.py:880-914): the   */
/* dense (im_h, im_w) canvas is never written.  pycocotools (third party, not vendored,      */
/* unpinned: README.md:63-74) algorithm = common/maskApi.c rleEncode + rleToString.          */
/*   inputs as vosd_paste_masks.  Per detection r:                                          */
/*     run_arena[run_offset[r] .. + run_count[r]]  uint32 run lengths of the column-major    */
/*       (Fortran-order) pixel sequence, first run = zeros (may be 0);                       */
/*     str_arena[str_offset[r] .. + str_len[r]]    the 'counts' string of the RLE dict       */
/*       (ASCII, not NUL-terminated);                                                        */
/*     status[r] 0 ok, 1 run arena too small, 2 string arena too small (nothing written for  */
/*       that detection; cursors still advance, so cursors[0] / cursors[1] = the capacities  */
/*       a retry needs).                                                                     */
/*   cursors: 2 x uint64 scratch, cleared by the call; arenas are caller-allocated; segments */
/*   are reserved with one atomicAdd per detection, so their ORDER in the arena varies from  */
/*   run to run while their contents do not.  Dynamic shared memory:                         */
/*   vosd_paste_rle_smem_bytes(mask_size, im_h, im_w) must be <= 220 KB (1333x800 frames:    */
/*   171 KB), else VOSD_ERR_UNSUPPORTED.                                                     */
/* ------------------------------------------------------------------------------------ */
VOSD_API size_t vosd_paste_rle_smem_bytes(int mask_size, int im_h, int im_w);
VOSD_API int vosd_paste_rle(const float* masks, const int* cls, const float* ref_boxes,
                            int num_dets, int num_classes, int mask_size, int im_h, int im_w, float thresh,
                            uint32_t* run_arena, long long run_capacity, uint8_t* str_arena, long long str_capacity,
                            unsigned long long* cursors, long long* run_offset, int* run_count,
                            long long* str_offset, int* str_len, int* status, cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* Box-head post-processing.                                */
/* vosd_bbox_transform replaces the decode of im_detect_bbox (lib/core/test.py:166-181):  */
/* box_utils.bbox_transform(boxes, deltas, weights) (lib/utils/boxes.py:156-205) followed */
/* by clip_tiled_boxes (:138-153) when clip_h, clip_w > 0.                                 */
/*   boxes (n,4), deltas (n,4k), weights HOST float[4], out (n,4k); all fp32, 16-byte      */
/*   aligned.                                                                              */
/* ------------------------------------------------------------------------------------ */
VOSD_API int vosd_bbox_transform(const float* boxes, const float* deltas, int n, int k,
                        const float* weights /*host*/, float clip_h, float clip_w,
                        float* out, cudaStream_t stream);

/* vosd_box_results replaces box_results_with_nms_and_limit (lib/core/test.py:733-797;    */
/* twin lib_vos/tools/vos_test.py:748-865) for a batch of images: per class j >= 1 keep    */
/* scores >= score_thresh (:747), greedy NMS at nms_thresh (:761-762, cython_nms           */
/* arithmetic), results in ascending proposal index; then the over-all-classes limit       */
/* (:775-784): if more than max_per_image survive, keep score >= the max_per_image-th      */
/* largest score (ties kept, so the count can exceed max_per_image).  Soft-NMS and box     */
/* voting (off by default) are not implemented.                                            */
/*   scores (N,R,K) fp32; boxes (N,R,4K) fp32 (pred_boxes, class-major per row);           */
/*   rows (N) int32 or NULL: valid proposals per image (NULL = R);                         */
/*   out_dets (N,cap,6) fp32 [x1,y1,x2,y2,score,class], class-major / proposal index       */
/*   order = np.vstack(cls_boxes[1:]) (:790); out_count (N) = true count (rows beyond cap  */
/*   are dropped); out_cls_count (N,K) int32 or NULL.  max_per_image <= 0: no limit.       */
/* ------------------------------------------------------------------------------------ */
VOSD_API size_t vosd_box_results_workspace_bytes(int num_images, int rois_per_image, int num_classes);
VOSD_API int vosd_box_results(const float* scores, const float* boxes, const int* rows, int num_images,
                     int rois_per_image, int num_classes, float score_thresh, float nms_thresh,
                     int max_per_image, int cap, float* out_dets, int* out_count, int* out_cls_count,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Dense {0,1} uint8 masks (num_masks, pixels_per_mask) -> bit-packed (num_masks, ceil(pixels/8)) uint8,
 * pixel 8j+k in bit k of byte j.  Lossless payload of the frame-sharded all-gather (the reference ships
 * whole pickled results between its per-GPU processes instead, lib/core/test_engine.py:193-200). */
VOSD_API int vosd_pack_mask_bits(const uint8_t* masks, int num_masks, long long pixels_per_mask,
                                 uint8_t* packed, cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* FlowAlign: warp a feature map by an optical-flow field ("next" row). */
/* Replaces FlowAlignForward / FlowAlignBackward                                          */
/* (lib_vos/vos_model/flow_align/src/flow_align_cuda_kernel.h:11,15; kernels              */
/* flow_align_cuda_kernel.cu:15-55 and :57-117) with the SAME argument order.             */
/*   bottom (N,C,H,W) fp32; flow (N,2,H,W) fp32, plane 0 = x displacement, plane 1 = y,   */
/*   in pixels of THIS map; top (N,C,H,W): top[n,c,h,w] = bilinear(bottom[n,c], h+fy,     */
/*   w+fx), or 0 when the sample lies outside [0,H-1) x [0,W-1) (:37-41).                 */
/* Forward: bit-identical to the reference kernel built for sm_100a (same mixed           */
/* float/double expression, :48-51).  Backward: the reference's addends (:89-112); where   */
/* neighbouring pixels hit the same texel they are summed in registers first, and the     */
/* flow gradient is summed over a channel chunk before it is added, so only the order of  */
/* the fp32 sum differs (the reference's atomics are unordered as well).                  */
/* bottomdiff / flowdiff are ACCUMULATED into; FlowAlignFunction.backward zero-fills them */
/* (functions/flow_align.py:41-43): pass zero_init != 0 to have the library clear them on */
/* `stream`.  N*C*H*W must be < 2^31 (the reference indexes with int).  flowdiff == NULL:  */
/* the flow gradient is not wanted (flow from a frozen estimator); its tap loads and        */
/* arithmetic (:95-112) are skipped.                                                       */
/* ------------------------------------------------------------------------------------ */
VOSD_API int vosd_flow_align_fwd(int batches, int height, int width, int channels, const float* bottom,
                                 const float* flow, float* top, cudaStream_t stream);
VOSD_API int vosd_flow_align_bwd(int batches, int height, int width, int channels, const float* topdiff,
                                 const float* bottom, const float* flow, float* bottomdiff, float* flowdiff,
                                 int zero_init, cudaStream_t stream);

/* All FPN levels in ONE launch: replaces the `for i in range(5): self.FlowAligns[i](hidden_states[i], flow)` */
/* loop of lib_vos/vos_modeling/vos_model_builder.py:329-335 (per level: its own map and its own, already     */
/* down-sampled, flow).  level_h / level_w and the pointer tables are HOST arrays of num_levels entries       */
/* (<= VOSD_MAX_LEVELS); every level has `batches` images and `channels` channels.                            */
VOSD_API int vosd_flow_align_ml_fwd(int num_levels, int batches, int channels, const int* level_h,
                                    const int* level_w, const float* const* bottom, const float* const* flow,
                                    float* const* top, cudaStream_t stream);
VOSD_API int vosd_flow_align_ml_bwd(int num_levels, int batches, int channels, const int* level_h,
                                    const int* level_w, const float* const* topdiff, const float* const* bottom,
                                    const float* const* flow, float* const* bottomdiff, float* const* flowdiff,
                                    int zero_init, cudaStream_t stream);
/* Test / tuning hook selecting the arithmetic of the FlowAlign kernels, so every variant stays covered:      */
/*   0 = default: the reference's expression as written (F2F conversions on the XU pipe; bit-identical),     */
/*   2 = the same double-precision operation sequence with the float->double widenings done as bit shuffles  */
/*       on the ALU pipe (bit-identical; non-finite taps take the expression as written; measured slower),   */
/*   1 = plain fp32 bilinear weights in the forward (NOT bit-identical, |err| <= 1e-6 relative).             */
/* Returns the previous setting.  Takes effect only with VOSD_B200_TEST_HOOKS=1 in the environment (see
 * vosd_debug_force_generic); a no-op elsewhere. */
VOSD_API int vosd_debug_flow_align_fast(int on);

/* ------------------------------------------------------------------------------------ */
/* Mask-IoU suppression ("next" row, SURVEY 8f rank 2, second half).  Replaces the O(R^2)  */
/* loop of nms_with_mask_iou (lib_vos/tools/vos_test.py:985-1029) over decoded full-frame   */
/* masks with AND + POPC over bit-packed masks.                                            */
/*   packed (num_masks, bytes_per_mask) uint8, 1 bit per pixel, any pixel order shared by   */
/*   all masks (vosd_paste_masks_packed or vosd_rle_to_bits), bytes_per_mask % 4 == 0,      */
/*   padding bits zero; order (num_masks) int32: mask index at each position of the         */
/*   descending-score order (vos_test.py:995-998), NULL = identity.                         */
/*   Position j is removed when an earlier surviving position i has                         */
/*   inter/(area_i + 1e-6) > iou_th or inter/(area_j + 1e-6) > iou_th in float64            */
/*   (iou_half_numpy, :953-959; test :1009).  removed (num_masks) int32 0/1 per POSITION,   */
/*   num_keep (1) int32.  num_masks <= 2048.                                                */
/* ------------------------------------------------------------------------------------ */
VOSD_API size_t vosd_mask_iou_nms_workspace_bytes(int num_masks);
VOSD_API int vosd_mask_iou_nms(const uint8_t* packed, int num_masks, long long bytes_per_mask, const int* order,
                               double iou_th, int* removed, int* num_keep, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream);
/* COCO run lengths -> bits (the mask_util.decode(segms) of vos_test.py:993, as 1 bit per pixel in the RLE's    */
/* own column-major pixel order): mask m = runs[run_offset[m] .. + run_count[m]) uint32, alternating 0-runs and  */
/* 1-runs, first run zeros; out_packed (num_masks, ceil(pixels/32)*4) uint8, every byte written.                 */
/* max_run_count = max over run_count (host value) must be <= 12000.                                             */
VOSD_API int vosd_rle_to_bits(const uint32_t* runs, const long long* run_offset, const int* run_count,
                              int num_masks, long long pixels, uint8_t* out_packed, int max_run_count,
                              cudaStream_t stream);

/* ------------------------------------------------------------------------------------ */
/* bbox_overlaps ("next" row, SURVEY 8f rank 3: first piece of the training label         */
/* assignment).  Replaces cython_bbox.bbox_overlaps (lib/utils/cython_bbox.pyx:32-73,     */
/* bound at lib/utils/boxes.py:55) and the `.argmax(axis=1)` / `.max(axis=1)` that follow  */
/* it in datasets/json_dataset.py:450-456 and roi_data/rpn.py:149-158.  The arithmetic of   */
/* the compiled .pyx (areas and union in float64, products and the division in float32):    */
/* bit-identical.                                                                          */
/*   boxes (N,4), query_boxes (K,4) fp32, 16-byte aligned; overlaps (N,K) fp32 or NULL;    */
/*   row_max (N) fp32 or NULL; row_argmax (N) int32 or NULL (first maximum, as np.argmax;  */
/*   0 and 0.0 when K == 0).                                                               */
/* ------------------------------------------------------------------------------------ */
VOSD_API int vosd_bbox_overlaps(const float* boxes, int num_boxes, const float* query_boxes, int num_query,
                                float* overlaps, float* row_max, int* row_argmax, cudaStream_t stream);

/* Box regression targets of the label assignment (rank 3, second piece): _compute_targets +      */
/* _expand_bbox_targets (lib/roi_data/fast_rcnn.py:216-260; bbox_transform_inv                     */
/* lib/utils/boxes.py:208-239) + bbox_outside_weights (:206-208) in one launch.                    */
/*   ex_rois, gt_rois (n,4) fp32, 16-byte aligned: sampled RoIs and the gt box each is assigned to; */
/*   labels (n) int32; weights HOST float[4] = MODEL.BBOX_REG_WEIGHTS; class_agnostic != 0:         */
/*   MODEL.CLS_AGNOSTIC_BBOX_REG (2 regression classes, labels clipped to 1).                       */
/*   bbox_targets, inside_weights, outside_weights (n, 4*K') fp32, K' = 2 or num_classes; every     */
/*   element is written (outside_weights may be NULL).  fp32 in NumPy's operation order; the two    */
/*   logarithms are logf (<= 1 ulp from NumPy's float32 log).                                       */
VOSD_API int vosd_bbox_targets(const float* ex_rois, const float* gt_rois, const int* labels, int num_rois,
                               int num_classes, int class_agnostic, const float* weights /*host*/,
                               float* bbox_targets, float* inside_weights, float* outside_weights,
                               cudaStream_t stream);

/* Sampling of the training RoIs (rank 3, third piece): the index selection of _sample_rois                */
/* (lib/roi_data/fast_rcnn.py:132-160) for every image of the minibatch in one launch.                     */
/*   max_overlaps, keys (num_images, stride) fp32: max overlap of every box of the image's roidb entry     */
/*   (gt rows included) and ONE UNIFORM RANDOM KEY per box; num_boxes (num_images) int32 valid boxes.      */
/*   RNG contract: npr.choice(inds, size, replace=False) of the reference = the `size` candidates with the */
/*   smallest keys, in ascending key order (ties: lower index first) -- reproducible, and the same         */
/*   distribution as the reference's permutation draw.  Foreground candidates: max_overlaps >= fg_thresh,  */
/*   at most fg_rois_per_image = round(FG_FRACTION * BATCH_SIZE_PER_IM); background: bg_thresh_lo <=       */
/*   max_overlaps < bg_thresh_hi, filling up to rois_per_image.                                            */
/*   keep_inds (num_images, rois_per_image) int32: foreground picks, then background picks, -1 beyond      */
/*   num_keep; num_fg, num_keep (num_images) int32.  rois_per_image <= VOSD_MAX_TOPK.                      */
VOSD_API int vosd_sample_rois(const float* max_overlaps, const float* keys, const int* num_boxes, int num_images, int stride,
                              int rois_per_image, int fg_rois_per_image, float fg_thresh, float bg_thresh_hi,
                              float bg_thresh_lo, int* keep_inds, int* num_fg, int* num_keep, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VOSD_B200_H_ */
