/*
 * osr.h - C ABI of libosr_sm100a.so: the B200-native RoI hot path of Openset-RCNN.
 *
 * One shared library, loaded with ctypes (Python host, like the reference) or linked from C/C++.
 * No torch / pybind types cross this boundary: plain device pointers, sizes, strides, scalars.
 *
 * Conventions
 *   - every entry point returns int: 0 = OK, >0 = cudaError_t of a failed launch / API call,
 *     <0 = argument or shape error (OSR_E_*).  osr_last_error() returns a thread-local message.
 *   - the library never allocates or frees device memory and keeps no global mutable state:
 *     the caller owns all outputs and the workspace (size from the matching *_workspace()).
 *   - every call is asynchronous on the cudaStream_t passed as `stream` (void*), never syncs,
 *     never touches the default stream implicitly; all exports are re-entrant.
 *   - pointers are DEVICE pointers unless the name starts with h_ (host).
 *   - no `device` argument: every entry point launches on the device that OWNS its output / workspace pointer
 *     (cudaPointerGetAttributes; the caller's current device is restored on return), so tensors on cuda:1 work while
 *     the current device is cuda:0 and from autograd worker threads; `stream` must belong to that device.
 *   - fp32 everywhere at the boundary (the reference has no AMP); boxes are xyxy absolute pixels.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * Yifei-Y/Openset-RCNN checkout).
 */
#ifndef OSR_H_
#define OSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSR_ABI_VERSION 1
#define OSR_MAX_LEVELS 8

#define OSR_E_ARG (-1)       /* null pointer / negative size / unsupported value            */
#define OSR_E_SHAPE (-2)     /* size outside what the kernels support (message says which)  */
#define OSR_E_WORKSPACE (-3) /* workspace too small                                         */

int osr_version(void);
const char* osr_last_error(void);
/* Number of kernel launches issued by this process since the last call to osr_reset_launch_count(). */
long long osr_launch_count(void);
void osr_reset_launch_count(void);
/* Kernel-variant switches for A/B measurements (bench.py, tools/): NOT part of the drop-in surface.  Each key starts
 * at its default (0 = the shipped kernel), or at the value of the environment variable OSR_TUNE_<KEY> read ONCE when
 * the library is loaded - no entry point calls getenv.  Returns the previous value, or OSR_E_ARG for an unknown key.
 *   OSR_TUNE_BWD_VARIANT   0 register accumulators + packed fp32x2 FMAs, two staging buffers per warp (shipped) | 1 the same
 *                          with one staging buffer | 2 shared-memory accumulators (round-1 kernel) | 3 pixel-per-thread kernel
 *   OSR_TUNE_FWD_VARIANT   0 default | 1 opt-in TMA-tiled NCHW kernel | 2 no prep records | 4 persistent channels_last kernel
 *                          | 5 one footprint row per row-loop iteration (round-1 loop; the default folds two)
 *                          | 6 ring cut into up to 12 row stages for narrow footprints (default 6; measured 0.5 % slower)
 *   OSR_TUNE_PLN_VARIANT   0 encoder GEMM on fp32 operands (tcgen05 kind::tf32, no cast pass; shipped) | 1 bf16 copies (kind::f16)
 *   OSR_TUNE_RPN_VARIANT   0 default | see csrc/rpn_select_decode.cu
 *   OSR_TUNE_BWD_SPLIT     0 shipped: for batches of <= 12 images the two coarsest levels launch two CTAs per 16x16 tile of
 *                          the ROIAlign backward (one 128-channel slab each) that only share DENSE tiles (>= 30 RoIs; the
 *                          sibling leaves a sparse tile after the RoI scan), larger batches one CTA per tile | -1 one CTA
 *                          per tile | > 0: hex digit l (finest level first) = log2 of level l's CTAs per tile, bits 20-27 =
 *                          dense-tile threshold in RoIs, 0 = static split (measurements: csrc/roi_align_bwd.cu fill_bwd) */
#define OSR_TUNE_BWD_VARIANT 0
#define OSR_TUNE_FWD_VARIANT 1
#define OSR_TUNE_PLN_VARIANT 2
#define OSR_TUNE_RPN_VARIANT 3
#define OSR_TUNE_NMS_VARIANT 4
#define OSR_TUNE_BWD_SPLIT 5
#define OSR_TUNE_COUNT 6
int osr_set_tuning(int key, int value);
int osr_get_tuning(int key);

/* ------------------------------------------------------------------------------------------
 * (1) CF-RPN proposal stage
 * Replaces ClsFreeRPN.predict_proposals + _decode_proposals
 *   openset_rcnn/modeling/proposal_generator/classification_free_rpn.py:558-610
 * and find_top_rpn_proposals steps 1-3 (top-k, gather, finite check, clip, small-box filter)
 *   openset_rcnn/modeling/find_top_proposals.py:63-110
 * Decode is detectron2 Box2BoxTransformLinear(normalize_by_size=True).apply_deltas, applied only
 * to the selected anchors (select-then-decode is bit-identical: decode is elementwise).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* deltas;    /* element (n, i, c) at deltas[n*delta_stride_n + i*delta_stride_a + c*delta_stride_c] */
  const float* scores;    /* element (n, i)    at scores[n*score_stride_n + i*score_stride_a]                    */
  const float* anchors;   /* (num_anchors, 4) contiguous xyxy; NULL => `deltas` already holds decoded xyxy boxes  */
  int64_t num_anchors;    /* Hi*Wi*A */
  int64_t delta_stride_n, delta_stride_a, delta_stride_c;
  int64_t score_stride_n, score_stride_a;
} osr_rpn_level_t;

/* Kmax = sum over levels of min(num_anchors, pre_nms_topk). */
int64_t osr_rpn_kmax(const osr_rpn_level_t* h_levels, int num_levels, int pre_nms_topk);
size_t osr_rpn_select_decode_workspace(const osr_rpn_level_t* h_levels, int num_levels, int num_images,
                                       int pre_nms_topk);
/*
 * Outputs (all (N, Kmax)-padded, valid prefix per image):
 *   out_boxes  (N, Kmax, 4) fp32  clipped boxes; per image = levels concatenated, each level score-descending,
 *                                 ties by lower anchor index; non-finite and empty boxes compacted out
 *   out_scores (N, Kmax)    fp32
 *   out_level  (N, Kmax)    int32 FPN level id of each kept proposal
 *   out_index  (N, Kmax)    int32 flat anchor index inside its level (test / debugging aid)
 *   out_counts (N, L+2)     int32 [kept per level ..., kept total, flags]; flags bit0 = a selected box or score
 *                                 was non-finite (the reference raises FloatingPointError when training)
 * image_hw: (N, 2) int32 device array of un-padded (h, w)  (find_top_proposals.py:91,105).
 */
int osr_rpn_select_decode(const osr_rpn_level_t* h_levels, int num_levels, int num_images, int pre_nms_topk,
                          float min_box_size, const int32_t* image_hw, float* out_boxes, float* out_scores,
                          int32_t* out_level, int32_t* out_index, int32_t* out_counts, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2) NMS
 * Replaces detectron2.layers.batched_nms -> torchvision.ops.nms at
 *   openset_rcnn/modeling/roi_heads/osrcnn_fast_rcnn.py:135, softmax_classifier.py:93 and :154,
 *   and the nominal-mode site openset_rcnn/modeling/find_top_proposals.py:112.
 * Segments are the units inside which suppression applies (image x category).  IoU arithmetic is
 * the one of torchvision's CUDA kernel: inter / (fma(bw, bh, rn(aw*ah)) - inter) > (float)thr.
 * ------------------------------------------------------------------------------------------ */
size_t osr_nms_workspace(int64_t total_boxes, int num_segments, int max_segment_len);
/*
 *   boxes (T,4) fp32 xyxy, scores (T) fp32, total_boxes = T (the same T given to osr_nms_workspace)
 *   seg_begin (S), seg_len (S) int32 DEVICE arrays: segment s = boxes[seg_begin[s] : seg_begin[s] + seg_len[s]]
 *       (device-resident so that counts produced by osr_rpn_select_decode never visit the host)
 *   max_segment_len: host-side upper bound on any seg_len (in-kernel sort: <= 16384; presorted: <= 65536)
 *   presorted != 0: every segment is already score-descending (stable) - the sort is skipped
 *   keep_idx   (T) int64: keep_idx[seg_begin[s] + j], j < keep_counts[s] = kept boxes of segment s as indices
 *                         RELATIVE to seg_begin[s], in score-descending (stable: ties by lower index) order
 *   keep_counts (S) int32
 *   keep_mask  (T) uint8 or NULL: 1 for kept boxes (indexed like boxes), 0 for suppressed ones inside segments
 *   iou_threshold >= 1 (the reference's shipped TEST thresholds): this arithmetic never yields an IoU above 1.0f, so
 *       nothing is suppressed and the call reduces to the stable sort (no mask, no sweep) - same result, bit for bit
 */
int osr_nms_segmented(const float* boxes, const float* scores, int64_t total_boxes, const int32_t* seg_begin,
                      const int32_t* seg_len, int num_segments, int max_segment_len, float iou_threshold,
                      int presorted, int64_t* keep_idx, int32_t* keep_counts, uint8_t* keep_mask, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (3) FPN level assignment + multi-level ROIAlignV2 forward / backward
 * Replaces detectron2 ROIPooler.forward (assign_boxes_to_levels + per-level torchvision roi_align,
 * aligned=True) called at openset_rcnn/modeling/roi_heads/osrcnn_roi_heads.py:306 (built :108-113),
 * and its autograd backward (train.py:145).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  float* data;                 /* (N, C, H, W) logical; element (n,c,y,x) at data[n*sN + c*sC + y*sH + x*sW] */
  int64_t sN, sC, sH, sW;      /* element strides: NCHW-contiguous or channels_last both accepted */
  int32_t H, W;
  float scale;                 /* 1/stride of the level */
} osr_feat_level_t;

/*
 *   rois (M,5) fp32 [image index, x1, y1, x2, y2] (detectron2 convert_boxes_to_pooler_format)
 *   out  (M, C, P, P) fp32, C-major (box_head.fc1 weight order)
 *   out_level (M) int32 = clamp(floor(canonical_level + log2(sqrt(area)/canonical_box_size + 1e-8)), min, max) - min
 *   sampling_ratio 0 = adaptive ceil(roi/P) grid; aligned must be 1 (ROIAlignV2)
 */
size_t osr_roi_align_fwd_workspace(int M);
/* workspace (optional, may be NULL): lets the library order the RoIs by (image, level, y band) so that concurrently
 * processed RoIs share feature rows in L2; results do not depend on it. */
int osr_roi_align_fwd(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, const float* rois,
                      int M, int P, int sampling_ratio, int aligned, int canonical_box_size, int canonical_level,
                      int min_level, float* out, int32_t* out_level, void* workspace, size_t workspace_bytes,
                      void* stream);

size_t osr_roi_align_bwd_workspace(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, int M);
/*
 *   grad_out (M, C, P, P); rois as above and image-major (all RoIs of image n are contiguous);
 *   roi_batch_offsets (N+1) int32 device: RoIs of image n = [off[n], off[n+1])
 *   h_grad_levels[l].data: dense (N,C,H,W) gradient, FULLY written (zeros where untouched); deterministic
 *   (gather formulation, no atomics, fixed accumulation order).
 */
int osr_roi_align_bwd(const osr_feat_level_t* h_grad_levels, int num_levels, int num_images, int C,
                      const float* grad_out, const float* rois, const int32_t* roi_batch_offsets, int M, int P,
                      int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The same in two calls.  osr_roi_align_bwd's first kernel (per-RoI level, footprint box and weight rows into the workspace,
 * ~11 us for 8192 RoIs) depends on the RoIs and the map GEOMETRY only - not on grad_out - so a training step can run it as
 * soon as the RoIs are sampled, on another stream, while the forward / loss kernels run:
 *   osr_roi_align_bwd_prepare  fills the workspace (h_levels: shapes / strides / scales as for the gradient maps; data unused);
 *   osr_roi_align_bwd_prepared runs the gather on a workspace prepared for the SAME rois / offsets / geometry.
 * prepare + prepared == osr_roi_align_bwd, bit for bit. */
int osr_roi_align_bwd_prepare(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, const float* rois,
                              const int32_t* roi_batch_offsets, int M, int P, int sampling_ratio, int aligned,
                              int canonical_box_size, int canonical_level, int min_level, void* workspace,
                              size_t workspace_bytes, void* stream);
int osr_roi_align_bwd_prepared(const osr_feat_level_t* h_grad_levels, int num_levels, int num_images, int C,
                               const float* grad_out, const float* rois, const int32_t* roi_batch_offsets, int M, int P,
                               int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Layout staging for NCHW callers (the reference's backbone emits NCHW maps): tiled transposes between
 * (N, C, H*W) and (N, H*W, C) contiguous fp32 buffers.  osr_b200.poolers.ROIPooler uses them to run the channels_last
 * kernels on NCHW maps (forward: maps in; backward: gradient maps out). */
int osr_nchw_to_nhwc(const float* src, float* dst, int N, int C, int64_t HW, void* stream);
int osr_nhwc_to_nchw(const float* src, float* dst, int N, int C, int64_t HW, void* stream);

/* ------------------------------------------------------------------------------------------
 * (3b) Box-head fully connected layers on tensor cores
 * Replaces detectron2 FastRCNNConvFCHead.forward (flatten -> fc1 -> ReLU -> fc2 -> ReLU), called at
 *   openset_rcnn/modeling/roi_heads/osrcnn_roi_heads.py:308  (built :119-121; cfg ROI_BOX_HEAD NUM_FC 2, FC_DIM 1024)
 * osr_roi_align_fwd_bf16: same as osr_roi_align_fwd but the pooled tile is written as bf16 (round to nearest even), C-major
 *   (M, C, P, P) - the A operand of fc1 - instead of fp32; dense channels_last maps, C % 8 == 0, C <= 256.
 * osr_linear_bf16_fwd:  out[R, N] = act(A[R, K] . W[N, K]^T + bias), A / W bf16 row-major, fp32 accumulate (tcgen05 + TMEM),
 *   bias fp32 or NULL, relu 0 / 1, out bf16 (out_is_bf16 = 1) or fp32; K % 64 == 0, N % 256 == 0.
 * osr_cast_bf16: fp32 -> bf16 copy (weights, once per optimizer step), n % 4 == 0.
 * ------------------------------------------------------------------------------------------ */
int osr_roi_align_fwd_bf16(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, const float* rois,
                           int M, int P, int sampling_ratio, int aligned, int canonical_box_size, int canonical_level,
                           int min_level, void* out_bf16, int32_t* out_level, void* workspace, size_t workspace_bytes,
                           void* stream);
int osr_linear_bf16_fwd(const void* a_bf16, const void* w_bf16, const float* bias, int R, int K, int N, int relu, void* out,
                        int out_is_bf16, void* stream);
int osr_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * (4) PLN prototype loss (COS distance) forward / backward
 * Replaces PLN.loss lines 134,137-187 of
 *   openset_rcnn/modeling/roi_heads/prototype_learning_network.py
 * (everything after the encoder nn.Linear) and its autograd backward (closed form).
 * labels are already id-mapped: [0,K) = known class, anything else is ignored (line 146-151).
 * r_norm = normaliser (the reference uses the number of sampled RoIs R, line 187);
 * center_weight = 1 (reference) or world_size for the gathered multi-GPU variant.
 * ------------------------------------------------------------------------------------------ */
size_t osr_pln_workspace(int R, int D, int K, int reps_per_class);
/* distance_type = MODEL.PLN.DISTANCE_TYPE (prototype_learning_network.py:156-161,171-176,210-215), between the UNIT
 * embedding and the UNIT prototypes: COS 1 - <a, b> (both shipped configurations), L1 = torch.cdist(a, b, p=1),
 * L2 = torch.cdist(a, b) (computed as sqrt(sum (a-b)^2); torch's mm formulation differs from it by rounding only). */
#define OSR_PLN_DIST_COS 0
#define OSR_PLN_DIST_L1 1
#define OSR_PLN_DIST_L2 2
/*
 * saved-for-backward outputs:
 *   emb_inv_norm (R) fp32, rep_inv_norm (K*rpc) fp32,
 *   intra_rep (R) int32: rep index of the own-class minimum if the intra hinge is active else -1
 *   inter_rep (R) int32: rep index of the nearest other-class rep if the inter hinge is active else -1
 *   center_rep (K*rpc) int32: nearest other-class rep if the separation hinge is active else -1
 *   saved_dist (2*R + K*rpc) fp32: [intra distance (R) | inter distance (R) | separation distance (K*rpc)] - the L2
 *     backward divides by them; REQUIRED for L2 when a backward follows, may be NULL otherwise
 *   loss_terms (4) fp32: [loss, intra_sum, inter_sum, center_sum]
 */
int osr_pln_loss_fwd(const float* emb, const float* reps, const int64_t* labels, const float* ious, int R, int D,
                     int K, int reps_per_class, int distance_type, float alpha, float beta, float loss_weight,
                     float iou_threshold, float r_norm, float center_weight, float* loss_terms, float* emb_inv_norm,
                     float* rep_inv_norm, int32_t* intra_rep, int32_t* inter_rep, int32_t* center_rep, float* saved_dist,
                     void* workspace, size_t workspace_bytes, void* stream);
/*
 *   grad_loss (1) device fp32 (upstream gradient of the scalar loss)
 *   grad_emb (R, D) fully written (zeros for non-foreground rows); grad_reps (K*rpc, D)
 */
int osr_pln_loss_bwd(const float* emb, const float* reps, const int64_t* labels, const float* emb_inv_norm,
                     const float* rep_inv_norm, const int32_t* intra_rep, const int32_t* inter_rep,
                     const int32_t* center_rep, const float* saved_dist, const float* grad_loss, int R, int D, int K,
                     int reps_per_class, int distance_type, float loss_weight, float r_norm, float center_weight,
                     float* grad_emb, float* grad_reps, void* workspace, size_t workspace_bytes, void* stream);
/* Forward and closed-form backward of the loss in one call (grad_loss: device scalar, usually 1): same outputs as
 * osr_pln_loss_fwd followed by osr_pln_loss_bwd, bit for bit, in three launches (d loss / d emb is written by the row
 * kernel itself).  Used by the training step when the caller does not route the loss through autograd. */
int osr_pln_loss_fwd_bwd(const float* emb, const float* reps, const int64_t* labels, const float* ious, const float* grad_loss,
                         int R, int D, int K, int reps_per_class, int distance_type, float alpha, float beta,
                         float loss_weight, float iou_threshold, float r_norm, float center_weight, float* loss_terms,
                         float* emb_inv_norm, float* rep_inv_norm, int32_t* intra_rep, int32_t* inter_rep,
                         int32_t* center_rep, float* saved_dist, float* grad_emb, float* grad_reps, void* workspace,
                         size_t workspace_bytes, void* stream);
/* The same call in two phases on the SAME buffers: phase 1 = the row launch (loss partials, saved state, d loss / d emb),
 * phase 2 = the prototype-gradient launches + the loss reduction (loss_terms, grad_reps).  d loss / d emb - what the rest of
 * the backward pass (encoder, box head, ROIAlign) waits for - is complete after phase 1; phase 2 is an independent branch of
 * the backward graph, and the training step runs it on a side stream next to the ROIAlign backward.  Phase 1 then phase 2 ==
 * osr_pln_loss_fwd_bwd, bit for bit. */
int osr_pln_loss_fwd_bwd_phase(int phase, const float* emb, const float* reps, const int64_t* labels, const float* ious,
                               const float* grad_loss, int R, int D, int K, int reps_per_class, int distance_type, float alpha,
                               float beta, float loss_weight, float iou_threshold, float r_norm, float center_weight,
                               float* loss_terms, float* emb_inv_norm, float* rep_inv_norm, int32_t* intra_rep,
                               int32_t* inter_rep, int32_t* center_rep, float* saved_dist, float* grad_emb, float* grad_reps,
                               void* workspace, size_t workspace_bytes, void* stream);


/*
 * PLN encoder on tensor cores: emb[R,E] = x[R,F] . W[E,F]^T + bias   (prototype_learning_network.py:133, nn.Linear)
 * bf16 operands (x and W are cast from fp32 into the workspace), fp32 accumulation in TMEM (tcgen05.mma), fp32 output.
 * F must be a multiple of 64, E a multiple of 128 (1024 / 256 in the reference).  bias may be NULL.
 */
size_t osr_pln_encode_workspace(int R, int F, int E);
int osr_pln_encode_fwd(const float* x, const float* W, const float* bias, int R, int F, int E, float* emb,
                       void* workspace, size_t workspace_bytes, void* stream);
/*
 * The same GEMM FUSED with the all-gather of the embeddings that the multi-GPU (gathered) PLN loss needs: the epilogue
 * stores each output tile straight into every rank's (world*R, E) fp32 buffer at rows [rank*R, rank*R + R) - its own
 * and the peers', which the caller maps into this process (CUDA IPC / torch symmetric memory; NVLink P2P stores).
 * h_peer_buffers: HOST array of `world` device addresses, entry r = rank r's buffer as seen from this process.
 * multicast_buffer: the NVLS multicast mapping of the same buffer (0 if none): one multimem.st per 16 bytes, replicated
 * by the NVSwitch into every rank's copy, instead of `world` P2P stores.
 * The caller brackets the call with a cross-rank barrier (peers done reading the previous contents / all stores landed).
 */
int osr_pln_encode_gather_fwd(const float* x, const float* W, const float* bias, int R, int F, int E,
                              const uint64_t* h_peer_buffers, int world, int rank, uint64_t multicast_buffer,
                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * PLN.inference nearest-prototype classification (prototype_learning_network.py:203-226), all images at once:
 *   pred[i] = class of the nearest prototype of normalize(emb[i]) (min over the reps of a class first),
 *   mapped through class_id_map (K int64, may be NULL = identity; the reference's self.class_id for GraspNet),
 *   or unknown_id when the minimum distance (distance_type as above) is > unk_thr.  min_dist (R) fp32 is also returned.
 */
int osr_pln_nearest(const float* emb, const float* reps, int R, int D, int K, int reps_per_class, int distance_type,
                    float unk_thr, int64_t unknown_id, const int64_t* class_id_map, int64_t* pred, float* min_dist,
                    void* stream);

/*
 * Proposal <-> ground-truth matching of the ROI-head sampling glue, all images at once
 * (label_and_sample_proposals, osrcnn_roi_heads.py:136-230: detectron2 pairwise_iou + Matcher([0.5],[0,1]) +
 * the matched-IoU gather at :193 + the class assignment of ROIHeads._sample_proposals).
 *   boxes        (P,4) fp32 xyxy, the per-image proposal lists concatenated (ground-truth boxes already appended
 *                when PROPOSAL_APPEND_GT is on); box_offsets (N+1) int32 DEVICE: image n owns [off[n], off[n+1]),
 *                or, when box_counts != NULL, [off[n], off[n] + box_counts[n * box_counts_stride]) - the padded
 *                (N, Kmax, 4) output of osr_rpn_select_decode with its counts column, no repacking
 *   gt_boxes     (G,4) fp32, gt_classes (G) int64, gt_offsets (N+1) int32 DEVICE
 *   max_boxes_per_image  host upper bound of off[n+1]-off[n] (grid sizing only)
 * Outputs, (P) each: matched_idx int32 = FIRST arg-max GT of the image (0 when the image has no GT),
 *   matched_iou fp32 = that IoU (bit-exact with torch's op-by-op fp32 evaluation), matched_label int32 = iou >= thr,
 *   matched_class int64 = gt_classes[matched] or background_label.  No workspace, one launch.
 */
int osr_match_label(const float* boxes, const int32_t* box_offsets, const int32_t* box_counts,
                    int box_counts_stride, const float* gt_boxes, const int64_t* gt_classes,
                    const int32_t* gt_offsets, int num_images, int max_boxes_per_image, float iou_threshold,
                    int64_t background_label, int32_t* matched_idx, float* matched_iou, int32_t* matched_label,
                    int64_t* matched_class, void* stream);

/*
 * Labelled RoI sampling for ALL images in one launch, fused with the gather of the sampled fields.
 * Replaces the per-image detectron2 subsample_labels (two torch.randperm draws + nonzero, host syncs) inside
 * ROIHeads._sample_proposals and the per-image field gathers of label_and_sample_proposals
 *   openset_rcnn/modeling/roi_heads/osrcnn_roi_heads.py:136-175 (_sample_proposals), :203-226 (sampled fields)
 *   labels   (P) int64 = matched_class of osr_match_label: -1 ignored, background_label = negative, else positive
 *   keys     (P) fp32 random keys, one per row (e.g. one torch.rand draw): per image the kept positives are the
 *            min(#pos, num_pos_max) rows with the SMALLEST keys, the kept negatives the min(#neg, num_samples - kept
 *            positives) smallest - ties by lower row index - i.e. positive[perm[:num_pos]] of the reference with
 *            perm = stable argsort of the kind's keys: a uniformly random subset in uniformly random order
 *   box_offsets / box_counts / box_counts_stride: row layout as in osr_match_label; max_boxes_per_image: host upper
 *            bound of an image's row count (shared-memory sizing: up to ~40 000 rows are cached on chip; 0 or more = the
 *            rows are re-read from global memory in every pass - same result)
 *   boxes (P,4), logits (P), ious (P), matched_idx (P) int32, gt_offsets (N+1) int32: sources of the optional outputs
 * Outputs: out_index (N, num_samples) int32 = sampled row inside its image, positives first, each kind in ascending
 *   (key, row) order, -1 beyond the image's count; out_count (N, 2) int32 = (kept positives, kept rows);
 *   optional (NULL to skip), each (N, num_samples[, 4]) and undefined beyond the count: out_boxes, out_logits,
 *   out_classes int64 (= labels), out_ious, out_gt int64 (= gt_offsets[n] + matched_idx: row in the concatenated targets),
 *   out_rois (N, num_samples, 5) fp32 = (image n, x1, y1, x2, y2): the `rois` rows of osr_roi_align_fwd.
 * num_samples <= 4096.  No workspace, one launch, no host sync.
 */
int osr_sample_rois(const int64_t* labels, const float* keys, const int32_t* box_offsets, const int32_t* box_counts,
                    int box_counts_stride, int num_images, int max_boxes_per_image, int num_samples, int num_pos_max,
                    int64_t background_label,
                    const float* boxes, const float* logits, const float* ious, const int32_t* matched_idx,
                    const int32_t* gt_offsets, int32_t* out_index, int32_t* out_count, float* out_boxes,
                    float* out_logits, int64_t* out_classes, float* out_ious, int64_t* out_gt, float* out_rois,
                    void* stream);

/*
 * ROI-head inference post-processing, stage 1 (osrcnn_fast_rcnn.py:380-450 and :89-126), all images in one launch:
 * detectron2 Box2BoxTransform(weights = wx,wy,ww,wh; scale_clamp = log(1000/16)).apply_deltas on the class-agnostic
 * (R,4) deltas, objectness = sqrt(iou * centerness) (geometric_mean != 0) or (iou + centerness) / 2, isfinite filter,
 * Boxes.clip to the image, score > score_thresh.  box_offsets (N+1) int32 DEVICE, image_hw (N,2) int32 DEVICE (h, w).
 * out_boxes (R,4): clipped predictions (zeros for dropped rows); out_scores (R): the objectness (NaN for rows with a
 * non-finite box or score, which the reference's isfinite filter removes);
 * out_effective_scores (R): objectness, or -inf for rows the reference drops - feed them to osr_nms_segmented with
 * segments = images: survivors come out first, in the reference's order.  Arithmetic is torch's op-by-op fp32 chain.
 */
int osr_rcnn_decode_score(const float* proposal_boxes, const float* deltas, const float* ious, const float* centerness,
                          const int32_t* box_offsets, const int32_t* image_hw, int num_images,
                          int max_boxes_per_image, float wx, float wy, float ww, float wh, float scale_clamp,
                          int geometric_mean, float score_thresh, float* out_boxes, float* out_scores,
                          float* out_effective_scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSR_H_ */
