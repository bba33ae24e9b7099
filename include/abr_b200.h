/*
 * abr_b200.h -- C ABI of libabr_b200.so: the RoI hot path of ABR_IOD on B200 (sm_100a).
 *
 * This is the drop-in boundary.  It replaces the pybind11 module `maskrcnn_benchmark._C`
 * of the reference (maskrcnn_benchmark/csrc/vision.cpp:9-24) for the ops on the hot path and
 * adds native entry points for the two ops the reference runs as Python (ARD loss, ABR paste).
 * INTEGRATION.md shows the reference-side binding for every function.
 *
 * Conventions (all functions):
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns every buffer, including workspaces; nothing is allocated, freed or
 *     synchronised here; work is enqueued on `stream` (a cudaStream_t) and the call returns;
 *   - return value: ABR_OK or an ABR_ERR_* code; abr_last_error() gives the thread-local text;
 *   - `dtype`  : element type of feature maps / pooled tensors / gradients (rois are always fp32);
 *   - `layout` : ABR_NCHW = the reference's contiguous [B,C,H,W] -> [R,C,PH,PW];
 *                ABR_NCHW_MAPS_NHWC_POOLED (ROIAlign only) = contiguous feature / gradient maps,
 *                channels-last pooled tensors: what a contiguous model gets when it accepts
 *                channels-last RoI features (saves both passes over the pooled tensor);
 *                ABR_NHWC = channels-last storage of the same logical tensors,
 *                [B,H,W,C] -> [R,PH,PW,C] (torch.channels_last), the vectorised fast path.
 *   - rois are [R,5] fp32 rows (batch_index, x1, y1, x2, y2) in image pixels, exactly the
 *     reference's format (modeling/poolers.py:73-78).
 */
#ifndef ABR_B200_H_
#define ABR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABR_B200_VERSION 100 /* major*100 + minor */

#if defined(__GNUC__)
#define ABR_API __attribute__((visibility("default")))
#else
#define ABR_API
#endif

typedef void* abr_stream_t; /* cudaStream_t */

enum abr_status {
  ABR_OK = 0,
  ABR_ERR_BAD_ARG = 1,     /* null pointer, negative size, rois not [R,5], ... */
  ABR_ERR_UNSUPPORTED = 2, /* dtype / layout / size combination not implemented */
  ABR_ERR_CUDA = 3,        /* launch or runtime error; text has cudaGetErrorString */
  ABR_ERR_WORKSPACE = 4    /* workspace pointer null or smaller than *_workspace_bytes() */
};

enum abr_dtype { ABR_F32 = 0, ABR_BF16 = 1 };
enum abr_layout { ABR_NCHW = 0, ABR_NHWC = 1, ABR_NCHW_MAPS_NHWC_POOLED = 2 };

ABR_API int abr_version(void);
ABR_API const char* abr_last_error(void);
/* Tuning switches for measurements and tests (process-wide; initial values from the environment variables ABR_ROI_V2,
 * ABR_FWD_TMA, ABR_BWD_TMA, ABR_ARD_CLUSTER).  Keys: "roi_v2" (-1 automatic, 0 never, 1 whenever the
 * output is at most 16x16: gather-form ROIAlign kernels instead of the TMA-staged ones), "fwd_tma",
 * "bwd_tma", "ard_cluster" (0 / 1).  They select between kernels that compute the same results. */
ABR_API int abr_set_option(const char* key, int value);
/* Stage timing of the multi-kernel entry points (measurement only; bench.py's per-kernel roofline).  Between
 * abr_stage_timing_begin(max_calls) and abr_stage_timing_end(), every abr_roi_ard_fused call records CUDA events on the
 * CALLER'S stream around its stages (0 plan, 1 teacher+student pooling, 2 ARD coefficients, 3 zero-fill + backward).
 * abr_stage_timing_end synchronises on the last event, writes the average milliseconds per call of each stage into
 * avg_ms[0..n_stages) and returns the number of calls recorded (< 0 on error).  Not re-entrant; one stream at a time.
 * abr_stage_timing_begin_every(max_calls, every) records only every `every`-th call (the five event records of a call
 * cost ~15 us of stream time at configs[0]; sampling keeps that out of most timed steps). */
#define ABR_FUSED_STAGES 4
ABR_API int abr_stage_timing_begin(int max_calls);
ABR_API int abr_stage_timing_begin_every(int max_calls, int every);
ABR_API int abr_stage_timing_end(float* avg_ms, int n_stages);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
ABR_API uint64_t abr_launch_count(void);

/* ---------------------------------------------------------------- ROIAlign
 * Replaces _C.roi_align_forward / _C.roi_align_backward
 *   (csrc/ROIAlign.h:11-46, csrc/cuda/ROIAlign_cuda.cu:257-346, csrc/cpu/ROIAlign_cpu.cpp:221-257).
 * sampling_ratio <= 0 selects the adaptive grid ceil(roi_size / pooled_size) (ROIAlign_cuda.cu:100-101).
 * R == 0 is a no-op (ROIAlign_cuda.cu:278-281).
 * `workspace` (optional, 16-byte aligned, abr_roi_align_workspace_bytes(R, PH, PW, max H over levels) bytes) holds the
 * per-RoI interpolation plans of the fast NHWC kernels; with NULL (or the NCHW layout) the self-contained kernels run.
 * A call leaves the plans of its RoIs in the workspace; a later call -- the backward, or another forward such as the
 * student's after the teacher's -- for the SAME rois / levels / output size / sampling ratio / scales / map shapes may be
 * handed that workspace with workspace_has_plan != 0 and then skips planning (0: contents are treated as scratch). */
ABR_API size_t abr_roi_align_workspace_bytes(int R, int PH, int PW, int max_h);
/* Workspace that additionally lets an ABR_NCHW call run through the channels-last kernels (256-byte aligned): plans + a
 * channels-last copy of every level's map (sum_hw = sum over levels of H*W) + one of the pooled tensor.  With less, an
 * NCHW call uses the direct NCHW kernels. */
ABR_API size_t abr_roi_align_workspace_bytes_nchw(int R, int PH, int PW, int max_h, int B, int C, long long sum_hw, int dtype);
/* The same for any layout: ABR_NHWC -> the plans only, ABR_NCHW -> as above, ABR_NCHW_MAPS_NHWC_POOLED -> plans + the
 * channels-last copies of the maps (required for that layout: there is no direct kernel for it). */
ABR_API size_t abr_roi_align_workspace_bytes_layout(int R, int PH, int PW, int max_h, int B, int C, long long sum_hw, int dtype,
                                                    int layout);
ABR_API int abr_roi_align_forward(const void* input, const float* rois, void* output,
                          int B, int C, int H, int W, int R, int PH, int PW,
                          float spatial_scale, int sampling_ratio,
                          int dtype, int layout, void* workspace, size_t workspace_bytes, int workspace_has_plan,
        abr_stream_t stream);

/* grad_input [B,C,H,W] is zero-filled first when zero_init != 0 (the reference always does,
 * ROIAlign_cuda.cu:316); pass 0 to accumulate into an existing gradient. */
ABR_API int abr_roi_align_backward(const void* grad_output, const float* rois, void* grad_input,
                           int B, int C, int H, int W, int R, int PH, int PW,
                           float spatial_scale, int sampling_ratio,
                           int dtype, int layout, int zero_init, void* workspace, size_t workspace_bytes,
                           int workspace_has_plan, abr_stream_t stream);

/* Multi-level (FPN) pooling in ONE launch: replaces the per-level nonzero / gather / launch / scatter
 * loop of Pooler.forward (modeling/poolers.py:93-105).  `levels[r]` in [0,L) selects the feature map
 * of RoI r (LevelMapper, modeling/poolers.py:31-42).  inputs_host / hs_host / ws_host / scales_host
 * are HOST arrays of length L (L <= ABR_MAX_LEVELS); every level has the same B and C.
 * Output rows stay in RoI order. */
#define ABR_MAX_LEVELS 8
ABR_API int abr_roi_align_multilevel_forward(const void* const* inputs_host, const int* hs_host, const int* ws_host,
                                     const float* scales_host, int L,
                                     const float* rois, const int32_t* levels, void* output,
                                     int B, int C, int R, int PH, int PW, int sampling_ratio,
                                     int dtype, int layout, void* workspace, size_t workspace_bytes, int workspace_has_plan,
        abr_stream_t stream);
ABR_API int abr_roi_align_multilevel_backward(const void* grad_output, const float* rois, const int32_t* levels,
                                      void* const* grad_inputs_host, const int* hs_host, const int* ws_host,
                                      const float* scales_host, int L,
                                      int B, int C, int R, int PH, int PW, int sampling_ratio,
                                      int dtype, int layout, int zero_init, void* workspace, size_t workspace_bytes,
                                      int workspace_has_plan, abr_stream_t stream);

/* FPN level of each RoI: floor(k0 + log2(sqrt(area)/s0 + eps)) clamped to [k_min,k_max], minus k_min,
 * area with the +1 convention (modeling/poolers.py:31-42, structures/bounding_box.py:227-231). */
ABR_API int abr_fpn_map_levels(const float* rois, int32_t* levels, int R, float k_min, float k_max,
                       float canonical_scale, float canonical_level, float eps, abr_stream_t stream);

/* ---------------------------------------------------------------- ROIPool
 * Replaces _C.roi_pool_forward / _C.roi_pool_backward
 *   (csrc/ROIPool.h:9-48, csrc/cuda/ROIPool_cuda.cu:16-202).
 * argmax is int32 with the flat index h*W+w inside the (image, channel) plane or -1 for an empty
 * bin (ROIPool_cuda.cu:57-75,125); it has the layout of `output`. */
ABR_API int abr_roi_pool_forward(const void* input, const float* rois, void* output, int32_t* argmax,
                         int B, int C, int H, int W, int R, int PH, int PW, float spatial_scale,
                         int dtype, int layout, abr_stream_t stream);
ABR_API int abr_roi_pool_backward(const void* grad_output, const int32_t* argmax, const float* rois, void* grad_input,
                          int B, int C, int H, int W, int R, int PH, int PW,
                          int dtype, int layout, int zero_init, abr_stream_t stream);

/* ---------------------------------------------------------------- NMS (batched, no host sync)
 * Replaces _C.nms (csrc/nms.h:10-28, csrc/cuda/nms.cu:70-131) and the per-image Python loop that
 * calls it (modeling/rpn/inference.py:111-117, structures/boxlist_ops.py:9-31).
 *   boxes  [total,4] fp32 xyxy, scores [total] fp32: the images' boxes back to back;
 *   offsets_host [n_images+1]: HOST prefix offsets, image i owns [offsets[i], offsets[i+1]);
 *   keep   [n_images, keep_stride] int64: per image the surviving indices RELATIVE TO THE IMAGE,
 *          ascending (nms.cu:127-130), truncated to max_keep when max_keep > 0
 *          (boxlist_ops.py:28-29); unused tail is filled with -1;  keep_stride >= min(n_i, max_keep);
 *   n_keep [n_images] int32: number of valid entries per image.
 * Greedy order is by score descending, ties by ascending index (torch.sort(stable=True)); the
 * reference's tie order is whatever its torch.sort(stable=False) yields.  Suppression test is
 * IoU > thresh with the +1 pixel convention, IEEE fp32 without contraction (nms.cu:13-21,60);
 * `ge` != 0 selects the CPU flavour IoU >= thresh (csrc/cpu/nms_cpu.cpp:60). */
ABR_API size_t abr_nms_workspace_bytes(const int* offsets_host, int n_images);
ABR_API int abr_nms_batched(const float* boxes, const float* scores, const int* offsets_host, int n_images,
                    float thresh, int ge, int max_keep, int64_t* keep, int keep_stride, int32_t* n_keep,
                    void* workspace, size_t workspace_bytes, abr_stream_t stream);

/* ---------------------------------------------------------------- RPN proposal selection (around NMS)
 * Replaces RPNPostProcessor.forward_for_single_feature_map (modeling/rpn/inference.py:76-118) for a
 * whole batch with no host synchronisation: sigmoid, top pre_nms_top_n per image (inference.py:89-95),
 * BoxCoder.decode (modeling/box_coder.py:52-95), clip_to_image (structures/bounding_box.py:214-219),
 * remove_small_boxes (structures/boxlist_ops.py:34-48) and boxlist_nms with max_proposals =
 * post_nms_top_n (structures/boxlist_ops.py:9-31).
 *   objectness     [N,A,H,W] fp32 logits and box_regression [N,4A,H,W] fp32, the RPN head outputs in
 *                  place: ABR_NCHW contiguous, or ABR_NHWC = channels-last storage of the same tensors;
 *   anchors        [.,A*H*W,4] fp32 xyxy in the reference's (h, w, a) order; image i uses
 *                  anchors + i*anchor_image_stride floats (0 = one set shared by all images);
 *   image_sizes_host [N][2] HOST ints (width, height) = BoxList.size of each image's anchors;
 *   weights4_host  HOST (wx, wy, ww, wh) and bbox_xform_clip of the BoxCoder;
 *   proposals      [N,out_stride,4] fp32, scores [N,out_stride] fp32 (= sigmoid(logit), the
 *                  "objectness" field), anchor_index [N,out_stride] int32 (optional, may be NULL):
 *                  per image the surviving proposals in the reference's order, zero / -1 padded;
 *   n_out          [N] int32 valid entries per image;  out_stride >= min(pre_nms_top_n, A*H*W,
 *                  post_nms_top_n if > 0)  (nms_thresh <= 0 skips NMS and the cut, like boxlist_nms).
 * Candidates are ranked by (logit descending, anchor index ascending); the reference's torch.topk
 * leaves the order of equal scores unspecified.  pre_nms_top_n <= 16384.  `ge` as in abr_nms_batched. */
ABR_API size_t abr_rpn_proposals_workspace_bytes(int N, int A, int H, int W, int pre_nms_top_n, int post_nms_top_n);
ABR_API int abr_rpn_proposals(const float* objectness, const float* box_regression, const float* anchors,
                              long long anchor_image_stride, const int* image_sizes_host, int N, int A, int H, int W,
                              int layout, int pre_nms_top_n, int post_nms_top_n, float nms_thresh, int ge, float min_size,
                              const float* weights4_host, float bbox_xform_clip, float* proposals, float* scores,
                              int32_t* anchor_index, int32_t* n_out, int out_stride, void* workspace,
                              size_t workspace_bytes, abr_stream_t stream);

/* ---------------------------------------------------------------- box-head post-processing (class-batched NMS)
 * Replaces PostProcessor.forward / filter_results (modeling/roi_heads/box_head/inference.py:42-151) for a
 * whole batch with no host synchronisation: softmax (:56), BoxCoder.decode of the class deltas (:65-67,
 * modeling/box_coder.py:52-95), clip_to_image (:80), score threshold (:117), per-class boxlist_nms
 * (:119-126) as ONE batched NMS over all (image, class) pairs, concatenation of the foreground classes
 * in class order (:139) and the detections_per_img cut (:142-149: scores >= the k-th largest, ties survive).
 *   class_logits   [R,C] fp32; box_regression rows of reg_row_stride floats: class j's deltas at
 *                  columns 4j..4j+3, or (cls_agnostic != 0, :63-64,68-69) the row's LAST four columns for
 *                  every class; proposals [R,4] fp32 xyxy: the images' proposals back to back;
 *   boxes_per_image_host [n_images], image_sizes_host [n_images][2] (width, height): HOST arrays;
 *   det_boxes [n_images,det_stride,4] fp32, det_scores [n_images,det_stride] fp32, det_labels
 *   [n_images,det_stride] int64, det_rows [n_images,det_stride] int32 (optional: the proposal each
 *   detection came from, image-relative), n_det [n_images] int32: the results in the reference's order,
 *   zero / -1 padded.  n_det[i] is the TRUE count: if it exceeds det_stride (equal scores tying at the
 *   cut, or no cut), only det_stride entries were written and the caller re-runs with a wider stride;
 *   bg_boxes [n_images,bg_stride,4], bg_scores [n_images,bg_stride], n_bg [n_images]: class 0 after its
 *   own NMS (the reference returns the last image's: :82,137-138); bg_stride >= max boxes per image.
 * nms_thresh <= 0 skips the NMS like boxlist_nms does.  `ge` as in abr_nms_batched.  Score ties inside
 * the NMS break by ascending proposal index (see abr_nms_batched). */
ABR_API size_t abr_box_postprocess_workspace_bytes(const int* boxes_per_image_host, int n_images, int num_classes);
ABR_API int abr_box_postprocess(const float* class_logits, const float* box_regression, int reg_row_stride, int cls_agnostic,
                                const float* proposals, const int* boxes_per_image_host, const int* image_sizes_host,
                                int n_images, int num_classes, float score_thresh, float nms_thresh, int ge,
                                int detections_per_img, const float* weights4_host, float bbox_xform_clip,
                                float* det_boxes, float* det_scores, int64_t* det_labels, int32_t* det_rows,
                                int32_t* n_det, int det_stride, float* bg_boxes, float* bg_scores, int32_t* n_bg,
                                int bg_stride, void* workspace, size_t workspace_bytes, abr_stream_t stream);

/* ---------------------------------------------------------------- Attentive RoI Distillation
 * Native form of calculate_attentive_roi_feature_distillation (distillation/distillation.py:86-130)
 * and of its autograd backward, in one kernel.  f_old = old model / teacher pooled features (argument 0
 * at the call site, tools/train_incremental.py:115; no gradient), f_new = student features.
 *   loss3 [3] fp32: { L_afd + gamma*L_pad, L_afd, L_pad }  (written, not accumulated)
 *   grad_new: dL/df_new * grad_scale, same dtype/layout as f_new; may be NULL (loss only).
 * Logical shape [N,C,H,W] with HW = H*W positions; layout ABR_NCHW ([N,C,HW]) or ABR_NHWC ([N,HW,C]). */
ABR_API size_t abr_ard_workspace_bytes(int N, int C, int HW);
ABR_API int abr_ard_forward_backward(const void* f_old, const void* f_new, void* grad_new, float* loss3,
                             int N, int C, int HW, float gamma, float grad_scale,
                             int dtype, int layout, void* workspace, size_t workspace_bytes,
                             abr_stream_t stream);
/* ---------------------------------------------------------------- fused ARD step (pool teacher + student, loss, backward)
 * One call for the RoI part of the distillation step of tools/train_incremental.py:84-115: the teacher's and the
 * student's ROIAlign over the SAME RoIs (generalized_rcnn.py:121-167 -> roi_box_feature_extractors.py:44-48 for the
 * teacher, train_incremental.py:93-95 for the student), the ARD loss (distillation/distillation.py:86-130) and the
 * backward of the ARD loss through the student's ROIAlign (csrc/cuda/ROIAlign_cuda.cu:177-254).  Three kernels:
 * both maps are pooled in one pass that also emits the per-position channel sums, a per-RoI kernel
 * turns them into the loss and two coefficients per position, and the backward forms
 * dL/df_new = ka*(f_new - f_old) + kb*f_new on the fly while scattering -- the ARD gradient tensor is never written
 * or re-read and the pooled tensors are read once instead of three times.
 *   teacher_map / student_map [B,H,W,C] and pooled_old / pooled_new [R,PH,PW,C]: ABR_NHWC, ABR_F32 only (other
 *   combinations: ABR_ERR_UNSUPPORTED -- use the separate ops); the pooled tensors are written (the box head consumes them);
 *   grad_student_map [B,H,W,C] (may be NULL: loss only) receives grad_scale * dL/d(student_map), zero-filled first
 *   when zero_init != 0;  loss3 as in abr_ard_forward_backward;  PH, PW <= 16;  R > 0.
 * workspace: 256-byte aligned, abr_roi_ard_fused_workspace_bytes(R, C, PH, PW) bytes (plans, channel sums, coefficients);
 * workspace_has_plan != 0: the workspace was last used by this function for the SAME rois / sizes / scale (skips planning). */
ABR_API size_t abr_roi_ard_fused_workspace_bytes(int R, int C, int PH, int PW);
ABR_API int abr_roi_ard_fused(const void* teacher_map, const void* student_map, const float* rois, void* pooled_old,
                              void* pooled_new, void* grad_student_map, float* loss3, int B, int C, int H, int W, int R,
                              int PH, int PW, float spatial_scale, int sampling_ratio, float gamma, float grad_scale,
                              int dtype, int layout, int zero_init, void* workspace, size_t workspace_bytes,
                              int workspace_has_plan, abr_stream_t stream);

/* data[i] *= (*scale_dev / expected) unless *scale_dev == expected (then no memory is touched):
 * lets the autograd backward apply an upstream gradient that differs from the grad_scale baked in. */
ABR_API int abr_scale_if_needed(void* data, size_t n, const float* scale_dev, float expected, int dtype,
                        abr_stream_t stream);

/* ---------------------------------------------------------------- proposal <-> ground-truth matching
 * abr_match_proposals replaces FastRCNNLossComputation.match_targets_to_proposals + prepare_targets
 * (modeling/roi_heads/box_head/loss.py:43-84) for a whole batch: boxlist_iou (structures/boxlist_ops.py:53-88),
 * Matcher without low-quality matches (modeling/matcher.py:52-81), labels and BoxCoder.encode
 * (modeling/box_coder.py:22-50).
 *   proposals [R,4] fp32 xyxy, the images' proposals back to back (boxes_per_image_host [n_images]);
 *   gt_boxes [G,4] fp32 xyxy + gt_labels [G] int64, likewise (gt_per_image_host [n_images], each > 0);
 *   matched_idxs [R] int64: index of the best ground-truth box inside its image (first maximum wins), -1 below
 *   low_threshold, -2 in [low_threshold, high_threshold);  labels [R] int64: the matched box's label, 0 for -1,
 *   -1 (ignored by the sampler) for -2;  regression_targets [R,4] fp32: encode(matched box, proposal), for
 *   unmatched proposals against ground-truth box 0 like the reference's clamp(min=0).
 * abr_box_iou is boxlist_iou itself: iou [N,M] of boxes1 [N,4] x boxes2 [M,4], +1 pixel convention. */
ABR_API int abr_match_proposals(const float* proposals, const int* boxes_per_image_host, const float* gt_boxes,
                                const int64_t* gt_labels, const int* gt_per_image_host, int n_images,
                                float high_threshold, float low_threshold, const float* weights4_host,
                                int64_t* matched_idxs, int64_t* labels, float* regression_targets, abr_stream_t stream);
ABR_API int abr_box_iou(const float* boxes1, int N, const float* boxes2, int M, float* iou, abr_stream_t stream);

/* ---------------------------------------------------------------- logit-level losses (forward + backward)
 * abr_roi_distillation_id replaces calculate_roi_distillation_losses(dist='id')
 * (distillation/distillation.py:164-241) and its autograd backward:
 *   soften_scores [R,C_old] / soften_bboxes [R,C_old,4]: the TEACHER's class logits and box deltas (no grad);
 *   target_scores [R,C_total] / target_bboxes [R,C_total,4]: the STUDENT's, C_total > C_old;
 *   loss3 = {class term + box term, class term, box term}: unbiased cross-entropy (:191-201) and the L2 box term
 *   over the old foreground classes 1..C_old-1 (:206-211);
 *   grad_scores [R,C_total] / grad_bboxes [R,C_total,4] (either may be NULL): grad_scale * d(total)/d(student).
 * abr_fastrcnn_loss replaces FastRCNNLossComputation.__call__ (modeling/roi_heads/box_head/loss.py:122-184):
 *   class_logits [R,C], box_regression rows of reg_row_stride floats, labels [R] int64 (-100 = ignored row, as
 *   F.nll_loss), regression_targets [R,4];  n_old >= 0: inclusive classification loss (:151-159), n_old < 0:
 *   F.cross_entropy (:162);  box term: smooth-L1 (layers/smooth_l1_loss.py:6-18, `beta`) on columns 4*label..+3 of
 *   the rows with label > 0 (columns 4..7 when cls_agnostic), summed and divided by R (:166-180);
 *   loss2 = {classification_loss, box_loss};  grad_logits [R,C] = grad_scale_cls * d(cls)/d(logits), grad_regression
 *   [R,reg_row_stride] = grad_scale_box * d(box)/d(regression) (either may be NULL).
 * workspace: abr_logit_loss_workspace_bytes(R).  Row sums are reduced in fixed order (deterministic). */
ABR_API size_t abr_logit_loss_workspace_bytes(int R);
ABR_API int abr_roi_distillation_id(const float* soften_scores, const float* soften_bboxes, const float* target_scores,
                                    const float* target_bboxes, int R, int C_old, int C_total, float grad_scale,
                                    float* grad_scores, float* grad_bboxes, float* loss3, void* workspace,
                                    size_t workspace_bytes, abr_stream_t stream);
ABR_API int abr_fastrcnn_loss(const float* class_logits, const float* box_regression, int reg_row_stride,
                              const int64_t* labels, const float* regression_targets, int R, int num_classes, int n_old,
                              int cls_agnostic, float beta, float grad_scale_cls, float grad_scale_box,
                              float* grad_logits, float* grad_regression, float* loss2, void* workspace,
                              size_t workspace_bytes, abr_stream_t stream);

/* ---------------------------------------------------------------- Prototype Box Selection (scoring)
 * abr_channel_mean: out[r][p] = mean over channels of pooled[r][:, p] -- the per-RoI descriptor of
 * tools/prototype_box_selection.py:96-101 (torch.mean(roi_align_features.cpu(), dim=1)); pooled is
 * [R,C,HW] (ABR_NCHW) or [R,HW,C] (ABR_NHWC), fp32 or bf16; out [R,HW] fp32.
 * abr_prototype_distances: Mem.mean_feature_sampling's scoring (tools/extract_memory.py:125-141) for one
 * class in float64: features [n,F] fp32 (F = 49); mean_out [F] = normalised class mean; dist [n] = distance
 * of each descriptor / |all descriptors|_F to it.  The n smallest-distance boxes (ascending) are the prototypes. */
ABR_API int abr_channel_mean(const void* pooled, int R, int C, int HW, int dtype, int layout, float* out, abr_stream_t stream);
ABR_API int abr_prototype_distances(const float* features, int n, int F, double* mean_out, double* dist, abr_stream_t stream);
/* abr_prototype_herding: the greedy loop of Mem.herding_feature_sampling (tools/extract_memory.py:163-211) for one class:
 * k times, among the boxes not yet chosen, the one whose inclusion brings the running centre
 * (centre*f/(f+1) + feature/(f+1)) closest to the normalised class mean `mean` [F] (as abr_prototype_distances returns it);
 * float64 in numpy's operation order, first index on ties.  selected [k] int64 (device) receives the order.
 * workspace: 8-byte aligned, F*8 + n bytes. */
ABR_API int abr_prototype_herding(const float* features, int n, int F, int k, const double* mean, int64_t* selected,
                                  void* workspace, size_t workspace_bytes, abr_stream_t stream);

/* ---------------------------------------------------------------- RoI sampling (fg / bg)
 * BalancedPositiveNegativeSampler.__call__ (modeling/balanced_positive_negative_sampler.py:19-68) for a whole batch in one
 * launch and without the per-image nonzero() synchronisations: per image (offsets_dev [n_images+1], device),
 * num_pos = min(#positives, max_positives) positives (matched_idxs >= 1; max_positives = int(batch_size_per_image *
 * positive_fraction), evaluated by the caller in double like the reference's Python) and
 * num_neg = min(#negatives, batch_size_per_image - num_pos) negatives (== 0) are drawn -- the ones with the smallest
 * caller-supplied random `keys` (one uniform float per RoI; ties by index), i.e. a uniformly random subset like the
 * reference's randperm prefix, from the caller's own random stream.  pos_mask / neg_mask [N] uint8; counts [n_images][2]. */
ABR_API int abr_sample_fg_bg(const int64_t* matched_idxs, const float* keys, const int* offsets_dev, int n_images,
                             int batch_size_per_image, int max_positives, uint8_t* pos_mask, uint8_t* neg_mask,
                             int* counts, abr_stream_t stream);

/* ---------------------------------------------------------------- ABR paste (mixup / mosaic)
 * Pixel part of PascalVOCDataset_ABR._start_mixup / _start_boxes_mosaic
 * (data/datasets/voc_abr.py:659-678 and :744-763) for a whole batch in one launch.  All images are
 * HWC uint8 (3 channels) inside one `canvas` buffer; prototypes are HWC uint8 inside one `pool`.
 * Ops of one image are applied in order per pixel (a later mixup box sees the earlier blend). */
enum abr_paste_kind {
  ABR_PASTE_FILL = 0,  /* dst rect <- value (mosaic canvas, 114; voc_abr.py:744) */
  ABR_PASTE_COPY = 1,  /* dst rect <- src rect (mosaic quadrant; voc_abr.py:763) */
  ABR_PASTE_BLEND = 2  /* dst rect <- trunc(lambda*dst + (1-lambda)*src) in fp64 (voc_abr.py:659-678) */
};
typedef struct abr_paste_image {
  int64_t offset; /* byte offset of pixel (0,0) inside canvas */
  int32_t height, width;
  int32_t first_op, n_ops; /* this image's slice of the ops array */
} abr_paste_image_t;
typedef struct abr_paste_op {
  int32_t kind;
  int32_t y0, x0, y1, x1; /* destination rectangle [y0,y1) x [x0,x1) */
  int32_t src_width;      /* row pitch of the prototype in pixels */
  int32_t sy0, sx0;       /* top-left of the source window inside the prototype */
  int64_t src_offset;     /* byte offset of the prototype inside pool */
  double lambda;          /* BLEND weight of the destination */
  int32_t fill;           /* FILL value 0..255 */
  int32_t pad_;
} abr_paste_op_t;
ABR_API int abr_paste_batch(uint8_t* canvas, const abr_paste_image_t* images, int n_images,
                    const abr_paste_op_t* ops, int n_ops, const uint8_t* pool, int max_pixels_per_image,
                    abr_stream_t stream);

/* ---------------------------------------------------------------- prototype resize (before the paste)
 * The rescaling of a Box-Rehearsal prototype in _sample_per_bbox_from_boxrehearsal (data/datasets/voc_abr.py:538-548:
 * PIL Image.resize((int(s*w), int(s*h))), default filter BICUBIC) for all crops of a batch, bit-exact with Pillow's 8-bit
 * resampling (horizontal pass into a uint8 intermediate, then the vertical pass; 22-bit fixed-point taps).
 *   pool: the prototypes (HWC uint8 RGB) at src_offset;  out: receives the crops at dst_offset (tmp_offset: src_h*dst_w*3
 *   bytes of scratch inside `out` for the intermediate);  taps (device int32): per axis and output coordinate
 *   { first input index, tap count, ksize taps } starting at x_taps / y_taps -- the tables of Pillow's precompute_coeffs +
 *   normalize_coeffs_8bpc, built by the host (abr_iod_b200/data/resample.py);  max_pixels: the largest pass output. */
typedef struct abr_resize_job {
  int64_t src_offset, dst_offset, tmp_offset;
  int32_t src_h, src_w, dst_h, dst_w;
  int32_t x_taps, x_ksize, y_taps, y_ksize; /* offsets in int32 units into `taps`; taps per output coordinate */
} abr_resize_job_t;
ABR_API int abr_resize_bicubic_batch(const uint8_t* pool, uint8_t* out, const abr_resize_job_t* jobs, int n_jobs,
                                     const int32_t* taps, int max_pixels, abr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ABR_B200_H_ */
