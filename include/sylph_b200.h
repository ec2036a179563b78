/*
 * sylph_b200 -- C ABI of the B200-native Meta-FCOS few-shot inference path.
 *
 * The reference (facebookresearch/sylph-few-shot-detection) is 100 % Python; its "FFI" for this path is the set of
 * nn.Module entry points listed below.  Each function names the reference interface it replaces (file:line under
 * /root/reference).  Plain pointers and sizes only; no torch types.  Conventions:
 *   - every function returns 0 on success, non-zero on failure; sylph_last_error(ctx) gives the message;
 *   - one context per device, thread-compatible (not thread-safe per context);
 *   - "dev" pointers are CUDA device pointers owned by the caller, "host" pointers are host memory;
 *   - all work is enqueued on the stream passed in (a cudaStream_t cast to void*; NULL = legacy default stream)
 *     and is asynchronous unless stated otherwise;
 *   - images are fp32 CHW, 3 channels in the order of MODEL.PIXEL_MEAN (BGR), values 0..255, NOT normalised;
 *   - a class code row is 257 floats: 256 conv weights (cls_conv[c, :, 0, 0]) followed by the bias.
 */
#ifndef SYLPH_B200_H_
#define SYLPH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sylph_ctx sylph_ctx;

#define SYLPH_CODE_STRIDE 257
#define SYLPH_NUM_LEVELS 5
#define SYLPH_DET_STRIDE 9 /* x0, y0, x1, y1, score, class, loc_x, loc_y, level  (all stored as float) */

/* Model hyper-parameters the hot path reads from the reference config ("Config / flags"). */
typedef struct sylph_model_config {
    int resnet_depth;          /* MODEL.RESNETS.DEPTH: 50 | 101 | 152 */
    int num_cls_convs;         /* MODEL.FCOS.NUM_CLS_CONVS */
    int num_box_convs;         /* MODEL.FCOS.NUM_BOX_CONVS */
    int use_scale;             /* MODEL.FCOS.USE_SCALE */
    int thresh_with_ctr;       /* MODEL.FCOS.THRESH_WITH_CTR */
    int box_quality;           /* bit 0: "ctrness", bit 1: "iou"  (MODEL.FCOS.BOX_QUALITY) */
    int pre_nms_topk;          /* MODEL.FCOS.PRE_NMS_TOPK_TEST */
    int post_nms_topk;         /* MODEL.FCOS.POST_NMS_TOPK_TEST */
    float inference_thresh;    /* MODEL.FCOS.INFERENCE_TH_TEST */
    float nms_thresh;          /* MODEL.FCOS.NMS_TH */
    float prior_prob;          /* MODEL.FCOS.PRIOR_PROB */
    float pixel_mean[3];       /* MODEL.PIXEL_MEAN */
    float pixel_std[3];        /* MODEL.PIXEL_STD */
    int cg_tower_layers;       /* len(CODE_GENERATOR.TOWER_LAYERS); each layer must be ["GN","ReLU"] */
    int cg_post_norm;          /* CODE_GENERATOR.POST_NORM == "GN" */
    int cg_conv_l2_norm;       /* CODE_GENERATOR.CONV_L2_NORM */
    int cg_bias_layer;         /* len(CODE_GENERATOR.BIAS_LAYER) == 3 */
    int cg_bias_l2_norm;       /* CODE_GENERATOR.BIAS_L2_NORM */
    int cg_use_bias;           /* CODE_GENERATOR.USE_BIAS (CondConvBasic use_bias) */
    int cg_has_conv_scale;     /* USE_WEIGHT_SCALE and (CONV_L2_NORM or POST_NORM) */
    int generator;             /* CODE_GENERATOR.NAME: 0 = "CodeGenerator" / "CodeGeneratorHead", 1 = "ROIEncoder";
                                  2 = none: base detector (MODEL.META_LEARN.EPISODIC_LEARNING off, meta_one_stage_detector.py:298-323):
                                  the rows handed to sylph_detect are cls_logits.weight / .bias (fcos.py:544-576) */
    int re_tok_convs;          /* ROIEncoder: CODE_GENERATOR.TOKENIZER.NUM_CONV (CONV_DIM 256, NORM "GN") */
    int re_tok_fcs;            /* ROIEncoder: CODE_GENERATOR.TOKENIZER.NUM_FC (FC_DIM 256 = transformer d_model) */
    int re_layers;             /* ROIEncoder: CODE_GENERATOR.TRANSFORMER_ENCODER.LAYERS (dim_feedforward = 4 x 256) */
    int re_head_fcs;           /* ROIEncoder: CODE_GENERATOR.HEAD.NUM_FC (OUTPUT_DIM 256) */
    int re_head_dim;           /* ROIEncoder: CODE_GENERATOR.HEAD.FC_DIM (<= 1024) */
    int cg_weight_layer;       /* len(CODE_GENERATOR.WEIGHT_LAYER) == 3: per-shot weights = softmax over the shots of a 256 -> 1 3x3
                                  convolution + global average pool (code_generator.py:583-612, 766-776, 969-981) instead of 1 / K */
} sylph_model_config;

/* Library / build identification (no GPU needed). */
const char* sylph_version(void);

/* Create / destroy a context bound to CUDA device `device`.  Replaces MetaOneStageDetector construction,
 * sylph/modeling/meta_arch/meta_one_stage_detector.py:72-88 (from_config). */
int sylph_create(sylph_ctx** ctx, int device, const sylph_model_config* cfg);
void sylph_destroy(sylph_ctx* ctx);
const char* sylph_last_error(const sylph_ctx* ctx);

/* Operand precision of every tensor-core convolution; call before sylph_finalize_weights (the weights are prepared for
 * one mode).  The reference computes in fp32 everywhere (sylph/modeling/meta_fcos/fcos.py:582-667,
 * fcos_outputs.py:904-1008; torch CPU kernels).
 *   exact = 1 (default): split-fp16 operands -- every activation and weight is a (hi, lo) fp16 pair with hi = rn(x),
 *       lo = rn(x - hi), every product is a_hi*w_hi + a_lo*w_hi + a_hi*w_lo accumulated in fp32 in the tensor core
 *       (three tcgen05.mma per k-step); outputs agree with the fp32 reference to ~1e-5 (max-norm), i.e. well inside
 *       the 1e-3 bar of the path.
 *   exact = 0: single fp16 operands (10-bit mantissa), one tcgen05.mma per k-step, half the activation bytes;
 *       deep activations / logits carry 1-2.5e-3 (max-norm) of accumulated operand rounding.
 * The environment variable SYLPH_PRECISION=fast|exact sets the default of new contexts. */
int sylph_set_precision(sylph_ctx* ctx, int exact);
int sylph_get_precision(const sylph_ctx* ctx);

/* Stage one state-dict tensor (host fp32, contiguous) under its reference key, then
 * finalize: fold FrozenBN into the convolutions, reorder to tap-major K-major, round to fp16 (or split into
 * fp16 hi/lo pairs), upload.
 * Replaces DetectionCheckpointer.load into the module tree (tools/train_net.py:51-61). Synchronous. */
int sylph_load_tensor(sylph_ctx* ctx, const char* key, const float* host_data, const int64_t* shape, int ndim);
int sylph_finalize_weights(sylph_ctx* ctx);

/* Feature slots: the context keeps NUM_SLOTS pyramids (p3..p7 for a batch of images) alive at a time. */
#define SYLPH_SLOT_SUPPORT 0
#define SYLPH_SLOT_QUERY 1
#define SYLPH_NUM_SLOTS 2

/* Normalise + pad to /32 + backbone.  Replaces convert_batched_inputs_to_image_list + backbone call,
 * meta_one_stage_detector.py:174-182 (and :245-247, :272-273).  images_dev[i] is a (3, heights[i], widths[i]) fp32
 * device tensor. */
int sylph_extract_features(sylph_ctx* ctx, int slot, int n_images, const float* const* images_dev, const int* heights,
                           const int* widths, void* stream);

/* Same, for uint8 images (what detectron2's DatasetMapper hands to the model; 4x less host-to-device traffic). */
int sylph_extract_features_u8(sylph_ctx* ctx, int slot, int n_images, const uint8_t* const* images_dev,
                              const int* heights, const int* widths, void* stream);

/* The backbone as the reference calls it: `self.backbone(images.tensor)` on an ALREADY normalised, zero-padded
 * (N, 3, H, W) fp32 batch (sylph/modeling/meta_arch/meta_one_stage_detector.py:174-182; build_fcos_resnet_fpn_backbone,
 * configs/COCO-Detection/Meta-FCOS/Base-FCOS.yaml:3-11).  Same kernels as sylph_extract_features with mean 0 / std 1;
 * H and W are padded up to multiples of 32 with zeros if they are not.  sylph_export_features returns p3..p7. */
int sylph_extract_features_normalized(sylph_ctx* ctx, int slot, int n_images, const float* batch_dev, int height, int width,
                                      void* stream);

/* Several image batches through ONE bottom-up trunk pass (stem, res2..res5), then the FPN of each batch into its own
 * slot: group g = images [sum(counts[0..g-1]), +counts[g]) -> slots[g].  Equivalent to n_groups sylph_extract_features
 * calls; the groups share a trunk batch only when they pad to the same size (ImageList.from_tensors pads each reference
 * call to its own batch maximum, meta_one_stage_detector.py:174-178), otherwise they run back to back.  The support and
 * query batches of an episode (forward_class_code :245-247 + forward_instances :272-273) are the intended use. */
int sylph_extract_features_multi(sylph_ctx* ctx, int n_groups, const int* slots, const int* counts,
                                 const void* const* images_dev, int is_u8, const int* heights, const int* widths, void* stream);

/* Plugin-level entry for features produced elsewhere (NCHW fp32 device tensors, one per level, (n, 256, H_l, W_l)):
 * the `features` argument of CodeGenerator.forward, sylph/modeling/code_generator/code_generator.py:1037-1053. */
int sylph_import_features(sylph_ctx* ctx, int slot, int n_images, int padded_h, int padded_w,
                          const float* const* level_ptrs_dev, const int* level_h, const int* level_w, void* stream);

/* Original (un-padded) sizes of the images whose features were imported with sylph_import_features: the
 * `image_sizes` of the ImageList handed to MetaFCOS.forward (sylph/modeling/meta_fcos/fcos.py:184-268).  Detection
 * boxes are scaled by out_size / image_size (detector_postprocess, meta_one_stage_detector.py:288-295), so a caller
 * that asks for out_sizes == image_sizes gets un-scaled proposals in the frame of the padded batch, as the reference's
 * proposal generator returns them.  Without this call the imported images count as padded_h x padded_w. Host-only. */
int sylph_set_image_sizes(sylph_ctx* ctx, int slot, int n_images, const int* heights, const int* widths);

/* Geometry of a slot after extract/import: level_h/level_w receive SYLPH_NUM_LEVELS entries. */
int sylph_feature_shape(sylph_ctx* ctx, int slot, int* n_images, int* padded_h, int* padded_w, int* level_h, int* level_w);

/* Copy one level of a slot out as NCHW fp32 (n, 256, H_l, W_l) -- what backbone(...)[f] returns in the reference. */
int sylph_export_features(sylph_ctx* ctx, int slot, int level, float* out_dev, void* stream);

/* Class-code generation for `n_classes` classes from the support slot.  boxes_host: (n_rois, 4) XYXY absolute pixels,
 * roi_image[i] = index of the support image of ROI i inside the slot, class_offsets: n_classes + 1 prefix offsets
 * into the ROI list (ROIs of a class are contiguous).  Writes RAW (un-normalised) codes, (n_classes, 257), to
 * codes_out_dev and, if non-NULL, the FPN level index of every ROI (int64, like assign_boxes_to_levels) to
 * levels_out_dev.  Replaces CodeGeneratorHead.forward_roi_align, code_generator.py:924-1002, called once per class
 * by forward_class_code, meta_one_stage_detector.py:229-254.
 * With generator == 1 the same entry runs ROIEncoder.forward (sylph/modeling/code_generator/roi_encoder.py:146-204):
 * ROIAlign -> conv3x3+GN+ReLU -> MS_CAM context gate -> tokenizer -> transformer encoder -> class-token mean ->
 * weight / bias hyper-network heads; a row is then the FINAL code (bias includes the -log((1-p)/p) prior) and
 * sylph_normalize_codes is not applicable. */
int sylph_generate_codes(sylph_ctx* ctx, int slot, int n_rois, const float* boxes_host /* host OR device pointer */, const int* roi_image,
                         int n_classes, const int* class_offsets, float* codes_out_dev, int64_t* levels_out_dev,
                         void* stream);

/* Copy the pooled ROI features of the last sylph_generate_codes call out as (n_rois, 256, 7, 7) fp32. */
int sylph_export_roi_features(sylph_ctx* ctx, float* out_dev, void* stream);

/* Code normalisation: GN(32) -> L2 -> x conv_scale; bias x bias_scale + prior.  In-place safe (out == in).
 * Replaces forward_normalize_code / code_process_module, code_generator.py:864-897. */
int sylph_normalize_codes(sylph_ctx* ctx, const float* raw_codes_dev, float* out_codes_dev, int n_classes, void* stream);

/* ---- Class-code exchange over NVLink peer memory: normalisation fused with the all-gather. ----
 * In the sharded episode every rank generates the raw codes of a contiguous class shard; all ranks need the NORMALISED
 * codes of all classes before detection.  The reference moves pickled code dicts with all_gather_object
 * (MetaFCOSRunner._gather_class_code, sylph/runner/meta_fcos_runner.py:381-396) and then normalises every class on
 * every rank (inference_normalization, sylph/evaluation/meta_learn_evaluation.py:105-116 ->
 * forward_normalize_code, code_generator.py:877-897).  Here the normalisation kernel stores each finished row directly
 * into the exchange buffer of EVERY rank of the box (peer memory mapped through CUDA IPC, stores travel over
 * NVLink / NVSwitch) and publishes it with a system-scope atomic; a one-block kernel on each rank waits for the rows of
 * the episode and hands them to the caller.  Two launches, no NCCL call, no host synchronisation.
 *
 * Setup (once per process group, all ranks of ONE box): sylph_exchange_create on every rank -> exchange the 64-byte
 * handles by any means (the Python layer uses one all_gather) -> sylph_exchange_connect with the handles of all ranks in
 * rank order.  world == 1 needs no handle exchange (handles_all may be NULL).  All ranks must then make the same
 * sequence of sylph_normalize_codes_exchange calls (same n_total per call). */
#define SYLPH_IPC_HANDLE_BYTES 64
int sylph_exchange_create(sylph_ctx* ctx, int world, int rank, int max_classes, uint8_t* handle_out /* 64 bytes */);
int sylph_exchange_connect(sylph_ctx* ctx, const uint8_t* handles_all /* world x 64 bytes, rank order */);

/* Normalise this rank's n_local raw codes (rows class_offset .. class_offset + n_local - 1 of the episode's n_total
 * classes; n_local may be 0), deliver them to every rank, wait for the rows of all other ranks and write the
 * (n_total, 257) normalised codes to all_codes_out_dev.  Same arithmetic as sylph_normalize_codes (bit-identical rows);
 * with generator == 1 (ROIEncoder, final codes) the rows travel unchanged.  A rank that waits longer than
 * SYLPH_EXCHANGE_TIMEOUT_MS (default 5000) gives up and raises the flag sylph_exchange_status reports. */
int sylph_normalize_codes_exchange(sylph_ctx* ctx, const float* raw_codes_dev, int n_local, int class_offset, int n_total,
                                   float* all_codes_out_dev, void* stream);

/* Non-blocking: non-zero (with a message) once an exchange of this context has given up waiting.  The flag travels to
 * pinned host memory behind every exchange, so it is current for every episode whose results the caller has already
 * synchronised on; sylph_normalize_codes_exchange checks it on entry as well. */
int sylph_exchange_poll(sylph_ctx* ctx);

/* Synchronous: *timed_out = 1 if any exchange of this context gave up waiting; *rows_arrived = rows received so far. */
int sylph_exchange_status(sylph_ctx* ctx, int* timed_out, int64_t* rows_arrived);
void sylph_exchange_destroy(sylph_ctx* ctx);

/* Base-class "all ground truths" path (MODEL.META_LEARN.USE_ALL_GTS_IN_BASE_CLASSES): a class arrives as several
 * chunks of <= 10 support boxes.  acc[class_of(k)] += chunk_codes[k] * chunk_weight[k] for k = 0..n_chunks-1 in order,
 * with the reference's fp32 rounding sequence (multiply, then add; no FMA).  acc_dev is [n_classes][257], zeroed by the
 * caller before the first call; chunk_weight = (float)len / total_len.  Replaces the accumulation loop of
 * inference_on_support_set_dataset_base, sylph/evaluation/meta_learn_evaluation.py:190-203. */
int sylph_accumulate_codes(sylph_ctx* ctx, const float* chunk_codes_dev, int n_chunks, const int* chunk_class_host,
                           const float* chunk_weight_host, float* acc_dev, int n_classes, void* stream);

/* Sum `n_parts` partial accumulators ([n_parts][n_classes][257], one per rank, in rank order starting from 0) and
 * divide class c by divisor_host[c] when it is non-zero (the caller sets it to the accumulated weight where
 * |1 - acc_weight| > 1e-6, else 0).  Replaces reduce_class_code, sylph/modeling/code_generator/utils.py:397-427. */
int sylph_reduce_codes(sylph_ctx* ctx, const float* parts_dev, int n_parts, int n_classes, const float* divisor_host,
                       float* codes_out_dev, void* stream);

/* Detection on the query slot with (n_classes, 257) normalised codes.  out_sizes: (n_images, 2) = (height, width)
 * requested output resolution per image (detector_postprocess).  dets_out_dev: (n_images, max_dets, 9) floats,
 * counts_out_dev: (n_images) int32; rows are in descending score order.  max_dets must be >= post_nms_topk.
 * Replaces MetaFCOS.forward + FCOSOutputs.predict_proposals + detector_postprocess:
 * sylph/modeling/meta_fcos/fcos.py:184-268,582-667, fcos_outputs.py:743-812,904-1028,
 * meta_one_stage_detector.py:279-295. */
int sylph_detect(sylph_ctx* ctx, int slot, const float* codes_dev, int n_classes, const int* out_sizes_host,
                 float* dets_out_dev, int* counts_out_dev, int max_dets, void* stream);

/* sylph_detect for callers that produce the class codes on ANOTHER stream.  The class / box towers and the box
 * predictors do not depend on the codes and are enqueued first; `stream` then waits on `codes_ready_event` (a cudaEvent_t
 * cast to void*, recorded by the caller after the last write of codes_dev; NULL = no wait) right before the
 * code-conditioned classifier (CondConvBasic, head_utils.py:60-81) reads codes_dev.  Code generation (ROIAlign, code
 * tower, K-shot mean, normalisation, the all-gather) thereby overlaps the 9 tensor-bound tower launches. */
int sylph_detect_after(sylph_ctx* ctx, int slot, const float* codes_dev, int n_classes, const int* out_sizes_host,
                       float* dets_out_dev, int* counts_out_dev, int max_dets, void* codes_ready_event, void* stream);

/* Health of the finished sylph_detect calls: non-zero (and a message) when an (image, level) candidate list overflowed its
 * 2^22 slots, i.e. detections were dropped -- reachable only at LVIS scale (1203 classes on p3) with a very low threshold.
 * The word travels to pinned host memory behind every detect call, so call this after synchronising on the detections;
 * it never touches the device.  (The reference has no such limit: fcos_outputs.py:960-984 sorts whatever passes.) */
int sylph_detect_poll(sylph_ctx* ctx);

/* Head intermediates of the last sylph_detect call, NCHW fp32: which = 0 logits (n, n_classes, H, W),
 * 1 bbox_reg after scale+ReLU (n, 4, H, W), 2 ctrness (n, 1, H, W), 3 iou (n, 1, H, W). */
int sylph_export_head_output(sylph_ctx* ctx, int which, int level, float* out_dev, void* stream);

/* ---- Episodic TRAINING forward of the proposal generator: losses; the backward of the code generator follows below. ---- */

/* The loss hyper-parameters FCOSOutputs._init_fcos reads (sylph/modeling/meta_fcos/fcos_outputs.py:70-102). */
typedef struct sylph_loss_config {
    float focal_alpha;         /* MODEL.FCOS.LOSS_ALPHA */
    float focal_gamma;         /* MODEL.FCOS.LOSS_GAMMA */
    int center_sample;         /* MODEL.FCOS.CENTER_SAMPLE */
    float pos_radius;          /* MODEL.FCOS.POS_RADIUS */
    int loc_loss_type;         /* MODEL.FCOS.LOC_LOSS_TYPE: 0 "iou", 1 "linear_iou", 2 "giou" */
    int sizes_of_interest[4];  /* MODEL.FCOS.SIZES_OF_INTEREST */
} sylph_loss_config;

#define SYLPH_LOSS_SUMS 5      /* focal-loss sum, positives, sum of centre-ness targets, sum of IoU-loss x target, centre-ness BCE sum */
#define SYLPH_BACKGROUND_ID 100000  /* FCOSOutputs.back_ground_id, fcos_outputs.py:101 */

/* Head forward on the query slot with (n_classes, 257) FINAL codes, ground-truth assignment of every location and the
 * per-rank loss sums.  The ground truths are ALREADY filtered to the episode's classes (MetaProposalNetwork._get_gt,
 * meta_one_stage_detector.py:184-221): gt_boxes_host (n_gt, 4) XYXY absolute pixels, gt_classes_host (n_gt) class
 * ids, gt_offsets_host (n_images + 1) prefix offsets per query image; support_targets_host (n_classes) = the class id
 * of every code row.  sums_out_dev receives SYLPH_LOSS_SUMS doubles.  Optional per-location outputs in the reference's
 * level-first order (level, image, y, x): labels (int64, SYLPH_BACKGROUND_ID for background), target_inds (int64, index
 * into the concatenated ground truths, -1 for an image without any) and reg_targets (fp32 l, t, r, b divided by the
 * level stride); pass NULL to skip.  Replaces MetaFCOS.forward (training branch, fcos.py:184-246) + FCOSOutputs.losses /
 * _get_ground_truth / compute_targets_for_locations / fcos_losses_episodic_learning, fcos_outputs.py:140-349, 351-592,
 * for CODE_GENERATOR.BOX_ON == False and DISTILLATION_LOSS_WEIGHT == 0 (the shipped configs). */
int sylph_fcos_loss_sums(sylph_ctx* ctx, int slot, const float* codes_dev, int n_classes,
                         const int64_t* support_targets_host, const sylph_loss_config* lc, int n_gt,
                         const float* gt_boxes_host, const int64_t* gt_classes_host, const int* gt_offsets_host,
                         double* sums_out_dev, int64_t* labels_out_dev, int64_t* target_inds_out_dev,
                         float* reg_targets_out_dev, void* stream);

/* enabled = 0: sylph_fcos_loss_sums skips the box tower and the box / centre-ness predictors and leaves the three box sums at 0
 * (only the positives are counted) -- the reference returns loss_fcos_cls alone when the box branch is frozen
 * (box_branch_loss_on, fcos_outputs.py:87-92, 626-632), so its outputs are never looked at.  Default 1.  sylph_detect is not affected. */
int sylph_set_loss_box_branch(sylph_ctx* ctx, int enabled);

/* losses_out_dev[0..2] = loss_fcos_cls, loss_fcos_loc, loss_fcos_ctr from this rank's sums.  global_pos_ctr_dev holds
 * {positives, centre-ness target sum} summed over ALL ranks -- the two reduce_sum calls of
 * fcos_losses_episodic_learning, fcos_outputs.py:520-523 and :557-558 -- or NULL for a single process. */
int sylph_fcos_loss_finalize(sylph_ctx* ctx, const double* local_sums_dev, const double* global_pos_ctr_dev,
                             int world_size, float* losses_out_dev, void* stream);

/* ---- Backward of the episodic training step for the CODE GENERATOR. ----
 * Scope: the meta-training configurations that train the hyper-network (the class tower's backward follows further below)
 * (configs/COCO-Detection/Meta-FCOS/Meta-FCOS-finetune-lvis.yaml: BACKBONE.FREEZE, PROPOSAL_GENERATOR.FREEZE_CLS_TOWER,
 * FREEZE_BBOX_BRANCH on, CODE_GENERATOR.FREEZE off).  The only loss that reaches the code generator is loss_fcos_cls
 * (the box losses depend on the frozen box branch alone).  Gradients equal what the reference's
 * `sum(model(batched_inputs).values()).backward()` leaves in `.grad` of every `code_generator.*` parameter
 * (detectron2 SimpleTrainer.run_step on forward_few_shot_detector_training, meta_one_stage_detector.py:325-388).
 * The class tower's backward (FREEZE_CLS_TOWER: False) follows further below; the backbone and the box branch have none. */

/* d loss_fcos_cls / d FINAL class codes, (n_classes, 257) fp32 rows [cls_conv 256 | cls_bias].  Call right after
 * sylph_fcos_loss_sums on the same slot with the same codes (the head's logits and class-tower output of that call
 * are read in place); labels_dev = that call's labels_out_dev; local_sums_dev / global_pos_ctr_dev / world_size as for
 * sylph_fcos_loss_finalize (num_pos_avg, fcos_outputs.py:520-523).  grad_loss_dev: one float, the upstream gradient of the
 * loss (NULL = 1).  Differentiates fvcore's sigmoid_focal_loss_jit (fcos_outputs.py:525-537) and CondConvBasic
 * (sylph/modeling/meta_fcos/head_utils.py:60-81). */
int sylph_fcos_cls_loss_backward(sylph_ctx* ctx, int slot, int n_classes, const int64_t* support_targets_host,
                                 const sylph_loss_config* lc, const int64_t* labels_dev, const double* local_sums_dev,
                                 const double* global_pos_ctr_dev, int world_size, const float* grad_loss_dev,
                                 float* grad_codes_out_dev, void* stream);

/* The code generator's trainable tensors, device fp32 in the reference's state_dict layouts
 * (`code_generator.code_generator_head.*`); NULL where the configuration has no such tensor. */
#define SYLPH_CG_MAX_TOWER 4
typedef struct sylph_codegen_tensors {
    float* tower_w[SYLPH_CG_MAX_TOWER];     /* support_set_shared_tower.{3i}.weight   (256, 256, 3, 3) */
    float* tower_b[SYLPH_CG_MAX_TOWER];     /* support_set_shared_tower.{3i}.bias     (256) */
    float* tower_gn_w[SYLPH_CG_MAX_TOWER];  /* support_set_shared_tower.{3i+1}.weight (256) */
    float* tower_gn_b[SYLPH_CG_MAX_TOWER];  /* support_set_shared_tower.{3i+1}.bias   (256) */
    float* cls_w;                           /* support_set_cls_conv.0.weight (256, 256, 3, 3) */
    float* cls_b;                           /* support_set_cls_conv.0.bias   (256) */
    float* bias_w;                          /* support_set_cls_bias.0.weight (1, 256, 3, 3) */
    float* bias_b;                          /* support_set_cls_bias.0.bias   (1) */
    float* post_norm_w;                     /* post_norm.weight (256) */
    float* post_norm_b;                     /* post_norm.bias   (256) */
    float* conv_scale;                      /* conv_scale.scale (1) */
    float* bias_scale;                      /* bias_scale.scale (1) */
} sylph_codegen_tensors;

/* Backward through code_process_module, compute_code, the pools, support_set_cls_conv / support_set_cls_bias and
 * support_set_shared_tower (sylph/modeling/code_generator/code_generator.py:648-688, 778-875, 941-994) for the ROIs of the last
 * sylph_generate_codes call (n_rois, class_offsets as in that call; its pooled ROI features are read in place).
 * raw_codes_dev = that call's output, grad_codes_dev = gradient with respect to the FINAL codes
 * (sylph_normalize_codes of raw_codes_dev); `params` are read, every non-NULL tensor of `grads` is overwritten
 * (not accumulated).  Deterministic (no atomics). */
int sylph_codegen_backward(sylph_ctx* ctx, int n_rois, int n_classes, const int* class_offsets_host,
                           const float* raw_codes_dev, const float* grad_codes_dev, const sylph_codegen_tensors* params,
                           const sylph_codegen_tensors* grads, void* stream);

/* Re-prepare the code generator's weights after an optimiser step: stage ALL `code_generator.*` tensors with
 * sylph_load_tensor, then call this instead of sylph_finalize_weights (the rest of the model keeps its prepared
 * weights; device buffers are reused).  Synchronises the device. */
int sylph_update_code_generator(sylph_ctx* ctx);

/* The same refresh from DEVICE tensors (the optimiser's parameters), prepared by kernels on `stream`: fp32 OIHW -> the
 * tap-major fp16 (hi | hi | lo in exact mode) operand layout, bit-identical to the host preparation.  Only conv_scale /
 * bias_scale travel to the host (8 bytes; the call waits for the stream). */
int sylph_update_code_generator_device(sylph_ctx* ctx, const sylph_codegen_tensors* params, void* stream);

/* ---- Backward of the FCOS class tower (PROPOSAL_GENERATOR.FREEZE_CLS_TOWER: False, the shipped Meta-FCOS-finetune.yaml
 * configurations of COCO and LVIS train the code generator AND the class tower; the backbone and the box branch stay frozen). ----
 * Reference: autograd of MetaFCOSHead.cls_tower (sylph/modeling/meta_fcos/fcos.py:72-122, 582-667: NUM_CLS_CONVS x
 * [conv3x3 + GroupNorm(32) + ReLU], shared by the five levels) under loss_fcos_cls. */

/* Keep the class tower's activations during the head pass of sylph_fcos_loss_sums (every layer's input planes, pre-GroupNorm
 * output and GroupNorm statistics stay in their own buffers instead of two ping-pong buffers).  Off by default. */
int sylph_set_training(sylph_ctx* ctx, int enabled);

/* The class tower's tensors, device fp32 in the state_dict layouts (`proposal_generator.fcos_head.cls_tower.*`). */
typedef struct sylph_tower_tensors {
    float* conv_w[SYLPH_CG_MAX_TOWER];      /* cls_tower.{3i}.weight   (256, 256, 3, 3) */
    float* conv_b[SYLPH_CG_MAX_TOWER];      /* cls_tower.{3i}.bias     (256) */
    float* gn_w[SYLPH_CG_MAX_TOWER];        /* cls_tower.{3i+1}.weight (256) */
    float* gn_b[SYLPH_CG_MAX_TOWER];        /* cls_tower.{3i+1}.bias   (256) */
} sylph_tower_tensors;

/* Gradients of loss_fcos_cls with respect to the class tower's tensors.  Call after sylph_fcos_loss_sums (training mode on) on
 * the same slot with the same FINAL codes; labels_dev / local_sums_dev / global_pos_ctr_dev / world_size / grad_loss_dev as
 * for sylph_fcos_cls_loss_backward.  Per layer, last first: ReLU + GroupNorm backward over the planes, the weight gradient
 * by a tcgen05 kernel that reads both operands MN-major straight from the planes (csrc/wgrad3x3.cuh), the input gradient
 * by the forward convolution kernel on the transposed, tap-reversed weights.  Every non-NULL tensor of `grads` is
 * overwritten.  Deterministic. */
int sylph_cls_tower_backward(sylph_ctx* ctx, int slot, int n_classes, const float* codes_dev,
                             const int64_t* support_targets_host, const sylph_loss_config* lc, const int64_t* labels_dev,
                             const double* local_sums_dev, const double* global_pos_ctr_dev, int world_size,
                             const float* grad_loss_dev, const sylph_tower_tensors* params,
                             const sylph_tower_tensors* grads, void* stream);

/* Re-prepare the class tower's weights (forward operands and the transposed copies of the backward) from the optimiser's
 * device tensors, by kernels on `stream`. */
int sylph_update_cls_tower_device(sylph_ctx* ctx, const sylph_tower_tensors* params, void* stream);

/* Debugging aid: copy the first `bytes` of one of the context's named scratch buffers (engine.cu `ensure` names, e.g.
 * "bwd.x", "det.logits") to out_dev.  Not part of the reference-facing surface. */
int sylph_debug_read_buffer(sylph_ctx* ctx, const char* name, void* out_dev, size_t bytes, void* stream);

/* Number of kernels this library launched on the context since creation (bench.py's gpu_launches). */
int64_t sylph_launch_count(const sylph_ctx* ctx);

/* Per-stage device time of the most recent profiled call: enable with sylph_set_profiling(ctx, 1); then
 * sylph_get_timings fills up to `cap` (name, ms) pairs and returns how many exist. */
int sylph_set_profiling(sylph_ctx* ctx, int enabled);
int sylph_get_timings(sylph_ctx* ctx, char (*names)[48], float* ms, double* flops, double* bytes, int cap);

#ifdef __cplusplus
}
#endif
#endif /* SYLPH_B200_H_ */
