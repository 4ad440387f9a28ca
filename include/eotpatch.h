/*
 * libeotpatch -- C ABI of the B200-native EOT patch-attack hot path.
 *
 * The reference exposes this path as Python/Keras layers, not as an FFI; each
 * entry point below names the reference interface it stands in for.  All pointers are DEVICE
 * pointers unless stated otherwise; every call is asynchronous on the caller's stream, never
 * allocates or frees device memory, never synchronises, keeps no pointer after it returns and
 * has no mutable global state (reentrant across host threads / streams).
 *
 * Return value: EOT_OK (0) or an EotStatus code; the message of the last failure on the calling
 * thread is available through eot_last_error().
 *
 * Tensors are float32, NHWC, contiguous.  Ragged per-image boxes are CSR: boxes[N,4] with
 * box_offsets[B+1].
 */
#ifndef EOTPATCH_H_
#define EOTPATCH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum EotStatus {
  EOT_OK = 0,
  EOT_ERR_NULL_POINTER = 1,
  EOT_ERR_BAD_SHAPE = 2,
  EOT_ERR_WORKSPACE_TOO_SMALL = 3,
  EOT_ERR_CUDA = 4,
  EOT_ERR_MISALIGNED = 5,
  EOT_ERR_GEOMETRY = 6 /* eot_check_workspace: a patch window did not fit the image */
} EotStatus;

/* Transform seeds of one person box: everything `Patcher.create` / `add_patch_to_image` draw from
 * TF's RNG (attacker.py:426-427,436,473-474; Masker: attack_detection.py:411,421,453).  48 bytes. */
typedef struct EotBoxParams {
  float uy, ux;        /* unit uniforms of the centre jitter U(-tol*h/2, tol*h/2), (.. w ..)        */
  float delta;         /* tf.image.random_brightness delta, U[-.3,.3)                                */
  float cos_t, sin_t;  /* rotation angle U[-20deg,20deg) as (cos, sin), evaluated by the caller      */
  float pa, pb;        /* projective row of the 8-parameter transform (0,0 = the reference rotation) */
  float scale;         /* >= 0: per-box patch scale (Masker training U(.3,.5)); < 0: shared *scale   */
  uint32_t key0, key1; /* Philox4x32-10 key of the per-texel noise tf.random.uniform(+-noise_amp)    */
  uint32_t rsv0, rsv1;
} EotBoxParams;

#define EOT_FLAG_MASK_OUTPUT 1u /* Masker: also write mask = original - pasted (attack_detection.py:429-430) */
#define EOT_FLAG_SERIAL_ADJOINT 2u /* accepted for compatibility: the backward always accumulates per image in shared
                                     memory (no per-box partial buffers) since version 101 */

typedef struct EotShape {
  int32_t batch;          /* B images held by this rank                                            */
  int32_t height, width;  /* H, W                                                                  */
  int32_t patch_size;     /* P: patch texture is [P,P,3]                                           */
  int32_t num_patches;    /* 1: one shared patch (Patcher); B: one texture per image (Masker train) */
  int32_t total_boxes;    /* N: box capacity of the call, >= box_offsets[B] (the count in use is read on the
                           * device: a caller whose boxes come out of a kernel needs no host read)      */
  uint32_t flags;         /* EOT_FLAG_*                                                            */
  float tolerance;        /* centre jitter fraction: .2 Patcher (attacker.py:465), .5/0 Masker     */
  float noise_amp;        /* .01 Patcher (attacker.py:426), .1 Masker (attack_detection.py:411)    */
  float min_patch_area;   /* 4 (attacker.py:347,392)                                               */
  float max_scale;        /* upper bound of scale used to size the workspace (<=0: 1.0)            */
  /* element strides of the patch view (channel stride is 1); 0 = contiguous [num_patches,P,P,3].
   * Negative strides express the Masker's flips of `images[:, :240, :240]` without a copy.       */
  int64_t patch_stride_n, patch_stride_y, patch_stride_x;
} EotShape;

/* Integer placement of a patch as the reference computes it (attacker.py:418-420,431-434). */
typedef struct EotBoxGeometry {
  int32_t y0, x0, ps, d, pad_lo, pad_hi, valid, span;
} EotBoxGeometry;

/* What the transform draw needs (eot_draw_transforms): the reference's distributions with its constants as
 * defaults -- max_angle 20 deg (attacker.py:436), max_delta .3 (:427), perspective 0 (no projective row),
 * scale_lo < 0: shared trainable scale; Masker training: scale ~ U(.3,.5), i.e. (.3, .2) (attack_detection.py:453). */
typedef struct EotDrawConfig {
  int64_t seed, step;      /* layer seed and call counter: the hash key                                  */
  int64_t first_image;     /* global index of this rank's first image (data-parallel sharding)           */
  float max_angle, max_delta, perspective;
  float scale_lo, scale_span; /* scale ~ scale_lo + U[0,1) * scale_span                                */
  float rsv;
} EotDrawConfig;

const char* eot_last_error(void);
int eot_version(void);
/* Number of CUDA kernels this library has launched since it was loaded (all threads). */
uint64_t eot_launch_count(void);

/* Bytes of the saved-state workspace eot_apply_fwd fills and eot_apply_bwd reads. */
int eot_workspace_bytes(const EotShape* shape, size_t* bytes);

/* `Patcher.create` + area filter + int cast for every box (attacker.py:392-394,418,448-488),
 * by the same device code eot_apply_fwd uses.  geometry_out: [N] EotBoxGeometry. */
int eot_box_geometry(const EotShape* shape, const float* boxes, const int32_t* box_offsets,
                     const EotBoxParams* params, const float* scale, EotBoxGeometry* geometry_out,
                     void* stream);

/* Everything `Patcher` / `Masker` draw from TF's RNG inside the graph (attacker.py:370-371, 426-427, 436, 473-474;
 * attack_detection.py:350-351, 411, 421, 451-453), as one launch: params_out[box_capacity] (slots past
 * box_offsets[B] zero-filled) and print_wb_out[B,6].  Counter-based on (seed, step, global image, box in image):
 * a sharded batch draws what the single-GPU batch draws. */
int eot_draw_transforms(const EotDrawConfig* cfg, int32_t batch, int32_t box_capacity,
                        const int32_t* box_offsets, EotBoxParams* params_out, float* print_wb_out,
                        void* stream);

/* `Patcher.call` (attacker.py:490-498) / `Masker.call` (attack_detection.py:478-498):
 * out_images[B,H,W,3] = images with every valid box patched in order; out_masks (optional,
 * needs EOT_FLAG_MASK_OUTPUT) as the Masker's second output.  `patch` points at element
 * [0,0,0,0] of the (possibly strided) patch view; `scale` is a device scalar (the trainable
 * scale_regressor); print_wb is [B,6] = (w0,w1,w2,b0,b1,b2) of random_print_adjust.
 * out_images may alias images (in-place). */
int eot_apply_fwd(const EotShape* shape, const float* patch, const float* scale, const float* images,
                  const float* boxes, const int32_t* box_offsets, const EotBoxParams* params,
                  const float* print_wb, float* out_images, float* out_masks, void* workspace,
                  size_t workspace_bytes, void* stream);

/* `tape.gradient(loss, patch)` through the patcher (attacker.py:217; chain of SURVEY.md 3.2):
 * grad_images = dL/d(out_images) [B,H,W,3]; grad_patch [P,P,3] (shared patch only).
 * accumulate != 0 adds into grad_patch, otherwise it is overwritten.  Needs the workspace of the
 * matching eot_apply_fwd call, unmodified. */
int eot_apply_bwd(const EotShape* shape, const float* patch, const float* print_wb,
                  const float* grad_images, void* workspace, size_t workspace_bytes,
                  float* grad_patch, int accumulate, void* stream);

/* `BrightnessMatcher()((src, tgt))` on its own (brightness_matcher.py:43-73): src [src_pixels,3],
 * tgt [tgt_pixels,3] -> out [src_pixels,3].  workspace: 16 bytes, 8-byte aligned. */
int eot_brightness_match(const float* src, int64_t src_pixels, const float* tgt, int64_t tgt_pixels,
                         float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Synchronises `stream` and reports whether any box failed the reference's implicit shape
 * requirements during the last eot_apply_fwd on this workspace (debug / tests only). */
int eot_check_workspace(const EotShape* shape, const void* workspace, void* stream);

/* --------------------------------------------------------------------------------------------
 * Person-score objective: pre_nms (max-reduce branch) + person filter + valid-box filter +
 * per-image max (attacker.py:118-141,190; tf2/postprocess.py:67-79,104-116,136-156;
 * tf2/anchors.py:30-58).
 * ------------------------------------------------------------------------------------------ */
#define SCORE_MAX_LEVELS 8

typedef struct ScoreShape {
  int32_t batch;
  int32_t num_levels;                       /* 5 for min_level 3 .. max_level 7              */
  int32_t num_classes;                      /* 90                                            */
  int32_t anchors_per_loc;                  /* 9                                             */
  int32_t level_locs[SCORE_MAX_LEVELS];     /* H_l * W_l                                     */
  int32_t total_anchors;                    /* A = 9 * sum(level_locs)                       */
  float image_height, image_width;          /* for the valid-box filter (attacker.py:79-83)  */
  float min_area;                           /* 100 (attacker.py:88)                          */
} ScoreShape;

int score_workspace_bytes(const ScoreShape* shape, size_t* bytes);
/* Byte offset, inside the score workspace, of the dense candidate scores score_max_fwd leaves there:
 * float [B, A], the person score of a valid candidate anchor or -1 (the reference's ragged `scores` of
 * attacker.py:134-139 in dense form).  It is what person_nms takes as `cand_score`; callers ask for the
 * offset instead of restating the workspace layout.  Host-only, no launch. */
int score_candidate_offset(const ScoreShape* shape, size_t* offset);

/* cls_levels / box_levels: HOST arrays of num_levels device pointers, level l is
 * [B, H_l, W_l, 9*num_classes] / [B, H_l, W_l, 9*4].  anchors: [A,4].
 * Outputs: max_scores[B] = maximum(reduce_max(candidate scores), 0); argmax_anchor[B] = lowest
 * anchor index attaining it (-1: no candidate); num_candidates[B]. */
int score_max_fwd(const ScoreShape* shape, const float* const* cls_levels,
                  const float* const* box_levels, const float* anchors, float* max_scores,
                  int32_t* argmax_anchor, int32_t* num_candidates, void* workspace,
                  size_t workspace_bytes, void* stream);

/* Dense, zero-filled dL/dcls per level (what the framework's conv backward consumes) for
 * loss = sum_b(M_b^2 + (M_b - scale)^2) (attacker.py:191,193): one non-zero per image, ties split
 * equally as TF's Max / UnsortedSegmentMax gradients do.  dscale_out: device scalar,
 * sum_b -2 (M_b - scale).  loss_out (optional): device scalar with the data term of the loss. */
int score_max_bwd(const ScoreShape* shape, const float* const* cls_levels, const float* max_scores,
                  const float* scale, float* const* dcls_levels, float* dscale_out,
                  float* loss_out, void* workspace, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * First-pass post-processing (attacker.py:100-116,143-170): the person candidates score_max_fwd
 * left in its workspace -> per-image NonMaxSuppressionV5 (tf2/postprocess.py:159-205; hard or
 * Gaussian soft-NMS, TF's lazy priority-queue order) -> clip_boxes -> ragged boxes as CSR.
 * Replaces the reference's host-synchronous tf.map_fn over the images + CPU-only NMS kernel.
 * ------------------------------------------------------------------------------------------ */
typedef struct NmsShape {
  int32_t batch;
  int32_t total_anchors;                    /* A                                                  */
  int32_t num_levels;
  int32_t max_output_size;                  /* nms_configs.max_output_size (100); <= 128          */
  int32_t max_candidates;                   /* candidates per image the workspace holds; 0 = A    */
  int32_t level_anchors[SCORE_MAX_LEVELS];  /* anchors per image of level l (9 * H_l * W_l)       */
  float iou_threshold;                      /* hard: nms_configs.iou_thresh; gaussian: 1.0        */
  float score_threshold;                    /* NMS score threshold (>= 0)                         */
  float soft_nms_sigma;                     /* 0: hard NMS; gaussian: sigma / 2                   */
  float score_floor;                        /* filter_valid_boxes(thresh=True): score >= floor    */
  float image_height, image_width;          /* clip_boxes                                         */
} NmsShape;

int person_nms_workspace_bytes(const NmsShape* shape, size_t* bytes);

/* cand_score: [B,A] candidate score or -1 (score_max_fwd's workspace; see ScoreShape).  box_levels:
 * HOST array of num_levels device pointers [B, H_l, W_l, 9*4].  anchors: [A,4].
 * Outputs (all device): nms_boxes [B,max_output_size,4] / nms_scores [B,max_output_size] in
 * selection order, zero padded (scores are the soft-NMS adjusted ones, as TF returns them);
 * valid_len [B]; row_splits [B+1] and ragged_boxes [B*max_output_size,4] / ragged_scores (optional)
 * = the same boxes compacted in image order.  An image with more candidates than max_candidates
 * sets valid_len[b] = -1 and row_splits[B] = -1 (nothing is truncated silently). */
int person_nms(const NmsShape* shape, const float* cand_score, const float* const* box_levels,
               const float* anchors, float* nms_boxes, float* nms_scores, int32_t* valid_len,
               int32_t* row_splits, float* ragged_boxes, float* ragged_scores, void* workspace,
               size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * Input pipeline (train_data_generator.py:55-75 `DataSequence._map_fn`; :201-226 augmentation):
 * decoded uint8 frames -> the [B,H,W,3] float32 batch the attack step consumes.
 * ------------------------------------------------------------------------------------------ */
/* frames: HOST array of `batch` DEVICE pointers, frame i is uint8 [heights[i], widths[i], 3]
 * contiguous; heights / widths / mean_rgb[3] / stddev_rgb[3]: HOST arrays.  out: [B,H,W,3]:
 * (v - mean) / stddev in float64, aspect-preserving bilinear resize (cv2.resize INTER_LINEAR on
 * the float64 image) into the top-left corner, zero padding, one rounding to float32.
 * channel_sums (optional, device double[B,3]): per-image per-channel sums of `out` for
 * eot_augment_batch. */
int eot_letterbox_normalize(const uint8_t* const* frames, const int32_t* heights,
                            const int32_t* widths, int32_t batch, int32_t out_height,
                            int32_t out_width, const double* mean_rgb, const double* stddev_rgb,
                            float* out, double* channel_sums, void* stream);

/* channel_sums[B,3] (device double) of a [B,H,W,3] batch that did not come from the call above. */
int eot_channel_sums(const float* images, int32_t batch, int32_t height, int32_t width,
                     double* channel_sums, void* stream);

/* tf.image.random_flip_left_right + RandomFlip('horizontal') (flip[b] != 0: mirror image b; NULL:
 * none), RandomContrast ((x - mean_hw) * contrast_factor + mean_hw per image and channel),
 * random_brightness (+ brightness_delta), clip to [-1,1] (train_data_generator.py:218-222); the
 * random draws are the caller's.  Out of place when flip is given. */
int eot_augment_batch(const float* images, float* out, int32_t batch, int32_t height, int32_t width,
                      const uint8_t* flip, const double* channel_sums, float contrast_factor,
                      float brightness_delta, void* stream);

/* --------------------------------------------------------------------------------------------
 * uint8 inference-time patcher: GPU twin of `adv_patch.AdversarialPatch` (adv_patch.py:40-201).
 * Frames and patches are uint8 [h,w,3] RGB on the device; results are bit-identical to the
 * reference's NumPy + OpenCV code (8-bit fixed-point colour conversion, INTER_LINEAR letter-box,
 * INTER_AREA patch down-sampling, INTER_CUBIC patch up-sampling as OpenCV's own 8-bit kernel computes
 * it -- pip wheels run that one call through Intel IPP, which differs by one grey level on ~4 % of
 * the elements).
 * ------------------------------------------------------------------------------------------ */
/* `_create` (adv_patch.py:61-92) for n boxes; HOST in, HOST out: placements[n,4] =
 * (ymin_patch, xmin_patch, patch_h, patch_w).  No GPU work (sizes the noise buffer). */
int adv_u8_box_geometry(int32_t frame_h, int32_t frame_w, double scale, const double* boxes,
                        int32_t n, int32_t* placements);

/* `print_patch` (adv_patch.py:40-59) on n_elems uint8 values. */
int adv_u8_print_patch(const uint8_t* patch, uint8_t* printed, int64_t n_elems, void* stream);

int adv_u8_workspace_bytes(int32_t patch_h, int32_t patch_w, size_t* bytes);

/* `add_adv_to_img` (adv_patch.py:179-190), in place on `frame`: boxes (HOST double[n,4]: ymin,
 * xmin, ymax, xmax) are pasted in order and every brightness match sees the earlier pastes.
 * out_h/out_w: the `h, w` of AdversarialPatch (letter-box size of `rescale`).  noise: DEVICE
 * float64, the np.random.uniform(-.01, .01) draws of `random_noise`, box after box, each
 * [patch_h_i, patch_w_i, 3]. */
int adv_u8_add_patches(uint8_t* frame, int32_t frame_h, int32_t frame_w, const uint8_t* patch_printed,
                       int32_t patch_h, int32_t patch_w, int32_t out_h, int32_t out_w, double scale,
                       const double* boxes, int32_t n, const double* noise, void* workspace,
                       size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * Patch update (attacker.py:191-193,307-316,51-54): total-variation term and Adam + constraint.
 * ------------------------------------------------------------------------------------------ */
/* grad_patch += weight * d TV(patch)/d patch ; tv_out (optional device scalar) = TV(patch). */
int patch_tv_grad(const float* patch, int32_t patch_size, float weight, float* grad_patch,
                  float* tv_out, void* stream);

/* Keras Adam step on `n` floats followed by clip to [lo,hi] (the tf.Variable constraint).
 * step is the 1-based iteration. */
int adam_clip_update(float* var, float* m, float* v, const float* grad, int64_t n, float lr,
                     float beta1, float beta2, float eps, int64_t step, float lo, float hi,
                     void* stream);

/* Scalars of the step around the data-parallel exchange (attacker.py:190-201, 217): out4 = [dL/dscale, data loss,
 * sum_b M_b, sum_b M_b^2] -- the tail of the packed buffer [dL/dpatch | out4] that is sum-all-reduced. */
int attack_pack_scalars(const float* max_scores, int32_t batch, const float* dscale, const float* data_loss,
                        float* out4, void* stream);
/* out6 = [loss, scale_loss, mean_max_score, std_max_score, tv_loss, scale] from the (reduced) tail, TV(patch) and the
 * scale variable: the add_metric values of attacker.py:196-201 as device scalars. */
int attack_step_metrics(const float* tail4, const float* tv, const float* scale, float global_batch, float tv_weight,
                        float* out6, void* stream);

/* --------------------------------------------------------------------------------------------
 * Epilogue helpers of the torch stand-in of the victim (victim.py) -- not a reference interface:
 * the convolutions stay on the framework's cuDNN path; the per-channel bias add and the SiLU that
 * PyTorch runs as two separate (partly non-vectorised) passes are one pass over the NHWC tensor.
 * x / y / dy / dx: [n_pixels, channels] float32, channels % 4 == 0 (the forward also takes even
 * channel counts), 16-byte aligned; y may alias x.
 * ------------------------------------------------------------------------------------------ */
int nhwc_bias_act_fwd(const float* x, const float* bias, float* y, int64_t n_pixels,
                      int32_t channels, int32_t act /* 0 identity, 1 SiLU */, void* stream);
int nhwc_bias_silu_bwd(const float* x, const float* bias, const float* dy, float* dx,
                       int64_t n_pixels, int32_t channels, void* stream);
/* Squeeze-and-excitation gate of MBConv on [n_images, hw, channels]: out = y * gate[n,c]
 * (+ shift[n,c] * shift_mul when shift != NULL: the backward's dy = dout * gate + dmean / HW). */
int nhwc_channel_scale(const float* y, const float* gate, const float* shift, float shift_mul,
                       float* out, int32_t n_images, int64_t hw, int32_t channels, void* stream);
/* out[n,c] = sum_hw a * b (dgate of the product above), deterministic two-stage reduction;
 * workspace: 16 * n_images * channels floats. */
int nhwc_channel_dot(const float* a, const float* b, float* out, float* workspace, int32_t n_images,
                     int32_t hw, int32_t channels, void* stream);
/* BiFPN fast normalised fusion: out = silu(sum_i weights[i] * xs[i]) for n = 2 or 3 same-layout
 * inputs (xs / dxs: HOST arrays of device pointers; weights: device [n]); backward
 * dxs[i] = weights[i] * dout * silu'(z) (dxs[i] may be NULL to skip an input). */
int nhwc_fuse_silu_fwd(const float* const* xs, int32_t n, const float* weights, float* out,
                       int64_t n_elems, void* stream);
int nhwc_fuse_silu_bwd(const float* const* xs, int32_t n, const float* weights, const float* dout,
                       float* const* dxs, int64_t n_elems, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EOTPATCH_H_ */
