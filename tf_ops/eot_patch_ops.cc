// TensorFlow custom-op shim over libeotpatch.so (include/eotpatch.h).
//
// STATUS: source only.  TensorFlow is not installed in the build image or on the GPU box, so this file is
// never compiled or tested there; it is what a maintainer of the reference builds next to
// a TensorFlow >= 2.8 install:
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC -O2 tf_ops/eot_patch_ops.cc -o tf_ops/_eot_patch_ops.so \
//       -Iinclude -DGOOGLE_CUDA=1 $TF_CFLAGS $TF_LFLAGS \
//       -Lmladversarialobjectdetection_b200 -l:libeotpatch.so -Wl,-rpath,'$ORIGIN/../mladversarialobjectdetection_b200'
//
// The ops are thin: they allocate outputs / the workspace through the OpKernelContext (the library never
// allocates device memory), take the op's CUDA stream, and forward DEVICE pointers -- zero copies.  Ragged boxes
// arrive as (flat_values [N,4], row_splits [B+1]) of the tf.RaggedTensor the reference already builds
// (attacker.py:160-170).
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

#include "eotpatch.h"

namespace tf = tensorflow;

namespace {

void* StreamOf(tf::OpKernelContext* ctx) {
  return static_cast<void*>(ctx->eigen_device<Eigen::GpuDevice>().stream());
}

#define EOT_OP_OK(ctx, call)                                                              \
  do {                                                                                    \
    int _rc = (call);                                                                     \
    OP_REQUIRES(ctx, _rc == EOT_OK, tf::errors::Internal(#call, ": ", eot_last_error())); \
  } while (0)

EotShape ShapeOf(const tf::Tensor& images, const tf::Tensor& patch, const tf::Tensor& boxes, float tolerance,
                 float noise_amp, float min_patch_area, bool want_mask) {
  EotShape s{};
  s.batch = images.dim_size(0);
  s.height = images.dim_size(1);
  s.width = images.dim_size(2);
  const bool per_image = patch.dims() == 4;
  s.patch_size = patch.dim_size(per_image ? 1 : 0);
  s.num_patches = per_image ? s.batch : 1;
  s.total_boxes = boxes.dim_size(0);
  s.flags = want_mask ? EOT_FLAG_MASK_OUTPUT : 0;
  s.tolerance = tolerance;
  s.noise_amp = noise_amp;
  s.min_patch_area = min_patch_area;
  s.max_scale = 1.0f;
  return s;
}

}  // namespace

REGISTER_OP("EotPatchApply")
    .Input("patch: float")        // [P,P,3] or [B,P,P,3]
    .Input("scale: float")        // []
    .Input("images: float")       // [B,H,W,3]
    .Input("boxes: float")        // [N,4]   RaggedTensor.flat_values
    .Input("row_splits: int32")   // [B+1]   RaggedTensor.row_splits
    .Input("params: uint8")       // [N,48]  EotBoxParams records (transform seeds)
    .Input("print_wb: float")     // [B,6]
    .Attr("tolerance: float = 0.2")
    .Attr("noise_amp: float = 0.01")
    .Attr("min_patch_area: float = 4.0")
    .Attr("want_mask: bool = false")
    .Output("patched: float")     // [B,H,W,3]
    .Output("mask: float")        // [B,H,W,3] (or [0] when want_mask is false)
    .Output("workspace: uint8")   // saved state for EotPatchApplyGrad
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(2));
      c->set_output(1, c->UnknownShape());
      c->set_output(2, c->Vector(c->UnknownDim()));
      return tf::Status::OK();
    });

class EotPatchApplyOp : public tf::OpKernel {
 public:
  explicit EotPatchApplyOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("tolerance", &tolerance_));
    OP_REQUIRES_OK(c, c->GetAttr("noise_amp", &noise_amp_));
    OP_REQUIRES_OK(c, c->GetAttr("min_patch_area", &min_patch_area_));
    OP_REQUIRES_OK(c, c->GetAttr("want_mask", &want_mask_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &patch = ctx->input(0), &scale = ctx->input(1), &images = ctx->input(2), &boxes = ctx->input(3),
                     &splits = ctx->input(4), &params = ctx->input(5), &wb = ctx->input(6);
    OP_REQUIRES(ctx, images.dims() == 4 && images.dim_size(3) == 3, tf::errors::InvalidArgument("images must be [B,H,W,3]"));
    const EotShape s = ShapeOf(images, patch, boxes, tolerance_, noise_amp_, min_patch_area_, want_mask_);
    size_t bytes = 0;
    EOT_OP_OK(ctx, eot_workspace_bytes(&s, &bytes));
    tf::Tensor *out = nullptr, *mask = nullptr, *ws = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, images.shape(), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, want_mask_ ? images.shape() : tf::TensorShape({0}), &mask));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({static_cast<tf::int64>(bytes)}), &ws));
    EOT_OP_OK(ctx, eot_apply_fwd(&s, patch.flat<float>().data(), scale.flat<float>().data(), images.flat<float>().data(),
                                 boxes.flat<float>().data(), splits.flat<tf::int32>().data(),
                                 reinterpret_cast<const EotBoxParams*>(params.flat<tf::uint8>().data()),
                                 wb.flat<float>().data(), out->flat<float>().data(),
                                 want_mask_ ? mask->flat<float>().data() : nullptr, ws->flat<tf::uint8>().data(), bytes,
                                 StreamOf(ctx)));
  }

 private:
  float tolerance_, noise_amp_, min_patch_area_;
  bool want_mask_;
};
REGISTER_KERNEL_BUILDER(Name("EotPatchApply").Device(tf::DEVICE_GPU), EotPatchApplyOp);

REGISTER_OP("EotPatchApplyGrad")
    .Input("patch: float")
    .Input("print_wb: float")
    .Input("grad_patched: float")   // [B,H,W,3]
    .Input("workspace: uint8")
    .Input("images_shape: int32")   // [4]
    .Input("num_boxes: int32")      // []
    .Attr("tolerance: float = 0.2")
    .Attr("noise_amp: float = 0.01")
    .Attr("min_patch_area: float = 4.0")
    .Output("grad_patch: float")    // [P,P,3]
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      return tf::Status::OK();
    });

class EotPatchApplyGradOp : public tf::OpKernel {
 public:
  explicit EotPatchApplyGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("tolerance", &tolerance_));
    OP_REQUIRES_OK(c, c->GetAttr("noise_amp", &noise_amp_));
    OP_REQUIRES_OK(c, c->GetAttr("min_patch_area", &min_patch_area_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &patch = ctx->input(0), &wb = ctx->input(1), &g = ctx->input(2), &ws = ctx->input(3);
    EotShape s{};
    s.batch = g.dim_size(0);
    s.height = g.dim_size(1);
    s.width = g.dim_size(2);
    s.patch_size = patch.dim_size(0);
    s.num_patches = 1;
    // host-memory inputs (see HostMemory below)
    s.total_boxes = ctx->input(5).scalar<tf::int32>()();
    s.tolerance = tolerance_;
    s.noise_amp = noise_amp_;
    s.min_patch_area = min_patch_area_;
    s.max_scale = 1.0f;
    tf::Tensor* gp = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, patch.shape(), &gp));
    // the workspace is an input tensor: TF inputs are immutable, the backward only writes its own scratch regions
    EOT_OP_OK(ctx, eot_apply_bwd(&s, patch.flat<float>().data(), wb.flat<float>().data(), g.flat<float>().data(),
                                 const_cast<tf::uint8*>(ws.flat<tf::uint8>().data()), ws.NumElements(),
                                 gp->flat<float>().data(), /*accumulate=*/0, StreamOf(ctx)));
  }

 private:
  float tolerance_, noise_amp_, min_patch_area_;
};
REGISTER_KERNEL_BUILDER(Name("EotPatchApplyGrad").Device(tf::DEVICE_GPU).HostMemory("images_shape").HostMemory("num_boxes"),
                        EotPatchApplyGradOp);

REGISTER_OP("EotScoreMax")
    .Input("cls_levels: num_levels * float")   // each [B,h,w,9*C]
    .Input("box_levels: num_levels * float")   // each [B,h,w,36]
    .Input("anchors: float")                   // [A,4]
    .Input("scale: float")
    .Attr("num_levels: int = 5")
    .Attr("num_classes: int = 90")
    .Attr("image_height: float")
    .Attr("image_width: float")
    .Output("max_scores: float")               // [B]
    .Output("argmax_anchor: int32")            // [B]
    .Output("num_candidates: int32")           // [B]
    .Output("dcls_levels: num_levels * float") // d loss / d cls level (dense, zero filled)
    .Output("dscale: float")
    .Output("loss: float");

class EotScoreMaxOp : public tf::OpKernel {
 public:
  explicit EotScoreMaxOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("num_levels", &num_levels_));
    OP_REQUIRES_OK(c, c->GetAttr("num_classes", &num_classes_));
    OP_REQUIRES_OK(c, c->GetAttr("image_height", &h_));
    OP_REQUIRES_OK(c, c->GetAttr("image_width", &w_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    ScoreShape s{};
    const float* cls[SCORE_MAX_LEVELS];
    const float* box[SCORE_MAX_LEVELS];
    float* dcls[SCORE_MAX_LEVELS];
    s.num_levels = num_levels_;
    s.num_classes = num_classes_;
    s.batch = ctx->input(0).dim_size(0);
    s.anchors_per_loc = ctx->input(0).dim_size(3) / num_classes_;
    int tot = 0;
    for (int l = 0; l < num_levels_; ++l) {
      const tf::Tensor& c = ctx->input(l);
      cls[l] = c.flat<float>().data();
      box[l] = ctx->input(num_levels_ + l).flat<float>().data();
      s.level_locs[l] = c.dim_size(1) * c.dim_size(2);
      tot += s.level_locs[l];
      tf::Tensor* d = nullptr;
      OP_REQUIRES_OK(ctx, ctx->allocate_output(3 + l, c.shape(), &d));
      dcls[l] = d->flat<float>().data();
    }
    s.total_anchors = tot * s.anchors_per_loc;
    s.image_height = h_;
    s.image_width = w_;
    s.min_area = 100.0f;
    const tf::Tensor& anchors = ctx->input(2 * num_levels_);
    const tf::Tensor& scale = ctx->input(2 * num_levels_ + 1);
    size_t bytes = 0;
    EOT_OP_OK(ctx, score_workspace_bytes(&s, &bytes));
    tf::Tensor ws, *M = nullptr, *am = nullptr, *nc = nullptr, *dscale = nullptr, *loss = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(bytes)}), &ws));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({s.batch}), &M));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({s.batch}), &am));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({s.batch}), &nc));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3 + num_levels_, tf::TensorShape({}), &dscale));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4 + num_levels_, tf::TensorShape({}), &loss));
    void* st = StreamOf(ctx);
    EOT_OP_OK(ctx, score_max_fwd(&s, cls, box, anchors.flat<float>().data(), M->flat<float>().data(),
                                 am->flat<tf::int32>().data(), nc->flat<tf::int32>().data(), ws.flat<tf::uint8>().data(),
                                 bytes, st));
    EOT_OP_OK(ctx, score_max_bwd(&s, cls, M->flat<float>().data(), scale.flat<float>().data(), dcls,
                                 dscale->flat<float>().data(), loss->flat<float>().data(), ws.flat<tf::uint8>().data(), bytes, st));
  }

 private:
  int num_levels_, num_classes_;
  float h_, w_;
};
REGISTER_KERNEL_BUILDER(Name("EotScoreMax").Device(tf::DEVICE_GPU), EotScoreMaxOp);

// First pass after the victim (attacker.py:100-116,143-170): person candidates -> NonMaxSuppressionV5 -> clip_boxes ->
// ragged boxes, in one op instead of tf.map_fn over the images with the CPU-only NMS kernel.
REGISTER_OP("EotFirstPassBoxes")
    .Input("cls_levels: num_levels * float")   // each [B,h,w,9*C]
    .Input("box_levels: num_levels * float")   // each [B,h,w,36]
    .Input("anchors: float")                   // [A,4]
    .Attr("num_levels: int = 5")
    .Attr("num_classes: int = 90")
    .Attr("image_height: float")
    .Attr("image_width: float")
    .Attr("max_output_size: int = 100")
    .Attr("iou_threshold: float = 1.0")        // gaussian: 1.0; hard: nms_configs.iou_thresh
    .Attr("score_threshold: float = 0.5")      // nms_configs.score_thresh (attacker_train.py:31)
    .Attr("soft_nms_sigma: float = 0.25")      // gaussian: sigma / 2 (tf2/postprocess.py:196-199); hard: 0
    .Attr("score_floor: float = 0.5")          // filter_valid_boxes(thresh=True), attacker.py:87-88
    .Output("nms_boxes: float")                // [B,max_output_size,4], selection order, zero padded
    .Output("nms_scores: float")               // [B,max_output_size]
    .Output("valid_len: int32")                // [B]
    .Output("row_splits: int32")               // [B+1]
    .Output("ragged_boxes: float")             // [B*max_output_size,4], first row_splits[B] rows valid
    .Output("ragged_scores: float");           // [B*max_output_size]

class EotFirstPassBoxesOp : public tf::OpKernel {
 public:
  explicit EotFirstPassBoxesOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("num_levels", &num_levels_));
    OP_REQUIRES_OK(c, c->GetAttr("num_classes", &num_classes_));
    OP_REQUIRES_OK(c, c->GetAttr("image_height", &h_));
    OP_REQUIRES_OK(c, c->GetAttr("image_width", &w_));
    OP_REQUIRES_OK(c, c->GetAttr("max_output_size", &max_out_));
    OP_REQUIRES_OK(c, c->GetAttr("iou_threshold", &iou_));
    OP_REQUIRES_OK(c, c->GetAttr("score_threshold", &thr_));
    OP_REQUIRES_OK(c, c->GetAttr("soft_nms_sigma", &sigma_));
    OP_REQUIRES_OK(c, c->GetAttr("score_floor", &floor_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    ScoreShape s{};
    NmsShape n{};
    const float* cls[SCORE_MAX_LEVELS];
    const float* box[SCORE_MAX_LEVELS];
    s.num_levels = n.num_levels = num_levels_;
    s.num_classes = num_classes_;
    s.batch = n.batch = ctx->input(0).dim_size(0);
    s.anchors_per_loc = ctx->input(0).dim_size(3) / num_classes_;
    int tot = 0;
    for (int l = 0; l < num_levels_; ++l) {
      const tf::Tensor& c = ctx->input(l);
      cls[l] = c.flat<float>().data();
      box[l] = ctx->input(num_levels_ + l).flat<float>().data();
      s.level_locs[l] = c.dim_size(1) * c.dim_size(2);
      n.level_anchors[l] = s.level_locs[l] * s.anchors_per_loc;
      tot += s.level_locs[l];
    }
    s.total_anchors = n.total_anchors = tot * s.anchors_per_loc;
    s.image_height = n.image_height = h_;
    s.image_width = n.image_width = w_;
    s.min_area = 100.0f;
    n.max_output_size = max_out_;
    n.max_candidates = 0;
    n.iou_threshold = iou_; n.score_threshold = thr_; n.soft_nms_sigma = sigma_; n.score_floor = floor_;
    const tf::Tensor& anchors = ctx->input(2 * num_levels_);
    size_t sbytes = 0, nbytes = 0;
    EOT_OP_OK(ctx, score_workspace_bytes(&s, &sbytes));
    EOT_OP_OK(ctx, person_nms_workspace_bytes(&n, &nbytes));
    tf::Tensor sws, nws, M, am, nc;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(sbytes)}), &sws));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(nbytes)}), &nws));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({s.batch}), &M));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT32, tf::TensorShape({s.batch}), &am));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT32, tf::TensorShape({s.batch}), &nc));
    tf::Tensor *boxes = nullptr, *scores = nullptr, *vlen = nullptr, *splits = nullptr, *rboxes = nullptr, *rscores = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({s.batch, max_out_, 4}), &boxes));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({s.batch, max_out_}), &scores));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({s.batch}), &vlen));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, tf::TensorShape({s.batch + 1}), &splits));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, tf::TensorShape({s.batch * max_out_, 4}), &rboxes));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(5, tf::TensorShape({s.batch * max_out_}), &rscores));
    void* st = StreamOf(ctx);
    EOT_OP_OK(ctx, score_max_fwd(&s, cls, box, anchors.flat<float>().data(), M.flat<float>().data(),
                                 am.flat<tf::int32>().data(), nc.flat<tf::int32>().data(), sws.flat<tf::uint8>().data(), sbytes, st));
    // the dense candidate scores score_max_fwd left in its workspace
    size_t cand_off = 0;
    EOT_OP_OK(ctx, score_candidate_offset(&s, &cand_off));
    const float* cand = reinterpret_cast<const float*>(sws.flat<tf::uint8>().data() + cand_off);
    EOT_OP_OK(ctx, person_nms(&n, cand, box, anchors.flat<float>().data(), boxes->flat<float>().data(),
                              scores->flat<float>().data(), vlen->flat<tf::int32>().data(), splits->flat<tf::int32>().data(),
                              rboxes->flat<float>().data(), rscores->flat<float>().data(), nws.flat<tf::uint8>().data(), nbytes, st));
  }

 private:
  int num_levels_, num_classes_, max_out_;
  float h_, w_, iou_, thr_, sigma_, floor_;
};
REGISTER_KERNEL_BUILDER(Name("EotFirstPassBoxes").Device(tf::DEVICE_GPU), EotFirstPassBoxesOp);
