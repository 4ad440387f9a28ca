// TEST INFRASTRUCTURE, NOT TENSORFLOW: see op_kernel.h in this directory.
#ifndef EOT_TESTS_TF_STUB_SHAPE_INFERENCE_H_
#define EOT_TESTS_TF_STUB_SHAPE_INFERENCE_H_
namespace tensorflow {
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
class InferenceContext {
 public:
  ShapeHandle input(int idx) const;
  void set_output(int idx, ShapeHandle shape);
  ShapeHandle UnknownShape();
  DimensionHandle UnknownDim();
  ShapeHandle Vector(DimensionHandle dim);
};
}  // namespace shape_inference
}  // namespace tensorflow
#endif
