// TEST INFRASTRUCTURE, NOT TENSORFLOW: see op_kernel.h in this directory.
#ifndef EOT_TESTS_TF_STUB_OP_H_
#define EOT_TESTS_TF_STUB_OP_H_
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
namespace tensorflow {
struct OpDefBuilderStub {
  explicit OpDefBuilderStub(const char* name);
  OpDefBuilderStub& Input(const char* spec);
  OpDefBuilderStub& Output(const char* spec);
  OpDefBuilderStub& Attr(const char* spec);
  OpDefBuilderStub& SetShapeFn(Status (*fn)(shape_inference::InferenceContext*));
};
}  // namespace tensorflow
#define REGISTER_OP(NAME) \
  static ::tensorflow::OpDefBuilderStub& EOT_STUB_CAT(eot_stub_op_, __LINE__) __attribute__((unused)) = ::tensorflow::OpDefBuilderStub(NAME)
#endif
