// TEST INFRASTRUCTURE, NOT TENSORFLOW.  A declarations-only stand-in for the handful of TensorFlow C++ API names
// tf_ops/eot_patch_ops.cc uses, so that `g++ -fsyntax-only` can type-check the shim's calls into include/eotpatch.h
// (argument counts, pointer types, constness) in an image without TensorFlow (tests/test_tf_shim_typecheck.py).
// It proves nothing about TensorFlow itself; signatures follow the TF 2.8 headers of the same names
// (tensorflow/core/framework/op_kernel.h, tensor.h, tensor_shape.h, op.h, shape_inference.h).
#ifndef EOT_TESTS_TF_STUB_OP_KERNEL_H_
#define EOT_TESTS_TF_STUB_OP_KERNEL_H_
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>

namespace Eigen {
struct GpuDevice {
  void* stream() const;   // cudaStream_t in the real header
};
}  // namespace Eigen

namespace tensorflow {
using int32 = std::int32_t;
using int64 = long long;
using uint8 = std::uint8_t;
enum DataType { DT_FLOAT = 1, DT_INT32 = 3, DT_UINT8 = 4 };
extern const char* const DEVICE_GPU;

class Status {
 public:
  static Status OK();
  bool ok() const;
};

namespace errors {
template <typename... A> Status Internal(A...);
template <typename... A> Status InvalidArgument(A...);
}  // namespace errors

class TensorShape {
 public:
  TensorShape();
  TensorShape(std::initializer_list<int64> dims);
};

template <typename T>
struct FlatView {
  T* data() const;
};
template <typename T>
struct ScalarView {
  T& operator()() const;
};

class Tensor {
 public:
  Tensor();
  int dims() const;
  int64 dim_size(int d) const;
  int64 NumElements() const;
  const TensorShape& shape() const;
  template <typename T> FlatView<T> flat();
  template <typename T> FlatView<const T> flat() const;
  template <typename T> ScalarView<const T> scalar() const;
};

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char* name, T* value) const;
  void CtxFailure(const Status& s);
  void CtxFailureWithWarning(const Status& s);
};

class OpKernelContext {
 public:
  const Tensor& input(int index) const;
  Status allocate_output(int index, const TensorShape& shape, Tensor** tensor);
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp);
  template <typename Device> const Device& eigen_device() const;
  void CtxFailure(const Status& s);
  void CtxFailureWithWarning(const Status& s);
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction* context);
  virtual ~OpKernel();
  virtual void Compute(OpKernelContext* context) = 0;
};

namespace register_kernel {
struct Name {
  explicit Name(const char* op);
  Name& Device(const char* device);
  Name& HostMemory(const char* arg);
};
}  // namespace register_kernel
}  // namespace tensorflow

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->CtxFailure((STATUS));    \
      return;                         \
    }                                 \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)                  \
  do {                                            \
    ::tensorflow::Status _s(__VA_ARGS__);         \
    if (!_s.ok()) {                               \
      (CTX)->CtxFailureWithWarning(_s);           \
      return;                                     \
    }                                             \
  } while (0)
#define EOT_STUB_CAT2(a, b) a##b
#define EOT_STUB_CAT(a, b) EOT_STUB_CAT2(a, b)
// the kernel class must derive from OpKernel and be constructible from an OpKernelConstruction*
#define REGISTER_KERNEL_BUILDER(BUILDER, ...)                                                              \
  static ::tensorflow::OpKernel* EOT_STUB_CAT(eot_stub_make_kernel_, __LINE__)(::tensorflow::OpKernelConstruction* c) { \
    (void)(::tensorflow::register_kernel::BUILDER);                                                        \
    return new __VA_ARGS__(c);                                                                             \
  }
#endif
