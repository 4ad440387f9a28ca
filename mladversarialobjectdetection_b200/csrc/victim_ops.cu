// Epilogue helpers for the torch stand-in of the EfficientDet victim (victim.py) -- NOT part of the reference's hot
// path: the convolutions stay on the framework's own GPU path (cuDNN), as the north star asks.  PyTorch applies a
// convolution's bias as a separate broadcast add (a non-vectorised kernel in channels_last) and SiLU as another pass;
// here the per-channel bias and the activation are one 128-bit pass over the NHWC activation, forward and backward.
//   k_bias_act      y = act(x + bias[c])            act: 0 = identity, 1 = SiLU (x * sigmoid(x))
//   k_bias_silu_bwd dx = dy * d/dz silu(z), z = x + bias[c]   (x = the convolution output saved by the forward)
#include "eot_common.cuh"

#include <math.h>

namespace eot {

__device__ __forceinline__ float silu_f(float z) { return z / (1.0f + expf(-z)); }
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = 1.0f / (1.0f + expf(-z));
  return s * (1.0f + z * (1.0f - s));
}

template <int ACT>
__global__ void __launch_bounds__(kThreads) k_bias_act(const float* __restrict__ x, const float* __restrict__ bias, float* y,
                                                       long long n4, int c4) {
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* b4 = reinterpret_cast<const float4*>(bias);
  float4* y4 = reinterpret_cast<float4*>(y);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    const float4 b = __ldg(b4 + (int)(i % c4));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if (ACT == 1) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
    y4[i] = v;
  }
}

__global__ void __launch_bounds__(kThreads) k_bias_silu_bwd(const float* __restrict__ x, const float* __restrict__ bias,
                                                            const float* __restrict__ dy, float* dx, long long n4, int c4) {
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* b4 = reinterpret_cast<const float4*>(bias);
  const float4* g4 = reinterpret_cast<const float4*>(dy);
  float4* o4 = reinterpret_cast<float4*>(dx);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x4[i], g = g4[i];
    const float4 b = __ldg(b4 + (int)(i % c4));
    o4[i] = make_float4(g.x * dsilu_f(v.x + b.x), g.y * dsilu_f(v.y + b.y), g.z * dsilu_f(v.z + b.z), g.w * dsilu_f(v.w + b.w));
  }
}

static int epilogue_grid(long long n4) {
  long long g = (n4 + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int check_epilogue(const void* a, const void* b, const void* c, long long n_pixels, int channels, const char* who) {
  if (!a || !b || !c) { set_error("%s: NULL pointer", who); return EOT_ERR_NULL_POINTER; }
  if (n_pixels <= 0 || channels <= 0 || channels % 4 != 0) { set_error("%s: needs channels %% 4 == 0 (got %d) and pixels > 0", who, channels); return EOT_ERR_BAD_SHAPE; }
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) != 0) { set_error("%s: pointers must be 16-byte aligned", who); return EOT_ERR_MISALIGNED; }
  return EOT_OK;
}

}  // namespace eot

using namespace eot;

extern "C" int nhwc_bias_act_fwd(const float* x, const float* bias, float* y, int64_t n_pixels, int32_t channels, int32_t act,
                                 void* stream) {
  if (int rc = check_epilogue(x, bias, y, n_pixels, channels, "nhwc_bias_act_fwd")) return rc;
  const long long n4 = n_pixels * (long long)(channels / 4);
  if (act == 1) k_bias_act<1><<<epilogue_grid(n4), kThreads, 0, (cudaStream_t)stream>>>(x, bias, y, n4, channels / 4);
  else if (act == 0) k_bias_act<0><<<epilogue_grid(n4), kThreads, 0, (cudaStream_t)stream>>>(x, bias, y, n4, channels / 4);
  else { set_error("nhwc_bias_act_fwd: act must be 0 (identity) or 1 (SiLU)"); return EOT_ERR_BAD_SHAPE; }
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int nhwc_bias_silu_bwd(const float* x, const float* bias, const float* dy, float* dx, int64_t n_pixels,
                                  int32_t channels, void* stream) {
  if (int rc = check_epilogue(x, bias, dy, n_pixels, channels, "nhwc_bias_silu_bwd")) return rc;
  if (!dx || ((uintptr_t)dx & 15)) { set_error("nhwc_bias_silu_bwd: dx NULL or misaligned"); return EOT_ERR_NULL_POINTER; }
  const long long n4 = n_pixels * (long long)(channels / 4);
  k_bias_silu_bwd<<<epilogue_grid(n4), kThreads, 0, (cudaStream_t)stream>>>(x, bias, dy, dx, n4, channels / 4);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
