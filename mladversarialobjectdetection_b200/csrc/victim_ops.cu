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

// channels % 4 != 0 but even (the 810-channel class logits): the same pass with 64-bit accesses
template <int ACT>
__global__ void __launch_bounds__(kThreads) k_bias_act2(const float* __restrict__ x, const float* __restrict__ bias, float* y,
                                                        long long n2, int c2) {
  const float2* x2 = reinterpret_cast<const float2*>(x);
  const float2* b2 = reinterpret_cast<const float2*>(bias);
  float2* y2 = reinterpret_cast<float2*>(y);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    float2 v = x2[i];
    const float2 b = __ldg(b2 + (int)(i % c2));
    v.x += b.x; v.y += b.y;
    if (ACT == 1) { v.x = silu_f(v.x); v.y = silu_f(v.y); }
    y2[i] = v;
  }
}

// ---- squeeze-and-excitation gate (MBConv): out = y * gate[n,c]; backward dy = dout * gate[n,c] + dmean[n,c] / HW and
// dgate[n,c] = sum_hw dout * y (deterministic two-stage reduction) -- PyTorch runs these as broadcast multiplies /
// expands through its non-vectorised kernel.
__global__ void __launch_bounds__(kThreads) k_channel_scale(const float* __restrict__ y, const float* __restrict__ gate,
                                                            const float* __restrict__ shift, float shift_mul, float* out,
                                                            long long hw4 /* float4 per image */, int c4) {
  const int n = blockIdx.y;
  const float4* y4 = reinterpret_cast<const float4*>(y) + (size_t)n * hw4;
  const float4* g4 = reinterpret_cast<const float4*>(gate) + (size_t)n * c4;
  const float4* s4 = shift ? reinterpret_cast<const float4*>(shift) + (size_t)n * c4 : nullptr;
  float4* o4 = reinterpret_cast<float4*>(out) + (size_t)n * hw4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw4; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % c4);
    const float4 v = y4[i], g = __ldg(g4 + q);
    float4 r = make_float4(v.x * g.x, v.y * g.y, v.z * g.z, v.w * g.w);
    if (s4) { const float4 s = __ldg(s4 + q); r.x += s.x * shift_mul; r.y += s.y * shift_mul; r.z += s.z * shift_mul; r.w += s.w * shift_mul; }
    o4[i] = r;
  }
}

constexpr int kDotChunks = 16;
// partial[chunk][n][c] = sum over the chunk's pixels of a * b; threads = (pixel lane, channel quad), channel-fastest
__global__ void __launch_bounds__(kThreads) k_channel_dot_partial(const float* __restrict__ a, const float* __restrict__ b,
                                                                  float* partial, int hw, int c4, int n_images) {
  __shared__ float4 red[kThreads];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int ql_n = c4 < kThreads ? c4 : kThreads;              // quad lanes
  const int pl_n = kThreads / ql_n;                              // pixel lanes
  const int ql = threadIdx.x % ql_n, pl = threadIdx.x / ql_n;
  const int per = (hw + kDotChunks - 1) / kDotChunks;
  const int p0 = chunk * per, p1 = min(hw, p0 + per);
  const float4* a4 = reinterpret_cast<const float4*>(a) + (size_t)n * hw * c4;
  const float4* b4 = reinterpret_cast<const float4*>(b) + (size_t)n * hw * c4;
  for (int q0 = 0; q0 < c4; q0 += ql_n) {
    const int q = q0 + ql;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < c4 && pl < pl_n) {
      for (int p = p0 + pl; p < p1; p += pl_n) {
        const float4 u = a4[(size_t)p * c4 + q], v = b4[(size_t)p * c4 + q];
        acc.x += u.x * v.x; acc.y += u.y * v.y; acc.z += u.z * v.z; acc.w += u.w * v.w;
      }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (pl == 0 && q < c4) {
      float4 s = red[ql];
      for (int k = 1; k < pl_n; ++k) { const float4 t = red[k * ql_n + ql]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
      reinterpret_cast<float4*>(partial)[((size_t)chunk * n_images + n) * c4 + q] = s;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kThreads) k_channel_dot_final(const float* __restrict__ partial, float* out, int nc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  float s = 0.0f;
  for (int k = 0; k < kDotChunks; ++k) s += partial[(size_t)k * nc + i];
  out[i] = s;
}

// ---- BiFPN fast normalised fusion (Fuse): out = silu(sum_i w_i x_i), n <= 3 same-shaped inputs; backward
// dx_i = w_i * dout * silu'(z).  PyTorch: n broadcast multiplies + n-1 adds + silu (and as many again backward).
struct FuseArgs { const float* x[3]; float* dx[3]; };
template <int N>
__global__ void __launch_bounds__(kThreads) k_fuse_silu(FuseArgs a, const float* __restrict__ w, float* out, long long n4) {
  float wv[N];
#pragma unroll
  for (int k = 0; k < N; ++k) wv[k] = __ldg(w + k);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float4 v = reinterpret_cast<const float4*>(a.x[k])[i];
      if (k == 0) z = make_float4(v.x * wv[0], v.y * wv[0], v.z * wv[0], v.w * wv[0]);
      else { z.x += v.x * wv[k]; z.y += v.y * wv[k]; z.z += v.z * wv[k]; z.w += v.w * wv[k]; }
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(silu_f(z.x), silu_f(z.y), silu_f(z.z), silu_f(z.w));
  }
}
template <int N>
__global__ void __launch_bounds__(kThreads) k_fuse_silu_bwd(FuseArgs a, const float* __restrict__ w, const float* __restrict__ dout,
                                                            long long n4) {
  float wv[N];
#pragma unroll
  for (int k = 0; k < N; ++k) wv[k] = __ldg(w + k);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float4 v = reinterpret_cast<const float4*>(a.x[k])[i];
      if (k == 0) z = make_float4(v.x * wv[0], v.y * wv[0], v.z * wv[0], v.w * wv[0]);
      else { z.x += v.x * wv[k]; z.y += v.y * wv[k]; z.z += v.z * wv[k]; z.w += v.w * wv[k]; }
    }
    const float4 g = reinterpret_cast<const float4*>(dout)[i];
    const float4 dz = make_float4(g.x * dsilu_f(z.x), g.y * dsilu_f(z.y), g.z * dsilu_f(z.z), g.w * dsilu_f(z.w));
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (a.dx[k]) reinterpret_cast<float4*>(a.dx[k])[i] = make_float4(dz.x * wv[k], dz.y * wv[k], dz.z * wv[k], dz.w * wv[k]);
  }
}

static int epilogue_grid(long long n4) {
  long long g = (n4 + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int check_epilogue(const void* a, const void* b, const void* c, long long n_pixels, int channels, const char* who,
                          int mult = 4) {
  if (!a || !b || !c) { set_error("%s: NULL pointer", who); return EOT_ERR_NULL_POINTER; }
  if (n_pixels <= 0 || channels <= 0 || channels % mult != 0) { set_error("%s: needs channels %% 4 == 0 (got %d) and pixels > 0", who, channels); return EOT_ERR_BAD_SHAPE; }
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & (mult * 4 - 1)) != 0) { set_error("%s: pointers must be %d-byte aligned", who, mult * 4); return EOT_ERR_MISALIGNED; }
  return EOT_OK;
}

}  // namespace eot

using namespace eot;

extern "C" int nhwc_bias_act_fwd(const float* x, const float* bias, float* y, int64_t n_pixels, int32_t channels, int32_t act,
                                 void* stream) {
  if (act != 0 && act != 1) { set_error("nhwc_bias_act_fwd: act must be 0 (identity) or 1 (SiLU)"); return EOT_ERR_BAD_SHAPE; }
  if (channels % 4 != 0) {
    if (int rc = check_epilogue(x, bias, y, n_pixels, channels, "nhwc_bias_act_fwd", 2)) return rc;
    const long long n2 = n_pixels * (long long)(channels / 2);
    if (act == 1) k_bias_act2<1><<<epilogue_grid(n2), kThreads, 0, (cudaStream_t)stream>>>(x, bias, y, n2, channels / 2);
    else k_bias_act2<0><<<epilogue_grid(n2), kThreads, 0, (cudaStream_t)stream>>>(x, bias, y, n2, channels / 2);
    count_launches(1);
    EOT_CHECK_CUDA(cudaPeekAtLastError());
    return EOT_OK;
  }
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

extern "C" int nhwc_channel_scale(const float* y, const float* gate, const float* shift, float shift_mul, float* out,
                                  int32_t n_images, int64_t hw, int32_t channels, void* stream) {
  if (int rc = check_epilogue(y, gate, out, hw, channels, "nhwc_channel_scale")) return rc;
  if (n_images <= 0) { set_error("nhwc_channel_scale: empty batch"); return EOT_ERR_BAD_SHAPE; }
  const long long hw4 = hw * (long long)(channels / 4);
  int gx = (int)((hw4 + kThreads - 1) / kThreads);
  const int cap = (sm_count() * 16 + n_images - 1) / n_images;
  gx = gx > cap ? cap : gx;
  k_channel_scale<<<dim3(gx < 1 ? 1 : gx, n_images), kThreads, 0, (cudaStream_t)stream>>>(y, gate, shift, shift_mul, out, hw4,
                                                                                    channels / 4);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int nhwc_channel_dot(const float* a, const float* b, float* out, float* workspace /* 16*N*C floats */,
                                int32_t n_images, int32_t hw, int32_t channels, void* stream) {
  if (int rc = check_epilogue(a, b, out, hw, channels, "nhwc_channel_dot")) return rc;
  if (!workspace || ((uintptr_t)workspace & 15) || n_images <= 0) { set_error("nhwc_channel_dot: bad workspace / batch"); return EOT_ERR_NULL_POINTER; }
  cudaStream_t st = (cudaStream_t)stream;
  k_channel_dot_partial<<<dim3(kDotChunks, n_images), kThreads, 0, st>>>(a, b, workspace, hw, channels / 4, n_images);
  const int nc = n_images * channels;
  k_channel_dot_final<<<(nc + kThreads - 1) / kThreads, kThreads, 0, st>>>(workspace, out, nc);
  count_launches(2);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int nhwc_fuse_silu_fwd(const float* const* xs, int32_t n, const float* weights, float* out, int64_t n_elems, void* stream) {
  if (!xs || !weights || !out) { set_error("nhwc_fuse_silu_fwd: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (n < 2 || n > 3 || n_elems <= 0 || n_elems % 4 != 0) { set_error("nhwc_fuse_silu_fwd: needs 2 or 3 inputs and a multiple of 4 elements"); return EOT_ERR_BAD_SHAPE; }
  FuseArgs a = {};
  uintptr_t al = (uintptr_t)out;
  for (int k = 0; k < n; ++k) { if (!xs[k]) { set_error("nhwc_fuse_silu_fwd: input %d is NULL", k); return EOT_ERR_NULL_POINTER; } a.x[k] = xs[k]; al |= (uintptr_t)xs[k]; }
  if (al & 15) { set_error("nhwc_fuse_silu_fwd: pointers must be 16-byte aligned"); return EOT_ERR_MISALIGNED; }
  const long long n4 = n_elems / 4;
  if (n == 2) k_fuse_silu<2><<<epilogue_grid(n4), kThreads, 0, (cudaStream_t)stream>>>(a, weights, out, n4);
  else k_fuse_silu<3><<<epilogue_grid(n4), kThreads, 0, (cudaStream_t)stream>>>(a, weights, out, n4);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int nhwc_fuse_silu_bwd(const float* const* xs, int32_t n, const float* weights, const float* dout, float* const* dxs,
                                  int64_t n_elems, void* stream) {
  if (!xs || !weights || !dout || !dxs) { set_error("nhwc_fuse_silu_bwd: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (n < 2 || n > 3 || n_elems <= 0 || n_elems % 4 != 0) { set_error("nhwc_fuse_silu_bwd: needs 2 or 3 inputs and a multiple of 4 elements"); return EOT_ERR_BAD_SHAPE; }
  FuseArgs a = {};
  uintptr_t al = (uintptr_t)dout;
  for (int k = 0; k < n; ++k) {
    if (!xs[k]) { set_error("nhwc_fuse_silu_bwd: input %d is NULL", k); return EOT_ERR_NULL_POINTER; }
    a.x[k] = xs[k]; a.dx[k] = dxs[k];
    al |= (uintptr_t)xs[k] | (uintptr_t)dxs[k];
  }
  if (al & 15) { set_error("nhwc_fuse_silu_bwd: pointers must be 16-byte aligned"); return EOT_ERR_MISALIGNED; }
  const long long n4 = n_elems / 4;
  if (n == 2) k_fuse_silu_bwd<2><<<epilogue_grid(n4), kThreads, 0, (cudaStream_t)stream>>>(a, weights, dout, n4);
  else k_fuse_silu_bwd<3><<<epilogue_grid(n4), kThreads, 0, (cudaStream_t)stream>>>(a, weights, dout, n4);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
