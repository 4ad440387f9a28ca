// Patch update pieces of PatchAttacker.train_step (attacker.py:191-193, 307-316, 51-54):
//   patch_tv_grad     1e-5 * tf.image.total_variation(patch) and its (sign) gradient
//   adam_clip_update  Keras Adam (ResourceApplyAdam formulas, tensorflow/core/kernels/training_ops.cc)
//                     followed by the tf.Variable clip constraint, in one pass
#include "eot_common.cuh"

#include <math.h>

namespace eot {

__device__ __forceinline__ float sgn(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

__global__ void __launch_bounds__(kThreads) k_tv_grad(const float* __restrict__ x, int P, float weight,
                                                      float* grad, float* tv_out) {
  __shared__ double red[32];
  const int n = P * P * 3;
  const int rs = P * 3;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / rs, rem = i - y * rs;
    const int xx = rem / 3;
    const float v = x[i];
    float g = 0.0f;
    if (y > 0) g += sgn(v - x[i - rs]);
    if (y < P - 1) { const float d = x[i + rs] - v; g -= sgn(d); acc += (double)fabsf(d); }
    if (xx > 0) g += sgn(v - x[i - 3]);
    if (xx < P - 1) { const float d = x[i + 3] - v; g -= sgn(d); acc += (double)fabsf(d); }
    if (grad) grad[i] = grad[i] + weight * g;
  }
  acc = block_sum(acc, red);
  if (tv_out && threadIdx.x == 0) atomicAdd(tv_out, (float)acc);
}

__global__ void __launch_bounds__(kThreads) k_adam_clip(float* var, float* m, float* v, const float* __restrict__ grad,
                                                        int64_t n, float alpha, float one_minus_b1, float one_minus_b2,
                                                        float eps, float lo, float hi) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = grad[i];
    const float mi = m[i] + (g - m[i]) * one_minus_b1;
    const float vi = v[i] + (g * g - v[i]) * one_minus_b2;
    m[i] = mi;
    v[i] = vi;
    const float w = var[i] - (mi * alpha) / (sqrtf(vi) + eps);
    var[i] = fminf(fmaxf(w, lo), hi);
  }
}

// tail of the packed all-reduce buffer: [dL/dscale, data loss, sum_b M_b, sum_b M_b^2] (attacker.py:190-201 needs the last two
// for the mean / std metrics of the GLOBAL batch)
__global__ void __launch_bounds__(kThreads) k_pack_scalars(const float* __restrict__ M, int batch, const float* __restrict__ dscale,
                                                           const float* __restrict__ data_loss, float* out4) {
  __shared__ double red[32];
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < batch; i += blockDim.x) { const double m = (double)M[i]; s1 += m; s2 += m * m; }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) { out4[0] = *dscale; out4[1] = *data_loss; out4[2] = (float)s1; out4[3] = (float)s2; }
}

// add_metric values of attacker.py:196-201 from the (all-reduced) tail: loss, scale_loss, mean / std of the max scores
__global__ void k_step_metrics(const float* __restrict__ tail4, const float* __restrict__ tv, const float* __restrict__ scale,
                               float global_batch, float tv_weight, float* out6) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float s1 = tail4[2], s2 = tail4[3], sc = *scale, B = global_batch;
  const float mean = s1 / B;
  out6[0] = tail4[1] + tv_weight * *tv;                           // loss = sum(M^2 + (M - scale)^2) + 1e-5 TV
  out6[1] = s2 - 2.0f * sc * s1 + B * sc * sc;                     // scale_loss = sum (M - scale)^2
  out6[2] = mean;
  out6[3] = sqrtf(fmaxf(s2 / B - mean * mean, 0.0f));
  out6[4] = *tv;
  out6[5] = sc;
}

}  // namespace eot

using namespace eot;

extern "C" int attack_pack_scalars(const float* max_scores, int32_t batch, const float* dscale, const float* data_loss,
                                   float* out4, void* stream) {
  if (!max_scores || !dscale || !data_loss || !out4) { set_error("attack_pack_scalars: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (batch <= 0) { set_error("attack_pack_scalars: bad batch %d", batch); return EOT_ERR_BAD_SHAPE; }
  k_pack_scalars<<<1, kThreads, 0, (cudaStream_t)stream>>>(max_scores, batch, dscale, data_loss, out4);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int attack_step_metrics(const float* tail4, const float* tv, const float* scale, float global_batch, float tv_weight,
                                   float* out6, void* stream) {
  if (!tail4 || !tv || !scale || !out6) { set_error("attack_step_metrics: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (!(global_batch > 0.0f)) { set_error("attack_step_metrics: bad batch"); return EOT_ERR_BAD_SHAPE; }
  k_step_metrics<<<1, 32, 0, (cudaStream_t)stream>>>(tail4, tv, scale, global_batch, tv_weight, out6);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int patch_tv_grad(const float* patch, int32_t patch_size, float weight, float* grad_patch, float* tv_out,
                             void* stream) {
  if (!patch) { set_error("patch_tv_grad: patch is NULL"); return EOT_ERR_NULL_POINTER; }
  if (patch_size <= 0) { set_error("patch_tv_grad: bad patch size %d", patch_size); return EOT_ERR_BAD_SHAPE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (tv_out) EOT_CHECK_CUDA(cudaMemsetAsync(tv_out, 0, sizeof(float), st));
  const int n = patch_size * patch_size * 3;
  const int blocks = min((n + kThreads - 1) / kThreads, sm_count() * 8);
  k_tv_grad<<<blocks, kThreads, 0, st>>>(patch, patch_size, weight, grad_patch, tv_out);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int adam_clip_update(float* var, float* m, float* v, const float* grad, int64_t n, float lr, float beta1,
                                float beta2, float eps, int64_t step, float lo, float hi, void* stream) {
  if (!var || !m || !v || !grad) { set_error("adam_clip_update: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (n <= 0 || step <= 0) { set_error("adam_clip_update: bad n/step"); return EOT_ERR_BAD_SHAPE; }
  const double b1p = pow((double)beta1, (double)step), b2p = pow((double)beta2, (double)step);
  const float alpha = (float)((double)lr * sqrt(1.0 - b2p) / (1.0 - b1p));
  int64_t nb = (n + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (nb > cap) nb = cap;
  const int blocks = (int)nb;
  k_adam_clip<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(var, m, v, grad, n, alpha, 1.0f - beta1, 1.0f - beta2, eps, lo, hi);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
