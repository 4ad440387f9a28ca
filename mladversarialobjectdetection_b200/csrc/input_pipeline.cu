// Input pipeline either side of the attack step
// (reference: train_data_generator.py:55-75 `DataSequence._map_fn`, :201-226 batch augmentation).
//
//   k_letterbox   decoded uint8 frames of any size -> [B,H,W,3] float32 in [-1,1]-ish: standardise in float64
//                 ((v - mean_c) / std_c, a 3 x 256 table per CTA), aspect-preserving half-pixel-centre bilinear resize
//                 in float64 (cv2.resize INTER_LINEAR on a CV_64F image: horizontal pass then vertical pass, border
//                 taps clamped), zero pad to the model size, one rounding to float32.  Also accumulates the per-image
//                 per-channel sums the contrast augmentation needs, so the batch is read once less.
//   k_channel_sums  the same sums for a batch that did not come from k_letterbox.
//   k_augment     random_flip_left_right ^ RandomFlip('horizontal'), RandomContrast ((x - mean_hw) * f + mean_hw),
//                 random_brightness (+ delta), clip to [-1,1]; float32 op by op as TF runs them (-fmad=false TU).
//                 The random draws are explicit inputs.
//
// HBM-bound: 12*H*W bytes written per frame by the letterbox (+ the uint8 frame read once), 24*H*W by the augmentation.
#include "eot_common.cuh"

#include <math.h>

namespace eot {

constexpr int kFramesPerLaunch = 64;   // frame descriptors travel as a kernel parameter (no device-side table, reentrant)

struct FrameBatch {
  const uint8_t* data[kFramesPerLaunch];
  int32_t h[kFramesPerLaunch], w[kFramesPerLaunch];       // decoded size
  int32_t sh[kFramesPerLaunch], sw[kFramesPerLaunch];     // scaled size inside the letterbox
  double scale_y[kFramesPerLaunch], scale_x[kFramesPerLaunch];   // 1 / (scaled / decoded), as cv2 derives it
  int32_t n, first;                                       // frames in this launch, index of the first one in the batch
};

struct Standardise { double mean[3], stddev[3]; };

__host__ __device__ inline double resize_scale(int dsize, int ssize) { return 1.0 / ((double)dsize / (double)ssize); }

__device__ __forceinline__ void bilinear_tap(int o, double scale, int ssize, int* s0, int* s1, double* w0, double* w1) {
  double f = ((double)o + 0.5) * scale - 0.5;
  int s = (int)floor(f);
  f = f - (double)s;
  if (s < 0) { s = 0; f = 0.0; }
  if (s >= ssize - 1) { s = ssize - 1; f = 0.0; }
  *s0 = s;
  *s1 = min(s + 1, ssize - 1);
  *w0 = 1.0 - f;
  *w1 = f;
}

__global__ void __launch_bounds__(kThreads) k_letterbox(FrameBatch fb, Standardise st, int H, int W, float* out,
                                                        double* channel_sums) {
  __shared__ double s_norm[3][256];
  __shared__ double s_red[32];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, v = i & 255;
    s_norm[c][v] = ((double)v - st.mean[c]) / st.stddev[c];         // image -= mean; image /= stddev (float64)
  }
  __syncthreads();
  const int k = blockIdx.y;
  const int b = fb.first + k;
  const uint8_t* src = fb.data[k];
  const int h = fb.h[k], w = fb.w[k], sh = fb.sh[k], sw = fb.sw[k];
  const bool same = (sh == h && sw == w);                            // cv2.resize to the same size copies
  float* o = out + (size_t)b * H * W * 3;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    const int oy = p / W, ox = p - oy * W;
    float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f;                           // np.zeros outside the scaled image
    if (oy < sh && ox < sw) {
      if (same) {
        const uint8_t* q = src + ((size_t)oy * w + ox) * 3;
        r0 = (float)s_norm[0][q[0]]; r1 = (float)s_norm[1][q[1]]; r2 = (float)s_norm[2][q[2]];
      } else {
        int x0, x1, y0, y1;
        double a0, a1, b0, b1;
        bilinear_tap(ox, fb.scale_x[k], w, &x0, &x1, &a0, &a1);
        bilinear_tap(oy, fb.scale_y[k], h, &y0, &y1, &b0, &b1);
        const uint8_t* q00 = src + ((size_t)y0 * w + x0) * 3;
        const uint8_t* q01 = src + ((size_t)y0 * w + x1) * 3;
        const uint8_t* q10 = src + ((size_t)y1 * w + x0) * 3;
        const uint8_t* q11 = src + ((size_t)y1 * w + x1) * 3;
        double v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double top = s_norm[c][q00[c]] * a0 + s_norm[c][q01[c]] * a1;    // horizontal pass, row y0
          const double bot = s_norm[c][q10[c]] * a0 + s_norm[c][q11[c]] * a1;    // horizontal pass, row y1
          v[c] = top * b0 + bot * b1;                                            // vertical pass
        }
        r0 = (float)v[0]; r1 = (float)v[1]; r2 = (float)v[2];
      }
    }
    o[(size_t)p * 3] = r0; o[(size_t)p * 3 + 1] = r1; o[(size_t)p * 3 + 2] = r2;
    acc0 += (double)r0; acc1 += (double)r1; acc2 += (double)r2;
  }
  if (channel_sums) {
    acc0 = block_sum(acc0, s_red);
    acc1 = block_sum(acc1, s_red);
    acc2 = block_sum(acc2, s_red);
    if (threadIdx.x == 0) {
      atomicAdd(channel_sums + (size_t)b * 3, acc0);
      atomicAdd(channel_sums + (size_t)b * 3 + 1, acc1);
      atomicAdd(channel_sums + (size_t)b * 3 + 2, acc2);
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_channel_sums(const float* __restrict__ in, int HW, double* channel_sums) {
  __shared__ double s_red[32];
  const int b = blockIdx.y;
  const float* x = in + (size_t)b * HW * 3;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    acc0 += (double)__ldg(x + (size_t)p * 3);
    acc1 += (double)__ldg(x + (size_t)p * 3 + 1);
    acc2 += (double)__ldg(x + (size_t)p * 3 + 2);
  }
  acc0 = block_sum(acc0, s_red);
  acc1 = block_sum(acc1, s_red);
  acc2 = block_sum(acc2, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(channel_sums + (size_t)b * 3, acc0);
    atomicAdd(channel_sums + (size_t)b * 3 + 1, acc1);
    atomicAdd(channel_sums + (size_t)b * 3 + 2, acc2);
  }
}

__device__ __forceinline__ float aug1(float v, float m, float contrast, float delta) {
  return clampf(((v - m) * contrast + m) + delta, -1.0f, 1.0f);
}

// kVec: W % 4 == 0 and 16-byte aligned batches -- a thread owns 4 consecutive output pixels (3 x 128-bit stores); their
// sources are 4 consecutive pixels as well (mirrored: the same quad read back to front).
template <bool kVec>
__global__ void __launch_bounds__(kThreads) k_augment(const float* __restrict__ in, float* out, int H, int W,
                                                      const uint8_t* __restrict__ flip,
                                                      const double* __restrict__ channel_sums, float contrast,
                                                      float delta) {
  const int b = blockIdx.y;
  const size_t img = (size_t)b * H * W * 3;
  const bool fl = flip && flip[b];
  const double n = (double)H * (double)W;
  const float m0 = (float)(channel_sums[(size_t)b * 3] / n), m1 = (float)(channel_sums[(size_t)b * 3 + 1] / n),
              m2 = (float)(channel_sums[(size_t)b * 3 + 2] / n);
  if (kVec) {
    const int wq = W >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < H * wq; q += gridDim.x * blockDim.x) {
      const int y = q / wq, xq = q - y * wq;
      const int sq = fl ? wq - 1 - xq : xq;
      const float4* src = reinterpret_cast<const float4*>(in + img + ((size_t)y * W + 4 * sq) * 3);
      const float4 A = __ldg(src), Bv = __ldg(src + 1), C = __ldg(src + 2);
      float p[12] = {A.x, A.y, A.z, A.w, Bv.x, Bv.y, Bv.z, Bv.w, C.x, C.y, C.z, C.w};
      float r[12];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sp = fl ? 3 - i : i;
        r[3 * i] = aug1(p[3 * sp], m0, contrast, delta);
        r[3 * i + 1] = aug1(p[3 * sp + 1], m1, contrast, delta);
        r[3 * i + 2] = aug1(p[3 * sp + 2], m2, contrast, delta);
      }
      float4* dst = reinterpret_cast<float4*>(out + img + ((size_t)y * W + 4 * xq) * 3);
      dst[0] = make_float4(r[0], r[1], r[2], r[3]);
      dst[1] = make_float4(r[4], r[5], r[6], r[7]);
      dst[2] = make_float4(r[8], r[9], r[10], r[11]);
    }
  } else {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
      const int y = p / W, x = p - y * W;
      const float* q = in + img + ((size_t)y * W + (fl ? W - 1 - x : x)) * 3;
      float* o = out + img + (size_t)p * 3;
      o[0] = aug1(__ldg(q), m0, contrast, delta);
      o[1] = aug1(__ldg(q + 1), m1, contrast, delta);
      o[2] = aug1(__ldg(q + 2), m2, contrast, delta);
    }
  }
}

}  // namespace eot

using namespace eot;

extern "C" int eot_letterbox_normalize(const uint8_t* const* frames, const int32_t* heights, const int32_t* widths,
                                       int32_t batch, int32_t out_height, int32_t out_width, const double* mean_rgb,
                                       const double* stddev_rgb, float* out, double* channel_sums, void* stream) {
  if (!frames || !heights || !widths || !mean_rgb || !stddev_rgb || !out) { set_error("eot_letterbox_normalize: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (batch <= 0 || out_height <= 0 || out_width <= 0) { set_error("eot_letterbox_normalize: bad shape B=%d H=%d W=%d", batch, out_height, out_width); return EOT_ERR_BAD_SHAPE; }
  cudaStream_t st = (cudaStream_t)stream;
  Standardise sd;
  for (int c = 0; c < 3; ++c) {
    sd.mean[c] = mean_rgb[c];
    sd.stddev[c] = stddev_rgb[c];
    if (!(stddev_rgb[c] != 0.0)) { set_error("eot_letterbox_normalize: stddev_rgb[%d] is 0", c); return EOT_ERR_BAD_SHAPE; }
  }
  if (channel_sums) EOT_CHECK_CUDA(cudaMemsetAsync(channel_sums, 0, (size_t)batch * 3 * sizeof(double), st));
  const int nsm = sm_count();
  for (int first = 0; first < batch; first += kFramesPerLaunch) {
    FrameBatch fb;
    fb.n = batch - first < kFramesPerLaunch ? batch - first : kFramesPerLaunch;
    fb.first = first;
    for (int k = 0; k < fb.n; ++k) {
      const int h = heights[first + k], w = widths[first + k];
      if (!frames[first + k] || h <= 0 || w <= 0) { set_error("eot_letterbox_normalize: frame %d is empty", first + k); return EOT_ERR_BAD_SHAPE; }
      // train_data_generator.py:66-70: Python float arithmetic, int() truncation
      const double sy = (double)out_height / (double)h, sx = (double)out_width / (double)w;
      const double image_scale = sx < sy ? sx : sy;
      const int sh = (int)((double)h * image_scale), sw = (int)((double)w * image_scale);
      if (sh <= 0 || sw <= 0 || sh > out_height || sw > out_width) {
        set_error("eot_letterbox_normalize: frame %d (%dx%d) scales to %dx%d (cv2.resize would fail)", first + k, h, w, sh, sw);
        return EOT_ERR_BAD_SHAPE;
      }
      fb.data[k] = frames[first + k];
      fb.h[k] = h; fb.w[k] = w; fb.sh[k] = sh; fb.sw[k] = sw;
      fb.scale_y[k] = resize_scale(sh, h); fb.scale_x[k] = resize_scale(sw, w);
    }
    const int per_frame = (out_height * out_width + kThreads - 1) / kThreads;
    int gx = (nsm * 8 + fb.n - 1) / fb.n;
    if (gx > per_frame) gx = per_frame;
    if (gx < 1) gx = 1;
    k_letterbox<<<dim3(gx, fb.n), kThreads, 0, st>>>(fb, sd, out_height, out_width, out, channel_sums);
    count_launches(1);
  }
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int eot_channel_sums(const float* images, int32_t batch, int32_t height, int32_t width, double* channel_sums,
                                void* stream) {
  if (!images || !channel_sums) { set_error("eot_channel_sums: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (batch <= 0 || height <= 0 || width <= 0) { set_error("eot_channel_sums: bad shape"); return EOT_ERR_BAD_SHAPE; }
  cudaStream_t st = (cudaStream_t)stream;
  EOT_CHECK_CUDA(cudaMemsetAsync(channel_sums, 0, (size_t)batch * 3 * sizeof(double), st));
  const int per = (height * width + kThreads - 1) / kThreads;
  int gx = (sm_count() * 8 + batch - 1) / batch;
  gx = gx > per ? per : (gx < 1 ? 1 : gx);
  k_channel_sums<<<dim3(gx, batch), kThreads, 0, st>>>(images, height * width, channel_sums);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int eot_augment_batch(const float* images, float* out, int32_t batch, int32_t height, int32_t width,
                                 const uint8_t* flip, const double* channel_sums, float contrast_factor,
                                 float brightness_delta, void* stream) {
  if (!images || !out || !channel_sums) { set_error("eot_augment_batch: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (batch <= 0 || height <= 0 || width <= 0) { set_error("eot_augment_batch: bad shape"); return EOT_ERR_BAD_SHAPE; }
  if (images == out && flip) { set_error("eot_augment_batch: the flip cannot run in place"); return EOT_ERR_BAD_SHAPE; }
  const bool vec = (width % 4 == 0) && (((uintptr_t)images | (uintptr_t)out) & 15) == 0;
  const int per = (height * width / (vec ? 4 : 1) + kThreads - 1) / kThreads;
  int gx = (sm_count() * 8 + batch - 1) / batch;
  gx = gx > per ? per : (gx < 1 ? 1 : gx);
  if (vec)
    k_augment<true><<<dim3(gx, batch), kThreads, 0, (cudaStream_t)stream>>>(images, out, height, width, flip, channel_sums,
                                                                           contrast_factor, brightness_delta);
  else
    k_augment<false><<<dim3(gx, batch), kThreads, 0, (cudaStream_t)stream>>>(images, out, height, width, flip, channel_sums,
                                                                            contrast_factor, brightness_delta);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
