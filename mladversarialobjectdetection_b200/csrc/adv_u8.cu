// GPU twin of the uint8 inference-time patcher `adv_patch.AdversarialPatch`
// (reference: adv_patch.py:40-201 -- print_patch, _create, rescale, brightness_match, resize, get_transformed_patch,
//  add_adv_to_img).  The reference runs ~10 OpenCV calls per person box on the CPU; here a box is three small launches
// on the caller's stream, the frame stays on the device and is patched in place:
//
//   k_adv_frame_ysum   sum of Y over `rescale(frame)` : letter-box resize of the frame (cv2.resize INTER_LINEAR on 8-bit
//                      data: 11-bit fixed point, horizontal then vertical; 2x2 box average for an exact 2x; copy for 1x)
//                      -> 8-bit fixed-point RGB->Y (shift 14), integer sum.  Padding (127 grey) is added in closed form.
//   k_adv_match        brightness match of the printed patch: Y' = uint8(clip(Y - mean_src + mean_tgt, 0, 255)) in
//                      float64, YUV -> RGB in 8-bit fixed point.  Every paste sees the earlier pastes (sequential).
//   k_adv_area_paste   cv2.resize INTER_AREA of the matched patch to the box's patch size (integer ratios: box sums;
//                      otherwise computeResizeAreaTab's float32 tap weights, accumulated in OpenCV's order), the
//                      float64 noise / clip / re-quantise chain of get_transformed_patch, and the slice assignment.
//
// Integer and float32 arithmetic is evaluated exactly as OpenCV's scalar code does (this TU is built with -fmad=false);
// results are bit-identical to the reference run with opencv-python 4.13 (tests/golden/adv_patch_u8.npz).
// The INTER_CUBIC branch (patch up-sampling, adv_patch.py:158-160) follows OpenCV's own 8-bit bicubic kernel bit for
// bit; pip wheels route that call through Intel IPP, whose closed-source evaluation differs by one grey level on ~4 % of
// the elements (see oracle/adv_patch_u8.py).
#include "eot_common.cuh"

#include <math.h>

namespace eot {

__device__ __forceinline__ int descale14(int x) { return (x + (1 << 13)) >> 14; }
__device__ __forceinline__ int sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
__device__ __forceinline__ int y_of(int r, int g, int b) { return sat_u8(descale14(r * 4899 + g * 9617 + b * 1868)); }

// horizontal tap of the 8-bit INTER_LINEAR path: source index and the two 11-bit weights
__device__ __forceinline__ void lin_tap_x(int d, double scale, int ssize, int* s0, int* s1, int* w0, int* w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = f - (float)s;
  if (s < 0) { s = 0; f = 0.0f; }
  if (s >= ssize - 1) { s = ssize - 1; f = 0.0f; }
  *s0 = s;
  *s1 = min(s + 1, ssize - 1);
  *w0 = __float2int_rn((1.0f - f) * 2048.0f);
  *w1 = __float2int_rn(f * 2048.0f);
}
// vertical tap: the weight keeps its fraction at the border, the ROWS are clamped
__device__ __forceinline__ void lin_tap_y(int d, double scale, int ssize, int* s0, int* s1, int* w0, int* w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  const int s = (int)floorf(f);
  f = f - (float)s;
  *s0 = min(max(s, 0), ssize - 1);
  *s1 = min(max(s + 1, 0), ssize - 1);
  *w0 = __float2int_rn((1.0f - f) * 2048.0f);
  *w1 = __float2int_rn(f * 2048.0f);
}

// mode 0: copy (same size), 1: exact 2x2 box average, 2: fixed-point bilinear
__global__ void __launch_bounds__(kThreads) k_adv_frame_ysum(const uint8_t* __restrict__ frame, int h, int w, int sh, int sw,
                                                             int mode, double scale_y, double scale_x,
                                                             unsigned long long* ysum) {
  __shared__ double red[32];
  double acc = 0.0;                                                   // exact: integers far below 2^53
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < sh * sw; p += gridDim.x * blockDim.x) {
    const int oy = p / sw, ox = p - oy * sw;
    int rgb[3];
    if (mode == 0) {
      const uint8_t* q = frame + ((size_t)oy * w + ox) * 3;
      rgb[0] = q[0]; rgb[1] = q[1]; rgb[2] = q[2];
    } else if (mode == 1) {
      const uint8_t* q = frame + ((size_t)(2 * oy) * w + 2 * ox) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) rgb[c] = ((int)q[c] + q[3 + c] + q[(size_t)w * 3 + c] + q[(size_t)w * 3 + 3 + c] + 2) >> 2;
    } else {
      int x0, x1, a0, a1, y0, y1, b0, b1;
      lin_tap_x(ox, scale_x, w, &x0, &x1, &a0, &a1);
      lin_tap_y(oy, scale_y, h, &y0, &y1, &b0, &b1);
      const uint8_t* q00 = frame + ((size_t)y0 * w + x0) * 3;
      const uint8_t* q01 = frame + ((size_t)y0 * w + x1) * 3;
      const uint8_t* q10 = frame + ((size_t)y1 * w + x0) * 3;
      const uint8_t* q11 = frame + ((size_t)y1 * w + x1) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = q00[c] * a0 + q01[c] * a1, h1 = q10[c] * a0 + q11[c] * a1;
        rgb[c] = sat_u8((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
      }
    }
    acc += (double)y_of(rgb[0], rgb[1], rgb[2]);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(ysum, (unsigned long long)acc);
}

__global__ void __launch_bounds__(kThreads) k_adv_patch_ysum(const uint8_t* __restrict__ patch, int n_px, unsigned long long* ysum) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += gridDim.x * blockDim.x) {
    const uint8_t* q = patch + (size_t)p * 3;
    acc += (double)y_of(q[0], q[1], q[2]);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(ysum, (unsigned long long)acc);
}

// sums[0] = sum Y(patch), sums[1] = sum Y(scaled frame); the 127-grey padding of rescale() has Y = 127 exactly
__global__ void __launch_bounds__(kThreads) k_adv_match(const uint8_t* __restrict__ patch, int n_px,
                                                        const unsigned long long* __restrict__ sums, long long pad_px,
                                                        long long tgt_px, uint8_t* matched) {
  const double source_mean = (double)sums[0] / (double)n_px;
  const double target_mean = (double)(sums[1] + 127ull * (unsigned long long)pad_px) / (double)tgt_px;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += gridDim.x * blockDim.x) {
    const uint8_t* q = patch + (size_t)p * 3;
    const int r = q[0], g = q[1], b = q[2];
    const int y = sat_u8(descale14(r * 4899 + g * 9617 + b * 1868));
    const int v = sat_u8(descale14((r - y) * 14369 + (128 << 14)));
    const int u = sat_u8(descale14((b - y) * 8061 + (128 << 14)));
    double res = ((double)y - source_mean) + target_mean;
    res = fmin(fmax(res, 0.0), 255.0);
    const int y2 = (int)res;                                          // astype('uint8'): truncation
    const int bb = y2 + descale14((u - 128) * 33292);
    const int gg = y2 + descale14((u - 128) * (-6472) + (v - 128) * (-9519));
    const int rr = y2 + descale14((v - 128) * 18678);
    uint8_t* o = matched + (size_t)p * 3;
    o[0] = (uint8_t)sat_u8(rr); o[1] = (uint8_t)sat_u8(gg); o[2] = (uint8_t)sat_u8(bb);
  }
}

// computeResizeAreaTab for one destination index: source range and the weights of its partial first / last cells
struct AreaTaps { int first, n; float a_first, a_mid, a_last; bool has_first, has_last; int s_mid0, s_mid1, s_last; };
__device__ __forceinline__ AreaTaps area_taps(int d, double scale, int ssize) {
  AreaTaps t;
  const double fsx1 = (double)d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  t.has_first = ((double)sx1 - fsx1) > 1e-3;
  t.a_first = (float)(((double)sx1 - fsx1) / cell);
  t.first = sx1 - 1;
  t.s_mid0 = sx1; t.s_mid1 = sx2;
  t.a_mid = (float)(1.0 / cell);
  t.has_last = (fsx2 - (double)sx2) > 1e-3;
  t.a_last = (float)(fmin(fmin(fsx2 - (double)sx2, 1.0), cell) / cell);
  t.s_last = sx2;
  t.n = 0;
  return t;
}

__device__ __forceinline__ void hsum3(const uint8_t* __restrict__ row, const AreaTaps& tx, float buf[3]) {
  buf[0] = buf[1] = buf[2] = 0.0f;
  if (tx.has_first) {
    const uint8_t* q = row + (size_t)tx.first * 3;
    buf[0] = buf[0] + (float)q[0] * tx.a_first; buf[1] = buf[1] + (float)q[1] * tx.a_first; buf[2] = buf[2] + (float)q[2] * tx.a_first;
  }
  for (int sx = tx.s_mid0; sx < tx.s_mid1; ++sx) {
    const uint8_t* q = row + (size_t)sx * 3;
    buf[0] = buf[0] + (float)q[0] * tx.a_mid; buf[1] = buf[1] + (float)q[1] * tx.a_mid; buf[2] = buf[2] + (float)q[2] * tx.a_mid;
  }
  if (tx.has_last) {
    const uint8_t* q = row + (size_t)tx.s_last * 3;
    buf[0] = buf[0] + (float)q[0] * tx.a_last; buf[1] = buf[1] + (float)q[1] * tx.a_last; buf[2] = buf[2] + (float)q[2] * tx.a_last;
  }
}

// cv2.resize INTER_CUBIC on 8-bit data (resize.cpp, OpenCV's own kernel): source offset and 11-bit taps of a
// destination index (interpolateCubic, A = -0.75, float32; saturate_cast<short> = round half to even).
struct CubicTaps { int s; int a[4]; };
__device__ __forceinline__ CubicTaps cubic_taps(int d, double scale) {
  CubicTaps t;
  const float f = (float)(((double)d + 0.5) * scale - 0.5);
  const float fl = floorf(f);
  t.s = (int)fl;
  const float x = f - fl, A = -0.75f;
  const float c0 = ((A * (x + 1.0f) - 5.0f * A) * (x + 1.0f) + 8.0f * A) * (x + 1.0f) - 4.0f * A;
  const float c1 = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  const float y = 1.0f - x;
  const float c2 = ((A + 2.0f) * y - (A + 3.0f)) * y * y + 1.0f;
  const float c3 = 1.0f - c0 - c1 - c2;
  t.a[0] = __float2int_rn(c0 * 2048.0f); t.a[1] = __float2int_rn(c1 * 2048.0f);
  t.a[2] = __float2int_rn(c2 * 2048.0f); t.a[3] = __float2int_rn(c3 * 2048.0f);
  return t;
}

// mode 0: no resize (patch side == target), 1: integer box sums (ix x iy cells), 2: general area tables,
// 3: INTER_CUBIC up-sampling (adv_patch.py:158-160)
__global__ void __launch_bounds__(kThreads) k_adv_area_paste(const uint8_t* __restrict__ matched, int src_h, int src_w,
                                                             int ph, int pw, int mode, int ix, int iy, double scale_y,
                                                             double scale_x, const double* __restrict__ noise,
                                                             uint8_t* frame, int frame_w, int y0, int x0) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < ph * pw; p += gridDim.x * blockDim.x) {
    const int dy = p / pw, dx = p - dy * pw;
    int v[3];
    if (mode == 0) {
      const uint8_t* q = matched + ((size_t)dy * src_w + dx) * 3;
      v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
    } else if (mode == 1) {
      int s[3] = {0, 0, 0};
      for (int yy = 0; yy < iy; ++yy) {
        const uint8_t* q = matched + ((size_t)(dy * iy + yy) * src_w + (size_t)dx * ix) * 3;
        for (int xx = 0; xx < ix; ++xx) { s[0] += q[xx * 3]; s[1] += q[xx * 3 + 1]; s[2] += q[xx * 3 + 2]; }
      }
      if (ix == 2 && iy == 2) {
        v[0] = (s[0] + 2) >> 2; v[1] = (s[1] + 2) >> 2; v[2] = (s[2] + 2) >> 2;
      } else {
        const float sc = 1.0f / (float)(ix * iy);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = sat_u8(__float2int_rn((float)s[c] * sc));
      }
    } else if (mode == 3) {
      // HResizeCubic (int32, replicated borders) of the four source rows, then VResizeCubic: the vectorised part of the
      // row (8 elements per step in the baseline build) in float32 without fused multiply-adds, its tail with the
      // 22-bit rounding shift
      const CubicTaps tx = cubic_taps(dx, scale_x), ty = cubic_taps(dy, scale_y);
      int hrow[4][3];
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int sy = min(max(ty.s - 1 + ky, 0), src_h - 1);
        hrow[ky][0] = hrow[ky][1] = hrow[ky][2] = 0;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int sx = min(max(tx.s - 1 + kx, 0), src_w - 1);
          const uint8_t* q = matched + ((size_t)sy * src_w + sx) * 3;
          hrow[ky][0] += (int)q[0] * tx.a[kx]; hrow[ky][1] += (int)q[1] * tx.a[kx]; hrow[ky][2] += (int)q[2] * tx.a[kx];
        }
      }
      const int nvec = (pw * 3 / 8) * 8;
      const float sc = 1.0f / (2048.0f * 2048.0f);
      const float b0 = (float)ty.a[0] * sc, b1 = (float)ty.a[1] * sc, b2 = (float)ty.a[2] * sc, b3 = (float)ty.a[3] * sc;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (dx * 3 + c < nvec) {
          float t = (float)hrow[3][c] * b3;
          t = (float)hrow[2][c] * b2 + t;
          t = (float)hrow[1][c] * b1 + t;
          t = (float)hrow[0][c] * b0 + t;
          v[c] = sat_u8(__float2int_rn(t));
        } else {
          const long long acc = (long long)hrow[0][c] * ty.a[0] + (long long)hrow[1][c] * ty.a[1] +
                                (long long)hrow[2][c] * ty.a[2] + (long long)hrow[3][c] * ty.a[3];
          v[c] = sat_u8((int)((acc + (1ll << 21)) >> 22));
        }
      }
    } else {
      const AreaTaps tx = area_taps(dx, scale_x, src_w), ty = area_taps(dy, scale_y, src_h);
      float acc[3] = {0.0f, 0.0f, 0.0f}, buf[3];
      bool started = false;
      if (ty.has_first) {
        hsum3(matched + (size_t)ty.first * src_w * 3, tx, buf);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = acc[c] + ty.a_first * buf[c];
        started = true;
      }
      for (int sy = ty.s_mid0; sy < ty.s_mid1; ++sy) {
        hsum3(matched + (size_t)sy * src_w * 3, tx, buf);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = started ? acc[c] + ty.a_mid * buf[c] : 0.0f + ty.a_mid * buf[c];
        started = true;
      }
      if (ty.has_last) {
        hsum3(matched + (size_t)ty.s_last * src_w * 3, tx, buf);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = acc[c] + ty.a_last * buf[c];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = sat_u8(__float2int_rn(acc[c]));   // saturate_cast<uchar>: round half to even
    }
    // get_transformed_patch (adv_patch.py:171-177), float64
    uint8_t* o = frame + ((size_t)(y0 + dy) * frame_w + (x0 + dx)) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double t = ((double)v[c] - 127.0) / 128.0;
      t = fmin(fmax(t + noise[(size_t)p * 3 + c], -1.0), 1.0);
      t = t * 128.0;
      t = t + 127.0;
      t = fmin(fmax(t, 0.0), 255.0);
      o[c] = (uint8_t)(int)t;
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_adv_print(const uint8_t* __restrict__ in, uint8_t* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double p = (double)in[i] - 127.0;
    p = p / 128.0;
    p = p * 0.5;
    p = p * 128.0;
    p = p + 127.0;
    out[i] = (uint8_t)(int)fmin(fmax(p, 0.0), 255.0);
  }
}

// adv_patch.py:61-92 in the reference's float arithmetic
static void create_host(int img_h, int img_w, const double* bb, double scale, int32_t out[4]) {
  const double ymin = bb[0], xmin = bb[1], ymax = bb[2], xmax = bb[3];
  const double h = ymax - ymin, w = xmax - xmin;
  const double long_side = h > w ? h : w;
  const int patch_w = (int)(long_side * scale), patch_h = patch_w;
  double ymin_patch = ymin + h / 2.0 - patch_h / 2.0;
  double xmin_patch = xmin + w / 2.0 - patch_w / 2.0;
  if (!(ymin_patch > 0.0)) ymin_patch = 0.0;
  if (!(xmin_patch > 0.0)) xmin_patch = 0.0;
  if (ymin_patch + patch_h > img_h) ymin_patch = img_h - patch_h;
  if (xmin_patch + patch_w > img_w) xmin_patch = img_w - patch_w;
  out[0] = (int)ymin_patch; out[1] = (int)xmin_patch; out[2] = patch_h; out[3] = patch_w;
}

static int grid_for(long long n) {
  long long g = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace eot

using namespace eot;

extern "C" int adv_u8_box_geometry(int32_t frame_h, int32_t frame_w, double scale, const double* boxes, int32_t n,
                                   int32_t* placements) {
  if (!boxes || !placements) { set_error("adv_u8_box_geometry: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (frame_h <= 0 || frame_w <= 0 || n < 0) { set_error("adv_u8_box_geometry: bad shape"); return EOT_ERR_BAD_SHAPE; }
  for (int i = 0; i < n; ++i) create_host(frame_h, frame_w, boxes + (size_t)i * 4, scale, placements + (size_t)i * 4);
  return EOT_OK;
}

extern "C" int adv_u8_print_patch(const uint8_t* patch, uint8_t* printed, int64_t n_elems, void* stream) {
  if (!patch || !printed) { set_error("adv_u8_print_patch: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (n_elems <= 0) { set_error("adv_u8_print_patch: empty patch"); return EOT_ERR_BAD_SHAPE; }
  k_adv_print<<<grid_for(n_elems), kThreads, 0, (cudaStream_t)stream>>>(patch, printed, n_elems);
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int adv_u8_workspace_bytes(int32_t patch_h, int32_t patch_w, size_t* bytes) {
  if (!bytes) { set_error("bytes is NULL"); return EOT_ERR_NULL_POINTER; }
  if (patch_h <= 0 || patch_w <= 0) { set_error("adv_u8_workspace_bytes: bad patch size"); return EOT_ERR_BAD_SHAPE; }
  *bytes = 256 + align_up((size_t)patch_h * patch_w * 3, 256);
  return EOT_OK;
}

extern "C" int adv_u8_add_patches(uint8_t* frame, int32_t frame_h, int32_t frame_w, const uint8_t* patch_printed,
                                  int32_t patch_h, int32_t patch_w, int32_t out_h, int32_t out_w, double scale,
                                  const double* boxes, int32_t n, const double* noise, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!frame || !patch_printed || !workspace || (n > 0 && (!boxes || !noise))) { set_error("adv_u8_add_patches: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (frame_h <= 0 || frame_w <= 0 || patch_h <= 0 || patch_w <= 0 || out_h <= 0 || out_w <= 0 || n < 0) {
    set_error("adv_u8_add_patches: bad shape");
    return EOT_ERR_BAD_SHAPE;
  }
  size_t need = 0;
  adv_u8_workspace_bytes(patch_h, patch_w, &need);
  if (workspace_bytes < need) { set_error("workspace too small: %zu < %zu", workspace_bytes, need); return EOT_ERR_WORKSPACE_TOO_SMALL; }
  if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return EOT_ERR_MISALIGNED; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* sums = static_cast<unsigned long long*>(workspace);
  uint8_t* matched = static_cast<uint8_t*>(workspace) + 256;
  // rescale(): adv_patch.py:101-107
  const double isy = (double)out_h / (double)frame_h, isx = (double)out_w / (double)frame_w;
  const double image_scale = isx < isy ? isx : isy;
  const int sh = (int)((double)frame_h * image_scale), sw = (int)((double)frame_w * image_scale);
  if (sh <= 0 || sw <= 0 || sh > out_h || sw > out_w) { set_error("adv_u8_add_patches: frame %dx%d rescales to %dx%d", frame_h, frame_w, sh, sw); return EOT_ERR_BAD_SHAPE; }
  const int fmode = (sh == frame_h && sw == frame_w) ? 0 : ((frame_h == 2 * sh && frame_w == 2 * sw) ? 1 : 2);
  const double fscale_y = 1.0 / ((double)sh / (double)frame_h), fscale_x = 1.0 / ((double)sw / (double)frame_w);
  const int n_px = patch_h * patch_w;
  // geometry of every box first: nothing is enqueued if one of them is not supported
  size_t noise_off = 0;
  for (int i = 0; i < n; ++i) {
    int32_t pl[4];
    create_host(frame_h, frame_w, boxes + (size_t)i * 4, scale, pl);
    if (pl[2] <= 0 || pl[3] <= 0) { set_error("adv_u8_add_patches: box %d gives an empty patch (cv2.resize would fail)", i); return EOT_ERR_BAD_SHAPE; }
    if (pl[0] < 0 || pl[1] < 0 || pl[0] + pl[2] > frame_h || pl[1] + pl[3] > frame_w) { set_error("adv_u8_add_patches: box %d does not fit the frame", i); return EOT_ERR_BAD_SHAPE; }
    // resize() compares the heights only (adv_patch.py:154-160): equal height + different width would fail in the paste
    if (pl[2] == patch_h && pl[3] != patch_w) { set_error("adv_u8_add_patches: box %d: patch height matches the texture but the width does not", i); return EOT_ERR_BAD_SHAPE; }
  }
  EOT_CHECK_CUDA(cudaMemsetAsync(sums, 0, 16, st));
  k_adv_patch_ysum<<<grid_for(n_px), kThreads, 0, st>>>(patch_printed, n_px, sums);
  count_launches(1);
  for (int i = 0; i < n; ++i) {
    int32_t pl[4];
    create_host(frame_h, frame_w, boxes + (size_t)i * 4, scale, pl);
    const int ph = pl[2], pw = pl[3];
    EOT_CHECK_CUDA(cudaMemsetAsync(sums + 1, 0, 8, st));
    k_adv_frame_ysum<<<grid_for((long long)sh * sw), kThreads, 0, st>>>(frame, frame_h, frame_w, sh, sw, fmode, fscale_y, fscale_x, sums + 1);
    k_adv_match<<<grid_for(n_px), kThreads, 0, st>>>(patch_printed, n_px, sums, (long long)out_h * out_w - (long long)sh * sw,
                                                    (long long)out_h * out_w, matched);
    int mode = 2, ix = 1, iy = 1;
    const double sc_x = (double)patch_w / (double)pw, sc_y = (double)patch_h / (double)ph;
    if (patch_h == ph) {
      mode = 0;                                                      // same height: no resize (validated above)
    } else if (patch_h < ph) {
      mode = 3;                                                      // bicubic up-sampling (adv_patch.py:158-160)
    } else {
      ix = (int)lrint(sc_x); iy = (int)lrint(sc_y);
      if (fabs(sc_x - ix) < 2.220446049250313e-16 && fabs(sc_y - iy) < 2.220446049250313e-16) mode = 1;
    }
    k_adv_area_paste<<<grid_for((long long)ph * pw), kThreads, 0, st>>>(matched, patch_h, patch_w, ph, pw, mode, ix, iy, sc_y, sc_x,
                                                                       noise + noise_off, frame, frame_w, pl[0], pl[1]);
    count_launches(3);
    noise_off += (size_t)ph * pw * 3;
  }
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
