// Forward of the EOT patch application: Patcher.call / Masker.call
// (reference: attacker.py:344-498, attack_detection.py:321-498, brightness_matcher.py:25-73).
//
// Kernels (all on the caller's stream, no host sync; ragged counts come in as CSR):
//   k_prepass         ONE launch, three roles by block index:
//                       geometry    Patcher.create + area filter + int cast, span / inverse-span / transposed
//                                   tables, -2 ring of the texel buffer, zeroed route map, work-item counts
//                       statistics  mean Y of the print-adjusted patch, per image
//                       image pass  copy image -> out (128-bit I/O) while accumulating mean Y of the target
//                                   image in float64 (the HBM-bound part)
//   k_match           print adjust + brightness match -> matched patch per image
//   k_resize          antialiased triangle resize (rows then columns, sequential float32
//                     accumulation as ScaleAndTranslate) + noise + brightness delta -> u_j (RGBX texels)
//   k_composite       gather over image tiles: projective bilinear sample of the ring-padded u_j of the
//                     covering boxes, newest first, `< -1` mask, background select, clip, store, route bytes
#include "eot_common.cuh"

#include <math.h>

namespace eot {

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float centre_clamp(float c, float extent, float limit) {
  float lo = fmaxf(c - extent / 2.0f, 0.0f);
  if (lo + extent > limit) lo = limit - extent;
  return lo;
}

// Patcher.create (attacker.py:448-488), area filter (:392-394), int cast (:418), pad split (:431-433),
// rotation transform (tfa angles_to_projective_transforms) and its inverse.
__device__ inline BoxPlan make_plan(const float* __restrict__ box, float shared_scale, const EotBoxParams& p,
                                    const EotShape& s, int image, int first, int last, int64_t u_off,
                                    int* err_flag) {
  BoxPlan pl;
  const float H = (float)s.height, W = (float)s.width;
  const float ymin = box[0], xmin = box[1], ymax = box[2], xmax = box[3];
  const float scale = p.scale >= 0.0f ? p.scale : shared_scale;
  const float tol = s.tolerance;
  const float h = ymax - ymin, w = xmax - xmin;
  const float longer = fmaxf(h, w);
  const float psf = floorf(longer * scale);
  const float diag = fminf(EOT_SQRT2 * psf, W);
  const float lo_y = (-tol * h) / 2.0f, hi_y = (tol * h) / 2.0f;
  const float jy = p.uy * (hi_y - lo_y) + lo_y;
  const float lo_x = (-tol * w) / 2.0f, hi_x = (tol * w) / 2.0f;
  const float jx = p.ux * (hi_x - lo_x) + lo_x;
  const float cy = (ymin + h / 2.0f) + jy;
  const float cx = (xmin + w / 2.0f) + jx;
  const float y0f = centre_clamp(cy, diag, H);
  const float x0f = centre_clamp(cx, diag, W);
  pl.valid = (psf * psf > s.min_patch_area) ? 1 : 0;
  pl.y0 = (int)y0f;
  pl.x0 = (int)x0f;
  pl.ps = (int)psf;
  pl.d = (int)diag;
  const int off2 = pl.d - pl.ps;
  pl.pad_lo = off2 >= 0 ? off2 / 2 : -((-off2 + 1) / 2);   // floor((d-ps)/2)
  pl.pad_hi = off2 - pl.pad_lo;                            // ceil((d-ps)/2)
  pl.span = 0;
  pl.image = image;
  pl.first_box = first;
  pl.last_box = last;
  pl.u_off = u_off;
  pl.delta = p.delta;
  pl.key0 = p.key0;
  pl.key1 = p.key1;
  if (pl.valid) {
    const Layout L = make_layout(s);
    const bool bad = !(psf == psf) || pl.y0 < 0 || pl.x0 < 0 || pl.y0 + pl.d > s.height || pl.x0 + pl.d > s.width ||
                     pl.d < pl.ps || pl.ps > L.lmin;
    if (bad) {   // TF would fail in tf.pad / tensor_scatter_nd_update; we skip the box and flag it
      pl.valid = 0;
      if (err_flag) atomicExch(err_flag, 1);
    }
  }
  const float c = p.cos_t, sn = p.sin_t;
  const float wm1 = (float)(pl.d - 1), hm1 = (float)(pl.d - 1);
  pl.T[0] = c;
  pl.T[1] = -sn;
  pl.T[2] = (wm1 - (c * wm1 - sn * hm1)) / 2.0f;
  pl.T[3] = sn;
  pl.T[4] = c;
  pl.T[5] = (hm1 - (sn * wm1 + c * hm1)) / 2.0f;
  pl.T[6] = p.pa;
  pl.T[7] = p.pb;
  {  // inverse by the adjugate in float64, rounded once (image_ops.py: _image_projective_transform_v3_grad)
    const double a = pl.T[0], b = pl.T[1], cc = pl.T[2], d = pl.T[3], e = pl.T[4], f = pl.T[5], g = pl.T[6], hh = pl.T[7];
    const double A = e - f * hh, Bm = -(d - f * g), C = d * hh - e * g;
    const double Dm = -(b - cc * hh), E = a - cc * g, Fm = -(a * hh - b * g);
    const double G = b * f - cc * e, Hm = -(a * f - cc * d), I = a * e - b * d;
    pl.Ti[0] = (float)(A / I);  pl.Ti[1] = (float)(Dm / I); pl.Ti[2] = (float)(G / I);
    pl.Ti[3] = (float)(Bm / I); pl.Ti[4] = (float)(E / I);  pl.Ti[5] = (float)(Hm / I);
    pl.Ti[6] = (float)(C / I);  pl.Ti[7] = (float)(Fm / I);
  }
  return pl;
}

// ScaleAndTranslate ComputeSpansCore (triangle kernel, antialias) for one output index.
struct SpanCfg { float inv_scale, ks, one_over; int span; };
__device__ __forceinline__ SpanCfg span_cfg(int out_size, int in_size) {
  SpanCfg c;
  const float scale = (float)out_size / (float)in_size;
  c.inv_scale = (float)(1.0 / (double)scale);
  c.ks = fmaxf(c.inv_scale, 1.0f);
  int span = 2 * (int)ceilf(1.0f * c.ks) + 1;
  c.span = span < in_size ? span : in_size;
  c.one_over = 1.0f / c.ks;
  return c;
}
__device__ __forceinline__ float tri_weight(int src, float sample_f, float one_over) {
  const float kernel_pos = ((float)src + 0.5f) - sample_f;
  const float a = fabsf(kernel_pos * one_over);
  return a < 1.0f ? 1.0f - a : 0.0f;
}
__device__ inline void span_row(int o, const SpanCfg& c, int in_size, int* start_out, float* __restrict__ w_out) {
  const float col_f = (float)o + 0.5f;
  const float sample_f = col_f * c.inv_scale + (-c.inv_scale * 0.0f);
  if (sample_f < 0.0f || sample_f > (float)in_size) {
    *start_out = 0;
    for (int k = 0; k < c.span; ++k) w_out[k] = 0.0f;
    return;
  }
  long long s0 = (long long)ceilf((sample_f - c.ks) - 0.5f);
  long long s1 = (long long)floorf((sample_f + c.ks) - 0.5f);
  s0 = s0 < 0 ? 0 : (s0 > in_size - 1 ? in_size - 1 : s0);
  s1 = (s1 < 0 ? 0 : (s1 > in_size - 1 ? in_size - 1 : s1)) + 1;
  int n = (int)(s1 - s0);
  if (n > c.span) n = c.span;
  float total = 0.0f;
  for (int k = 0; k < n; ++k) total = total + tri_weight((int)s0 + k, sample_f, c.one_over);
  const bool ok = fabsf(total) >= 1000.0f * 1.17549435e-38f;
  const float inv_total = ok ? 1.0f / total : 0.0f;
  for (int k = 0; k < c.span; ++k)
    w_out[k] = (k < n && ok) ? tri_weight((int)s0 + k, sample_f, c.one_over) * inv_total : 0.0f;
  *start_out = (int)s0;
}

// first output index whose span [start, start+span) can hold input index i, and the last one
// (starts[] is non-decreasing).  Used by the resize adjoint of the backward.
__device__ __forceinline__ int2 inverse_span_search(const int* starts, int n_out, int span, int i) {
  int a = 0, b = n_out;
  const int key = i - span + 1;
  while (a < b) { const int m = (a + b) >> 1; if (starts[m] < key) a = m + 1; else b = m; }
  const int lo = a;
  a = 0; b = n_out;
  while (a < b) { const int m = (a + b) >> 1; if (starts[m] <= i) a = m + 1; else b = m; }
  return make_int2(lo, a - 1);
}

// One CTA per box: plan, span table, inverse-span table, work-item counts.
__device__ int geometry_block(const EotShape& s, const Layout& L, int j, const float* __restrict__ boxes,
                               const int32_t* __restrict__ offsets, const EotBoxParams* __restrict__ params,
                               const float* __restrict__ scale, char* ws, EotBoxGeometry* geom_out) {
  __shared__ BoxPlan spl;
  __shared__ int s_maxcnt;
  int* counters = ws ? reinterpret_cast<int*>(ws + L.off_counters) : nullptr;
  if (threadIdx.x == 0) {
    s_maxcnt = 0;
    int a = 0, b = s.batch;                       // image of box j: last b with offsets[b] <= j
    while (a < b) { const int m = (a + b) >> 1; if (offsets[m + 1] <= j) a = m + 1; else b = m; }
    BoxPlan pl = make_plan(boxes + (size_t)j * 4, *scale, params[j], s, a, offsets[a], offsets[a + 1],
                           (int64_t)j * L.slot, counters ? counters + 2 : nullptr);
    if (pl.valid) pl.span = span_cfg(pl.ps, s.patch_size).span;
    spl = pl;
    if (ws) {
      reinterpret_cast<BoxPlan*>(ws + L.off_plans)[j] = pl;
      reinterpret_cast<int2*>(ws + L.off_cnt)[j] =
          pl.valid ? make_int2(fwd_strips(pl.ps, L.resize_rows), 0) : make_int2(0, 0);
    }
    if (geom_out) {
      EotBoxGeometry g = {pl.y0, pl.x0, pl.ps, pl.d, pl.pad_lo, pl.pad_hi, pl.valid, pl.span};
      geom_out[j] = g;
    }
  }
  __syncthreads();
  const int image = spl.image;
  if (!ws || !spl.valid) return image;
  const int ps = spl.ps, P = s.patch_size;
  int* starts = reinterpret_cast<int*>(ws + L.off_starts) + (size_t)j * L.lmin;
  float* weights = reinterpret_cast<float*>(ws + L.off_weights) + (size_t)j * L.wcap;
  const SpanCfg cfg = span_cfg(ps, P);
  for (int o = threadIdx.x; o < ps; o += blockDim.x) span_row(o, cfg, P, starts + o, weights + (size_t)o * cfg.span);
  __syncthreads();                                  // starts[] of this box are complete
  int2* inv = reinterpret_cast<int2*>(ws + L.off_inv) + (size_t)j * P;
  float* wt = reinterpret_cast<float*>(ws + L.off_wt) + (size_t)j * P * L.tcap;
  int2* stt = reinterpret_cast<int2*>(ws + L.off_stt) + (size_t)j * (P + 1);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const int2 rng = inverse_span_search(starts, ps, cfg.span, i);
    inv[i] = rng;
    int cnt = rng.y - rng.x + 1;
    if (cnt < 0) cnt = 0;
    if (cnt > L.tcap) { cnt = L.tcap; if (counters) atomicExch(counters + 2, 4); }
    stt[i] = make_int2(min(max(rng.x, 0), ps - 1), cnt);
    atomicMax(&s_maxcnt, cnt);
  }
  __syncthreads();
  // transposed weights (ScaleAndTranslateGrad = scatter of the same weights): row i of W^T, dense from rng.x, row
  // stride = the largest tap count of this box
  const int tstride = max(s_maxcnt, 1);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const int2 rng = inv[i];
    const int cnt = stt[i].y;
    for (int k = 0; k < tstride; ++k) {
      const int o = rng.x + k;
      float w = 0.0f;
      if (k < cnt) {
        const int kk = i - starts[o];
        if (kk >= 0 && kk < cfg.span) w = weights[(size_t)o * cfg.span + kk];
      }
      wt[(size_t)i * tstride + k] = w;
    }
  }
  if (threadIdx.x == 0) {
    const int rows = bwd_strip_rows(ps);
    stt[P] = make_int2(tstride, rows);
    reinterpret_cast<int2*>(ws + L.off_cnt)[j].y = (P + rows - 1) / rows;   // backward resize strips of this box
  }
  // route map of the box starts all-zero; the composite only writes the non-zero bytes
  uint4* rz = reinterpret_cast<uint4*>(ws + L.off_route + (size_t)j * L.rslot);
  const int n16 = (spl.d * spl.d + 15) / 16;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) rz[i] = make_uint4(0u, 0u, 0u, 0u);
  // -2 ring (pad of attacker.py:435 / fill of :437) around the ps x ps core the resize items write
  float4* u4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(ws + L.off_u) + spl.u_off);
  const int S = u_stride(ps);
  const float4 fill = make_float4(-2.0f, -2.0f, -2.0f, 0.0f);
  for (int i = threadIdx.x; i < 4 * S; i += blockDim.x) {
    const int r = i / S, c = i - r * S;
    u4[(r < 2 ? r : S - 4 + r) * S + c] = fill;
  }
  for (int i = threadIdx.x; i < 4 * ps; i += blockDim.x) {
    const int r = i >> 2, k = i & 3;
    u4[(r + 2) * S + (k < 2 ? k : S - 4 + k)] = fill;
  }
  return image;
}

// Exclusive prefix sums of the per-box work-item counts (one CTA; N is a few hundred to a few thousand).
__device__ void scan_block(int N, const int2* cnt, int2* base, int2* part /* [blockDim.x] shared */) {
  const int T = blockDim.x;
  const int per = (N + T - 1) / T;
  const int j0 = threadIdx.x * per, j1 = min(N, j0 + per);
  int2 sum = make_int2(0, 0);
  for (int j = j0; j < j1; ++j) { const int2 c = __ldcg(cnt + j); sum.x += c.x; sum.y += c.y; }
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < T; d <<= 1) {
    int2 v = make_int2(0, 0);
    if ((int)threadIdx.x >= d) v = part[threadIdx.x - d];
    __syncthreads();
    part[threadIdx.x].x += v.x;
    part[threadIdx.x].y += v.y;
    __syncthreads();
  }
  int2 run = threadIdx.x ? part[threadIdx.x - 1] : make_int2(0, 0);
  for (int j = j0; j < j1; ++j) {
    base[j] = run;
    const int2 c = __ldcg(cnt + j);
    run.x += c.x;
    run.y += c.y;
  }
  if ((int)threadIdx.x == T - 1) base[N] = part[T - 1];
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) k_geometry_only(EotShape s, Layout L, const float* __restrict__ boxes,
                                                            const int32_t* __restrict__ offsets,
                                                            const EotBoxParams* __restrict__ params,
                                                            const float* __restrict__ scale, EotBoxGeometry* geom_out) {
  geometry_block(s, L, blockIdx.x, boxes, offsets, params, scale, nullptr, geom_out);
}

// ------------------------------------------------------------------------------------------------
// patch luma statistics: mean Y of rescale(print_adjust(patch)) per image
// (attacker.py:372 -> brightness_matcher.py:54,58,62-63)
// ------------------------------------------------------------------------------------------------
__device__ void patch_stats_block(const EotShape& s, int b, int chunk, int nchunks, const float* __restrict__ patch,
                                  const float* __restrict__ print_wb, double* ysum_patch, double* red) {
  const int P = s.patch_size;
  const float* wb = print_wb + (size_t)b * 6;
  const float* base = patch + (s.num_patches > 1 ? (int64_t)b * s.patch_stride_n : 0);
  double acc = 0.0;
  for (int t = chunk * blockDim.x + threadIdx.x; t < P * P; t += nchunks * blockDim.x) {
    const int py = t / P, px = t - py * P;
    const float* p = base + (int64_t)py * s.patch_stride_y + (int64_t)px * s.patch_stride_x;
    acc += (double)texel_yuv(__ldg(p), __ldg(p + 1), __ldg(p + 2), wb).y;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(ysum_patch + b, acc);
}

// ------------------------------------------------------------------------------------------------
// image pass: out = image (128-bit loads/stores, 4 pixels = 3 x float4 per thread per step) and
// sum of Y over the image (brightness_matcher.py:55,59,62,64) in float64.
// ------------------------------------------------------------------------------------------------
#ifndef EOT_PASS_PIX
#define EOT_PASS_PIX 16
#endif
constexpr int kPassPixPerThread = EOT_PASS_PIX;   // batches of 2 x 4 pixels
constexpr int kPassPixPerBlock = kThreads * kPassPixPerThread;

__device__ __forceinline__ float luma_of(float r, float g, float b) {
  const float t0 = (r + 1.0f) * EOT_C127_255, t1 = (g + 1.0f) * EOT_C127_255, t2 = (b + 1.0f) * EOT_C127_255;
  return (t0 * EOT_K00 + t1 * EOT_K10) + t2 * EOT_K20;
}

template <bool kVec>
__device__ __forceinline__ void image_pass_block(int HW, int b, int chunk, const float* __restrict__ images, float* out,
                                                 float* mask, double* ysum_img, int* oor_flags, double* red) {
  const size_t img_off = (size_t)b * HW * 3;
  bool oor = false;                                     // any value outside [-1,1] (or NaN)
  const float* in = images + img_off;
  float* o = (out && out != images) ? out + img_off : nullptr;
  float* mk = mask ? mask + img_off : nullptr;
  double acc = 0.0;
  const int pix0 = chunk * kPassPixPerBlock;
  if (kVec) {
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* o4 = reinterpret_cast<float4*>(o);
    float4* m4 = reinterpret_cast<float4*>(mk);
    // two batches of 2 x (3 x float4): 96 bytes per thread in flight per batch, modest register footprint
#pragma unroll 1
    for (int half = 0; half < kPassPixPerThread / 8; ++half) {
      float4 v[2][3];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int pix = pix0 + ((half * 2 + k) * kThreads + threadIdx.x) * 4;
        if (pix < HW) {
          const int q = (pix >> 2) * 3;
          v[k][0] = __ldg(in4 + q);
          v[k][1] = __ldg(in4 + q + 1);
          v[k][2] = __ldg(in4 + q + 2);
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int pix = pix0 + ((half * 2 + k) * kThreads + threadIdx.x) * 4;
        if (pix < HW) {
          const int q = (pix >> 2) * 3;
          const float4 a = v[k][0], bb = v[k][1], c = v[k][2];
          if (o) { o4[q] = a; o4[q + 1] = bb; o4[q + 2] = c; }
          if (mk) { const float4 z = make_float4(0.f, 0.f, 0.f, 0.f); m4[q] = z; m4[q + 1] = z; m4[q + 2] = z; }
          acc += (double)luma_of(a.x, a.y, a.z);
          acc += (double)luma_of(a.w, bb.x, bb.y);
          acc += (double)luma_of(bb.z, bb.w, c.x);
          acc += (double)luma_of(c.y, c.z, c.w);
          const float m0 = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
          const float m1 = fmaxf(fmaxf(fabsf(bb.x), fabsf(bb.y)), fmaxf(fabsf(bb.z), fabsf(bb.w)));
          const float m2 = fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w)));
          const float sum = (a.x + a.y + a.z + a.w) + (bb.x + bb.y + bb.z + bb.w) + (c.x + c.y + c.z + c.w);
          oor = oor || !(fmaxf(fmaxf(m0, m1), m2) <= 1.0f) || (sum != sum);      // fmaxf drops NaN: test the sum too
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < kPassPixPerBlock; i += kThreads) {
      const int pix = pix0 + i;
      if (pix < HW) {
        const float r = in[(size_t)pix * 3], g = in[(size_t)pix * 3 + 1], bl = in[(size_t)pix * 3 + 2];
        if (o) { o[(size_t)pix * 3] = r; o[(size_t)pix * 3 + 1] = g; o[(size_t)pix * 3 + 2] = bl; }
        if (mk) { mk[(size_t)pix * 3] = 0.f; mk[(size_t)pix * 3 + 1] = 0.f; mk[(size_t)pix * 3 + 2] = 0.f; }
        acc += (double)luma_of(r, g, bl);
        oor = oor || !(fabsf(r) <= 1.0f) || !(fabsf(g) <= 1.0f) || !(fabsf(bl) <= 1.0f);
      }
    }
  }
  const int any_oor = __syncthreads_or(oor ? 1 : 0);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(ysum_img + b, acc);
    if (any_oor) atomicOr(oor_flags + b, 1);
  }
}

// One launch, three independent roles selected by the block index (they only meet at k_match):
//   [0, n_geom)                         geometry of box j                   (n_geom = N or 0)
//   [.., + n_stat_imgs*pchunks)         patch luma statistics               (n_stat_imgs = B or 0)
//   [.., + (b1-b0)*cpi)                 image pass (copy + luma sum) of images [b0,b1), the HBM-bound bulk
template <bool kVec>
__global__ void __launch_bounds__(kThreads) k_prepass(EotShape s, Layout L, const float* __restrict__ patch,
                                                      const float* __restrict__ print_wb, const float* __restrict__ boxes,
                                                      const int32_t* __restrict__ offsets,
                                                      const EotBoxParams* __restrict__ params,
                                                      const float* __restrict__ scale, const float* __restrict__ images,
                                                      float* out, float* mask, char* ws, int n_geom, int n_stat_imgs,
                                                      int pchunks, int cpi, int b0) {
  __shared__ double red[32];
  int blk = blockIdx.x;
  if (blk < n_geom) {
    __shared__ int s_last;
    __shared__ int2 s_part[kThreads];
    geometry_block(s, L, blk, boxes, offsets, params, scale, ws, nullptr);
    __threadfence();
    __syncthreads();
    int* counters = reinterpret_cast<int*>(ws + L.off_counters);
    if (threadIdx.x == 0) s_last = (atomicAdd(counters + 5, 1) == n_geom - 1);
    __syncthreads();
    if (s_last) {                                        // last geometry block: prefix sums of the work-item counts
      __threadfence();
      scan_block(n_geom, reinterpret_cast<const int2*>(ws + L.off_cnt), reinterpret_cast<int2*>(ws + L.off_base), s_part);
    }
    return;
  }
  blk -= n_geom;
  if (blk < n_stat_imgs * pchunks) {
    const int b = blk / pchunks;
    patch_stats_block(s, b, blk - b * pchunks, pchunks, patch, print_wb, reinterpret_cast<double*>(ws + L.off_ysum_patch), red);
    if (blk == 0 && threadIdx.x == 0) {                  // CSR copy for the backward
      int32_t* off_copy = reinterpret_cast<int32_t*>(ws + L.off_offsets);
      for (int i = 0; i <= s.batch; ++i) off_copy[i] = offsets[i];
    }
    return;
  }
  blk -= n_stat_imgs * pchunks;
  const int b = b0 + blk / cpi, chunk = blk % cpi;
  image_pass_block<kVec>(s.height * s.width, b, chunk, images, out, mask, reinterpret_cast<double*>(ws + L.off_ysum_img),
                         reinterpret_cast<int*>(ws + L.off_oor), red);
}

// ------------------------------------------------------------------------------------------------
// print adjust + brightness match of the patch for image b (attacker.py:372; brightness_matcher.py:43-73)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void match_block(const EotShape& s, const Layout& L, const float* __restrict__ patch,
                                            const float* __restrict__ print_wb, char* ws, int b, int part, int nparts);

__global__ void __launch_bounds__(kThreads) k_match(EotShape s, Layout L, const float* __restrict__ patch,
                                                    const float* __restrict__ print_wb, char* ws, int b0) {
  match_block(s, L, patch, print_wb, ws, b0 + blockIdx.y, blockIdx.x, gridDim.x);
}

__device__ __forceinline__ void match_block(const EotShape& s, const Layout& L, const float* __restrict__ patch,
                                            const float* __restrict__ print_wb, char* ws, int b, int part, int nparts) {
  const int P = s.patch_size;
  const double* ysum_img = reinterpret_cast<const double*>(ws + L.off_ysum_img);
  const double* ysum_patch = reinterpret_cast<const double*>(ws + L.off_ysum_patch);
  const float mu_t = (float)(ysum_img[b] / (double)((size_t)s.height * s.width));
  const float mu_s = (float)(ysum_patch[b] / (double)((size_t)P * P));
  const float* wb = print_wb + (size_t)b * 6;
  const float* base = patch + (s.num_patches > 1 ? (int64_t)b * s.patch_stride_n : 0);
  float* m = reinterpret_cast<float*>(ws + L.off_match) + (size_t)b * P * P * 3;
  for (int t = part * blockDim.x + threadIdx.x; t < P * P; t += nparts * blockDim.x) {
    const int py = t / P, px = t - py * P;
    const float* p = base + (int64_t)py * s.patch_stride_y + (int64_t)px * s.patch_stride_x;
    const TexelYuv y = texel_yuv(__ldg(p), __ldg(p + 1), __ldg(p + 2), wb);
    const float yp = clampf((y.y - mu_s) + mu_t, 0.0f, 1.0f);
    // rgb = [Y',U,V] . K'  as ((Y'*K'0c + U*K'1c) + V*K'2c)
    const float r = (yp * 1.0f + y.u * EOT_I10) + y.v * EOT_I20;
    const float g = (yp * 1.0f + y.u * EOT_I11) + y.v * EOT_I21;
    const float bl = (yp * 1.0f + y.u * EOT_I12) + y.v * EOT_I22;
    m[(size_t)t * 3 + 0] = clampf(r, 0.0f, 1.0f) * EOT_C255_127 - 1.0f;
    m[(size_t)t * 3 + 1] = clampf(g, 0.0f, 1.0f) * EOT_C255_127 - 1.0f;
    m[(size_t)t * 3 + 2] = clampf(bl, 0.0f, 1.0f) * EOT_C255_127 - 1.0f;
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone BrightnessMatcher()((src, tgt)) (brightness_matcher.py:43-73): no print adjust, no clip
// of the source before the rescale.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_bm_ysum(const float* __restrict__ x, long long n_pix, double* sum) {
  __shared__ double red[32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x)
    acc += (double)luma_of(__ldg(x + i * 3), __ldg(x + i * 3 + 1), __ldg(x + i * 3 + 2));
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(sum, acc);
}

__global__ void __launch_bounds__(kThreads) k_bm_apply(const float* __restrict__ src, long long n_src, long long n_tgt,
                                                       const double* __restrict__ sums, float* out) {
  const float mu_s = (float)(sums[0] / (double)n_src), mu_t = (float)(sums[1] / (double)n_tgt);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += (long long)gridDim.x * blockDim.x) {
    const float s0 = (src[i * 3] + 1.0f) * EOT_C127_255, s1 = (src[i * 3 + 1] + 1.0f) * EOT_C127_255,
                s2 = (src[i * 3 + 2] + 1.0f) * EOT_C127_255;
    const float y = (s0 * EOT_K00 + s1 * EOT_K10) + s2 * EOT_K20;
    const float u = (s0 * EOT_K01 + s1 * EOT_K11) + s2 * EOT_K21;
    const float v = (s0 * EOT_K02 + s1 * EOT_K12) + s2 * EOT_K22;
    const float yp = clampf((y - mu_s) + mu_t, 0.0f, 1.0f);
    const float r = (yp * 1.0f + u * EOT_I10) + v * EOT_I20;
    const float g = (yp * 1.0f + u * EOT_I11) + v * EOT_I21;
    const float b = (yp * 1.0f + u * EOT_I12) + v * EOT_I22;
    out[i * 3] = clampf(r, 0.0f, 1.0f) * EOT_C255_127 - 1.0f;
    out[i * 3 + 1] = clampf(g, 0.0f, 1.0f) * EOT_C255_127 - 1.0f;
    out[i * 3 + 2] = clampf(b, 0.0f, 1.0f) * EOT_C255_127 - 1.0f;
  }
}

// ------------------------------------------------------------------------------------------------
// resize + noise + delta + clip for one strip of L.resize_rows output rows of one box
// (attacker.py:425-428; ScaleAndTranslate GatherRows then GatherColumns).
// The box's span table is staged in shared memory first (no dependent global loads in the loops).
// Rows pass: a lane owns four texels (12 floats = 3 x 128-bit loads per tap) of a source row and writes
// them as RGBX texels into the shared intermediate.  Columns pass: a lane owns a quad of 4 consecutive
// output texels = 12 elements = exactly 3 Philox groups; one 128-bit shared load per tap.  Accumulation
// order == the oracle's.  Output texel: (clip r, clip g, clip b, inner-clip pass bits).
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t resize_smem_bytes(const EotShape& s, const Layout& L) {
  return ((size_t)L.resize_rows * s.patch_size * 4 + (size_t)L.wcap + (size_t)L.lmin) * sizeof(float);
}

// SPAN > 0: compile-time tap count (3 = up-sampling / unit scale, 5 and 7 = moderate down-sampling), every tap
// loop fully unrolled; weights past the true span are stored as 0 and the clamped source index is finite, so the
// padded taps add +0 -- the same sums as the oracle's.  SPAN == 0: run-time span (any scale).
template <int SPAN>
__device__ __forceinline__ void resize_passes(const EotShape& s, int P, int ps, int span, int oy0, int rows,
                                              const float* m, const int* s_st, const float* s_w, float4* inter,
                                              float4* u4, float delta, uint32_t key0, uint32_t key1) {
  const int P3 = P * 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  constexpr int NT = SPAN > 0 ? SPAN : 1;
  // ---- rows pass ----
  if ((P & 3) == 0) {
    const int nq = P >> 2;                                    // texel quads per source row
    // (row, quad) pairs flattened over the CTA (a warp per row would idle 7 lanes of 32 at P = 100)
    for (int idx = threadIdx.x; idx < rows * nq; idx += blockDim.x) {
      const int r = idx / nq;
      const int oy = oy0 + r;
      const int st = s_st[oy];
      const float* w = s_w + oy * (SPAN > 0 ? SPAN : span);
      const int nk = SPAN > 0 ? SPAN : min(span, P - st);
      {
        const int tq = idx - r * nq;
        float acc[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[i] = 0.0f;
        if (SPAN > 0) {
          float4 v[NT][3];
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            const float4* mp = reinterpret_cast<const float4*>(m + (size_t)min(st + k, P - 1) * P3) + 3 * tq;
            v[k][0] = mp[0]; v[k][1] = mp[1]; v[k][2] = mp[2];
          }
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            const float wk = w[k];
            const float x[12] = {v[k][0].x, v[k][0].y, v[k][0].z, v[k][0].w, v[k][1].x, v[k][1].y, v[k][1].z, v[k][1].w,
                                 v[k][2].x, v[k][2].y, v[k][2].z, v[k][2].w};
#pragma unroll
            for (int i = 0; i < 12; ++i) acc[i] = acc[i] + wk * x[i];
          }
        } else {
          for (int k = 0; k < nk; ++k) {
            const float4* mp = reinterpret_cast<const float4*>(m + (size_t)(st + k) * P3) + 3 * tq;
            const float4 a = mp[0], b = mp[1], c = mp[2];
            const float wk = w[k];
            const float x[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 12; ++i) acc[i] = acc[i] + wk * x[i];
          }
        }
        float4* o = inter + r * P + 4 * tq;
        o[0] = make_float4(acc[0], acc[1], acc[2], 0.0f);
        o[1] = make_float4(acc[3], acc[4], acc[5], 0.0f);
        o[2] = make_float4(acc[6], acc[7], acc[8], 0.0f);
        o[3] = make_float4(acc[9], acc[10], acc[11], 0.0f);
      }
    }
  } else {                                                    // any P: one lane per texel, scalar loads
    for (int r = warp; r < rows; r += nwarps) {
      const int oy = oy0 + r;
      const int st = s_st[oy];
      const float* w = s_w + oy * (SPAN > 0 ? SPAN : span);
      const int nk = SPAN > 0 ? SPAN : min(span, P - st);
      for (int x = lane; x < P; x += 32) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        for (int k = 0; k < nk; ++k) {
          const float* mp = m + (size_t)min(st + k, P - 1) * P3 + x * 3;
          const float wk = w[k];
          a0 = a0 + wk * mp[0];
          a1 = a1 + wk * mp[1];
          a2 = a2 + wk * mp[2];
        }
        inter[r * P + x] = make_float4(a0, a1, a2, 0.0f);
      }
    }
  }
  __syncthreads();
  // ---- columns pass + noise + delta + clip ----
  const int S = u_stride(ps);
  const int p_begin = oy0 * ps, p_end = (oy0 + rows) * ps;   // flat texel range of the strip
  for (int q = (p_begin >> 2) + threadIdx.x; q <= ((p_end - 1) >> 2); q += blockDim.x) {
    uint32_t words[12];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const uint4 rnd = philox4x32_10((uint32_t)(3 * q + g), key0, key1);
      words[4 * g] = rnd.x; words[4 * g + 1] = rnd.y; words[4 * g + 2] = rnd.z; words[4 * g + 3] = rnd.w;
    }
    int p = 4 * q;
    int oy = p / ps, ox = p - oy * ps;
#pragma unroll
    for (int t = 0; t < 4; ++t, ++p) {
      if (p >= p_begin && p < p_end) {
        const int st = s_st[ox];
        const float4* irow = inter + (oy - oy0) * P;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        if (SPAN > 0) {
          const float* w = s_w + ox * SPAN;
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            const float wk = w[k];
            const float4 v = irow[min(st + k, P - 1)];
            a0 = a0 + wk * v.x;
            a1 = a1 + wk * v.y;
            a2 = a2 + wk * v.z;
          }
        } else {
          const float* w = s_w + ox * span;
          const int nk = min(span, P - st);
          for (int k = 0; k < nk; ++k) {
            const float wk = w[k];
            const float4 v = irow[st + k];
            a0 = a0 + wk * v.x;
            a1 = a1 + wk * v.y;
            a2 = a2 + wk * v.z;
          }
        }
        const float v0 = (a0 + noise_from_word(words[3 * t], s.noise_amp)) + delta;
        const float v1 = (a1 + noise_from_word(words[3 * t + 1], s.noise_amp)) + delta;
        const float v2 = (a2 + noise_from_word(words[3 * t + 2], s.noise_amp)) + delta;
        const unsigned bits = (unsigned)(v0 >= -1.0f && v0 <= 1.0f) | ((unsigned)(v1 >= -1.0f && v1 <= 1.0f) << 1) |
                              ((unsigned)(v2 >= -1.0f && v2 <= 1.0f) << 2);
        u4[(oy + 2) * S + ox + 2] =
            make_float4(clampf(v0, -1.0f, 1.0f), clampf(v1, -1.0f, 1.0f), clampf(v2, -1.0f, 1.0f), __uint_as_float(bits));
      }
      if (++ox == ps) { ox = 0; ++oy; }
    }
  }
}

__device__ __forceinline__ void resize_item(const EotShape& s, const Layout& L, char* ws, int2 item, float* smem) {
  const int P = s.patch_size, P3 = P * 3;
  float4* inter = reinterpret_cast<float4*>(smem);           // [resize_rows][P] RGBX
  float* s_w = smem + (size_t)L.resize_rows * P * 4;         // [ps][span]
  int* s_st = reinterpret_cast<int*>(s_w + L.wcap);          // [ps]
  const BoxPlan* pl = reinterpret_cast<const BoxPlan*>(ws + L.off_plans) + item.x;
  const int j = item.x;
  const int ps = pl->ps, span = pl->span;
  const float delta = pl->delta;
  const uint32_t key0 = pl->key0, key1 = pl->key1;
  const int RR = strip_rows(ps, reinterpret_cast<const int2*>(ws + L.off_cnt)[j].x);
  const int oy0 = item.y * RR;
  const int rows = min(RR, ps - oy0);
  const float* m = reinterpret_cast<const float*>(ws + L.off_match) + (size_t)pl->image * P * P3;
  const int* starts = reinterpret_cast<const int*>(ws + L.off_starts) + (size_t)j * L.lmin;
  const float* wts = reinterpret_cast<const float*>(ws + L.off_weights) + (size_t)j * L.wcap;
  float4* u4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(ws + L.off_u) + pl->u_off);
  for (int i = threadIdx.x; i < ps; i += blockDim.x) s_st[i] = starts[i];
  for (int i = threadIdx.x; i < ps * span; i += blockDim.x) s_w[i] = wts[i];
  __syncthreads();
  if (span == 3) resize_passes<3>(s, P, ps, span, oy0, rows, m, s_st, s_w, inter, u4, delta, key0, key1);
  else if (span == 5) resize_passes<5>(s, P, ps, span, oy0, rows, m, s_st, s_w, inter, u4, delta, key0, key1);
  else if (span == 7) resize_passes<7>(s, P, ps, span, oy0, rows, m, s_st, s_w, inter, u4, delta, key0, key1);
  else resize_passes<0>(s, P, ps, span, oy0, rows, m, s_st, s_w, inter, u4, delta, key0, key1);
}

#ifndef EOT_RESIZE_MINB
#define EOT_RESIZE_MINB 4
#endif
#ifndef EOT_RESIZE_THREADS
#define EOT_RESIZE_THREADS 256
#endif
__global__ void __launch_bounds__(EOT_RESIZE_THREADS, EOT_RESIZE_MINB) k_resize(EotShape s, Layout L, char* ws, const int32_t* __restrict__ offsets,
                                                        int b0, int b1) {
  extern __shared__ __align__(16) float resize_smem[];
  __shared__ int2 s_base[kMaxBaseSmem];
  const int2* base = stage_base(reinterpret_cast<const int2*>(ws + L.off_base), s.total_boxes, s_base);
  const int lo = base[offsets[b0]].x, hi = base[offsets[b1]].x;
  __shared__ int2 s_item;
  for (int it = lo + blockIdx.x; it < hi; it += gridDim.x) {
    if (threadIdx.x == 0) s_item = find_item(base, s.total_boxes, 0, it);
    __syncthreads();
    resize_item(s, L, ws, s_item, resize_smem);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// composite (attacker.py:436-444) in gather form over IMAGE tiles.
//
// Sequential-paste semantics: an element's final value is clip(R_j) of the LAST box j (in paste order)
// covering it with R_j >= -1, else clip(original); pixels outside every window are untouched.  One work
// item = a band of kCompRows rows of one image; the boxes whose windows meet the band are staged in shared
// memory once; a warp owns a row and sweeps it in 32-pixel tiles, one lane per pixel (sampling coordinates
// shared by the three channels).  Per tile the candidate boxes are visited newest first, warp-uniformly;
// a box is sampled only where the pixel is inside its window, still misses a channel, and at least one of
// the four taps can touch the ps x ps core (everything else blends to -2 exactly as the padded image
// would).  Every output pixel is written once, by one thread: no ordering between work items, no waits.
//
// Each pixel leaves a route byte in the map of every box that provided one of its channels (bit c: channel
// c of the output came from this box and passes the outer clip): the backward's TensorScatterUpdate /
// SelectV2 / clip routing without re-sampling.
// ------------------------------------------------------------------------------------------------
#ifndef EOT_COMP_THREADS
#define EOT_COMP_THREADS 128
#endif
constexpr int kCompThreads = EOT_COMP_THREADS;   // threads per composite CTA: one warp per band row
constexpr int kMaxBandBoxes = 32;   // boxes of one image whose windows meet one band (more: flagged, first 32 handled)

struct BandBox {
  float t0, t1, t2, t3, t4, t5, t6, t7;
  float lo2, hi;              // clamp range of the floor coordinates
  float ia0, ia3;             // 1/t0, 1/t3 (0 when the coefficient is ~0): column range of the core per row
  int org, S;
  int y0, x0, d, j;
  const float4* u;
  uint8_t* route;
};

struct CompositeSmem {
  BandBox box[kMaxBandBoxes];
  int n;
};

// Conservative range of window columns x in window row y whose sample can touch the ps x ps core (affine T):
// outside it all four taps are pad/fill and the box contributes nothing.
__device__ __forceinline__ void core_range(const BandBox& bx, float yf, int* xa, int* xb) {
  float lo = 0.0f, hi = (float)(bx.d - 1);
  const float clo = bx.lo2 + 2.0f, chi = bx.hi;               // core bounds in padded coordinates
  const float c0 = bx.t1 * yf + bx.t2, c1 = bx.t4 * yf + bx.t5;
  {
    const float l = clo - 1.5f - c0, h = chi + 0.5f - c0;      // need l < t0*x < h (half-pixel safety margin)
    if (bx.ia0 == 0.0f) { if (!(0.0f > l - 1.0f && 0.0f < h + 1.0f)) { lo = 1.0f; hi = 0.0f; } }
    else { const float x1 = l * bx.ia0, x2 = h * bx.ia0; lo = fmaxf(lo, fminf(x1, x2) - 1.0f); hi = fminf(hi, fmaxf(x1, x2) + 1.0f); }
  }
  {
    const float l = clo - 1.5f - c1, h = chi + 0.5f - c1;
    if (bx.ia3 == 0.0f) { if (!(0.0f > l - 1.0f && 0.0f < h + 1.0f)) { lo = 1.0f; hi = 0.0f; } }
    else { const float x1 = l * bx.ia3, x2 = h * bx.ia3; lo = fmaxf(lo, fminf(x1, x2) - 1.0f); hi = fminf(hi, fmaxf(x1, x2) + 1.0f); }
  }
  *xa = (int)floorf(lo);
  *xb = (int)ceilf(hi);
}

__device__ __forceinline__ void composite_band(const EotShape& s, const Layout& L, char* ws,
                                               const float* __restrict__ images, float* out, float* mask, int b, int band,
                                               const int32_t* __restrict__ offsets, CompositeSmem& sm) {
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  const int H = s.height, W = s.width;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ya = band * kCompRows, yb = min(H, ya + kCompRows);
  if (warp == 0) {                                              // stage the boxes meeting the band, in paste order
    const float* ubuf = reinterpret_cast<const float*>(ws + L.off_u);
    uint8_t* routes = reinterpret_cast<uint8_t*>(ws + L.off_route);
    const int first = offsets[b], last = offsets[b + 1];
    int n = 0;
    for (int q0 = first; q0 < last; q0 += 32) {
      const int q = q0 + lane;
      bool hit = false;
      if (q < last) {
        const BoxPlan* o = plans + q;
        hit = o->valid && o->y0 < yb && o->y0 + o->d > ya;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      const int pos = n + __popc(m & ((1u << lane) - 1u));
      if (hit && pos < kMaxBandBoxes) {
        const BoxPlan* o = plans + q;
        BandBox& bx = sm.box[pos];
        bx.t0 = o->T[0]; bx.t1 = o->T[1]; bx.t2 = o->T[2]; bx.t3 = o->T[3];
        bx.t4 = o->T[4]; bx.t5 = o->T[5]; bx.t6 = o->T[6]; bx.t7 = o->T[7];
        bx.lo2 = (float)(o->pad_lo - 2);
        bx.hi = (float)(o->pad_lo + o->ps);
        bx.ia0 = fabsf(o->T[0]) < 1e-6f ? 0.0f : 1.0f / o->T[0];
        bx.ia3 = fabsf(o->T[3]) < 1e-6f ? 0.0f : 1.0f / o->T[3];
        bx.org = o->pad_lo - 2;
        bx.S = o->ps + 4;
        bx.y0 = o->y0; bx.x0 = o->x0; bx.d = o->d; bx.j = q;
        bx.u = reinterpret_cast<const float4*>(ubuf + o->u_off);
        bx.route = routes + (size_t)q * L.rslot;
      }
      n += __popc(m);
    }
    if (n > kMaxBandBoxes) {
      if (lane == 0) atomicExch(reinterpret_cast<int*>(ws + L.off_counters) + 2, 2);
      n = kMaxBandBoxes;
    }
    if (lane == 0) sm.n = n;
  }
  __syncthreads();
  const int n = sm.n;
  if (n == 0) return;
  const bool oor = reinterpret_cast<const int*>(ws + L.off_oor)[b] != 0;
  // with a background clip that is not the identity, or the Masker's mask, every pixel inside a window is written
  const bool all_px = oor || mask != nullptr;
  const size_t img_off = (size_t)b * H * W * 3;
  const float* img = images + img_off;
  float* o_img = out + img_off;
  float* m_img = mask ? mask + img_off : nullptr;
  for (int gy = ya + warp; gy < yb; gy += kCompThreads / 32) {
    // lane i: the column range of this row in which box i can matter (its whole window when every window pixel is
    // written or T is projective, else the conservative range of the rotated core)
    int rx0 = 1, rx1 = 0;
    if (lane < n) {
      const BandBox& bx = sm.box[lane];
      if (gy >= bx.y0 && gy < bx.y0 + bx.d) {
        int xa = 0, xb = bx.d - 1;
        if (!all_px && bx.t6 == 0.0f && bx.t7 == 0.0f) {
          core_range(bx, (float)(gy - bx.y0), &xa, &xb);
          xa = max(xa, 0);
          xb = min(xb, bx.d - 1);
        }
        rx0 = bx.x0 + xa;
        rx1 = bx.x0 + xb;
      }
    }
    if (!__any_sync(0xffffffffu, rx0 <= rx1)) continue;
    int xlo = rx0 <= rx1 ? rx0 : W, xhi = rx0 <= rx1 ? rx1 : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
      xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
    }
    const int row_off = gy * W * 3;
    // the box whose constants sit in registers (tiles of one row mostly meet the same single box)
    int cur = -1;
    float t0 = 0.f, t3 = 0.f, c0 = 0.f, c1 = 0.f, t6 = 0.f, c2 = 0.f, lo2 = 0.f, hi = 0.f;
    int org = 0, S = 0, bx0 = 0, bd = 0;
    const float4* bu = nullptr;
    uint8_t* brow = nullptr;
    bool proj_on = false;
    for (int xs = xlo & ~31; xs <= xhi; xs += 32) {
      const int gx = xs + lane;
      unsigned rest = __ballot_sync(0xffffffffu, rx0 <= rx1 && rx0 <= xs + 31 && rx1 >= xs);
      if (!rest) continue;
      unsigned found = 0;
      bool in_any = false;
      float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
      while (rest) {                                            // newest paste first (warp-uniform loop)
        const int i = 31 - __clz(rest);
        rest &= ~(1u << i);
        if (i != cur) {
          const BandBox& bx = sm.box[i];
          const float yf = (float)(gy - bx.y0);
          cur = i;
          t0 = bx.t0; t3 = bx.t3; c0 = bx.t1 * yf; c1 = bx.t4 * yf; c2 = bx.t7 * yf; t6 = bx.t6;
          proj_on = bx.t6 != 0.0f || bx.t7 != 0.0f;
          lo2 = bx.lo2; hi = bx.hi; org = bx.org; S = bx.S; bx0 = bx.x0; bd = bx.d;
          bu = bx.u;
          brow = bx.route + (gy - bx.y0) * bx.d;
        }
        const BandBox& bxs = sm.box[i];
        const int x = gx - bx0;
        const bool inwin = x >= 0 && x < bd;
        in_any = in_any || inwin;
        const float xf = (float)x;
        float ix = (t0 * xf + c0) + bxs.t2;
        float iy = (t3 * xf + c1) + bxs.t5;
        bool degenerate = false;
        if (proj_on) {
          const float proj = (t6 * xf + c2) + 1.0f;
          degenerate = proj == 0.0f;
          ix = ix / proj;
          iy = iy / proj;
        }
        const float x0f = floorf(ix), y0f = floorf(iy);
        // at least one tap inside the core <=> floor coordinate in [pad_lo - 1, pad_lo + ps - 1] on both axes
        const bool core = x0f > lo2 && x0f < hi && y0f > lo2 && y0f < hi;
        const bool take = inwin && found != 7u && core && !degenerate;
        if (!__any_sync(0xffffffffu, take)) continue;
        const float wx1 = (x0f + 1.0f) - ix, wx0 = ix - x0f, wy1 = (y0f + 1.0f) - iy, wy0 = iy - y0f;
        const int xi = (int)fminf(fmaxf(x0f, lo2), hi) - org;
        const int yi = (int)fminf(fmaxf(y0f, lo2), hi) - org;
        const float4* p = bu + (yi * S + xi);
        float R[3];
        blend3(p[0], p[1], p[S], p[S + 1], wx1, wx0, wy1, wy0, R);
        if (take) {
          unsigned bits = 0;
          if (!(found & 1u) && !(R[0] < -1.0f)) { v0 = R[0]; found |= 1u; bits |= (unsigned)(R[0] <= 1.0f); }
          if (!(found & 2u) && !(R[1] < -1.0f)) { v1 = R[1]; found |= 2u; bits |= (unsigned)(R[1] <= 1.0f) << 1; }
          if (!(found & 4u) && !(R[2] < -1.0f)) { v2 = R[2]; found |= 4u; bits |= (unsigned)(R[2] <= 1.0f) << 2; }
          if (bits) brow[x] = (uint8_t)bits;
        }
      }
      if (found || (in_any && all_px)) {
        const int e = row_off + gx * 3;
        float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
        if (found != 7u || m_img) { b0 = __ldg(img + e); b1 = __ldg(img + e + 1); b2 = __ldg(img + e + 2); }
        const float o0 = clampf((found & 1u) ? v0 : b0, -1.0f, 1.0f);
        const float o1 = clampf((found & 2u) ? v1 : b1, -1.0f, 1.0f);
        const float o2 = clampf((found & 4u) ? v2 : b2, -1.0f, 1.0f);
        o_img[e] = o0; o_img[e + 1] = o1; o_img[e + 2] = o2;
        if (m_img) {                                            // Masker: mask = original - pasted (attack_detection.py:429-430)
          m_img[e] = b0 - o0; m_img[e + 1] = b1 - o1; m_img[e + 2] = b2 - o2;
        }
      }
    }
  }
}

// Bands are handed out by an atomic ticket (their cost varies from nothing to several overlapping windows).
#ifndef EOT_COMP_MINB
#define EOT_COMP_MINB (1024 / EOT_COMP_THREADS)
#endif
__global__ void __launch_bounds__(kCompThreads, EOT_COMP_MINB) k_composite(EotShape s, Layout L, char* ws,
                                                        const float* __restrict__ images, float* out, float* mask,
                                                        const int32_t* __restrict__ offsets, int b0, int b1, int group) {
  __shared__ CompositeSmem sm;
  __shared__ int s_it;
  const int bands = (s.height + kCompRows - 1) / kCompRows;
  const int total = (b1 - b0) * bands;
  int* ticket = reinterpret_cast<int*>(ws + L.off_tickets) + group;
  for (;;) {
    if (threadIdx.x == 0) s_it = atomicAdd(ticket, 1);
    __syncthreads();
    const int it = s_it;
    if (it >= total) break;
    const int b = b0 + it / bands;
    composite_band(s, L, ws, images, out, mask, b, it - (it / bands) * bands, offsets, sm);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------------
static int check_shape(const EotShape* s) {
  if (!s) { set_error("shape is NULL"); return EOT_ERR_NULL_POINTER; }
  if (s->batch <= 0 || s->height <= 0 || s->width <= 0 || s->patch_size <= 0 || s->total_boxes < 0 ||
      (s->num_patches != 1 && s->num_patches != s->batch)) {
    set_error("bad shape: batch=%d H=%d W=%d P=%d num_patches=%d N=%d", s->batch, s->height, s->width,
              s->patch_size, s->num_patches, s->total_boxes);
    return EOT_ERR_BAD_SHAPE;
  }
  if ((int64_t)s->height * s->width * 3 >= (int64_t)1 << 31) { set_error("image too large for int32 indexing"); return EOT_ERR_BAD_SHAPE; }
  return EOT_OK;
}

static EotShape normalised(const EotShape& in) {
  EotShape s = in;
  if (s.patch_stride_x == 0) s.patch_stride_x = 3;
  if (s.patch_stride_y == 0) s.patch_stride_y = (int64_t)s.patch_size * 3;
  if (s.patch_stride_n == 0) s.patch_stride_n = (int64_t)s.patch_size * s.patch_size * 3;
  return s;
}

}  // namespace eot

using namespace eot;

extern "C" int eot_workspace_bytes(const EotShape* shape, size_t* bytes) {
  if (int rc = check_shape(shape)) return rc;
  if (!bytes) { set_error("bytes is NULL"); return EOT_ERR_NULL_POINTER; }
  *bytes = make_layout(*shape).total;
  return EOT_OK;
}

extern "C" int eot_box_geometry(const EotShape* shape, const float* boxes, const int32_t* box_offsets,
                                const EotBoxParams* params, const float* scale, EotBoxGeometry* geometry_out,
                                void* stream) {
  if (int rc = check_shape(shape)) return rc;
  if (!box_offsets || !scale || !geometry_out || (shape->total_boxes > 0 && (!boxes || !params))) {
    set_error("eot_box_geometry: NULL pointer");
    return EOT_ERR_NULL_POINTER;
  }
  const EotShape s = normalised(*shape);
  if (s.total_boxes > 0) {
    k_geometry_only<<<s.total_boxes, kThreads, 0, (cudaStream_t)stream>>>(s, make_layout(s), boxes, box_offsets, params, scale, geometry_out);
    count_launches(1);
  }
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

namespace eot {

// Enqueues the whole forward on `st`: memset of the small accumulators, the pre-pass (geometry + patch statistics +
// image pass in one launch), then match -> resize -> composite.
static int launch_forward(const EotShape& s, const Layout& L, const float* patch, const float* scale, const float* images,
                          const float* boxes, const int32_t* box_offsets, const EotBoxParams* params,
                          const float* print_wb, float* out_images, float* mask, char* ws, cudaStream_t st) {
  const int B = s.batch, P = s.patch_size, HW = s.height * s.width, N = s.total_boxes;
  EOT_CHECK_CUDA(cudaMemsetAsync(ws + L.off_ysum_img, 0, L.off_plans - L.off_ysum_img, st));
  const int pchunks = max(1, min((P * P + kThreads * 4 - 1) / (kThreads * 4), 64));
  const int cpi = (HW + kPassPixPerBlock - 1) / kPassPixPerBlock;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)images | (uintptr_t)out_images | (uintptr_t)(mask ? mask : out_images)) & 15) == 0;
  const int nsm = sm_count();
  const long long nblocks = (long long)N + (long long)B * pchunks + (long long)B * cpi;
  if (nblocks >= (1ll << 31)) { set_error("eot_apply_fwd: grid too large"); return EOT_ERR_BAD_SHAPE; }
  if (vec)
    k_prepass<true><<<(unsigned)nblocks, kThreads, 0, st>>>(s, L, patch, print_wb, boxes, box_offsets, params, scale, images,
                                                            out_images, mask, ws, N, B, pchunks, cpi, 0);
  else
    k_prepass<false><<<(unsigned)nblocks, kThreads, 0, st>>>(s, L, patch, print_wb, boxes, box_offsets, params, scale, images,
                                                             out_images, mask, ws, N, B, pchunks, cpi, 0);
  count_launches(1);
  if (N > 0) {
    const size_t smem = resize_smem_bytes(s, L);
    if (smem > 32 * 1024) EOT_CHECK_CUDA(cudaFuncSetAttribute(k_resize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // static + dynamic may pass 48 KB
    k_match<<<dim3(pchunks, B), kThreads, 0, st>>>(s, L, patch, print_wb, ws, 0);
    k_resize<<<nsm * EOT_RESIZE_MINB, EOT_RESIZE_THREADS, smem, st>>>(s, L, ws, box_offsets, 0, B);
    k_composite<<<nsm * EOT_COMP_MINB, kCompThreads, 0, st>>>(s, L, ws, images, out_images, mask, box_offsets, 0, B, 0);
    count_launches(3);
  }
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

static int forward_entry(const EotShape* shape, const float* patch, const float* scale, const float* images,
                         const float* boxes, const int32_t* box_offsets, const EotBoxParams* params, const float* print_wb,
                         float* out_images, float* out_masks, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  if (!patch || !scale || !images || !box_offsets || !print_wb || !out_images || !workspace ||
      (shape->total_boxes > 0 && (!boxes || !params))) {
    set_error("eot_apply_fwd: NULL pointer");
    return EOT_ERR_NULL_POINTER;
  }
  const bool want_mask = (shape->flags & EOT_FLAG_MASK_OUTPUT) != 0;
  if (want_mask && !out_masks) { set_error("EOT_FLAG_MASK_OUTPUT set but out_masks is NULL"); return EOT_ERR_NULL_POINTER; }
  if (want_mask && out_images == images) { set_error("the Masker's mask needs the original image: out_images may not alias images"); return EOT_ERR_BAD_SHAPE; }
  const EotShape s = normalised(*shape);
  const Layout L = make_layout(s);
  if (workspace_bytes < L.total) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, L.total);
    return EOT_ERR_WORKSPACE_TOO_SMALL;
  }
  if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return EOT_ERR_MISALIGNED; }
  return launch_forward(s, L, patch, scale, images, boxes, box_offsets, params, print_wb, out_images,
                        want_mask ? out_masks : nullptr, static_cast<char*>(workspace), (cudaStream_t)stream);
}

}  // namespace eot

extern "C" int eot_apply_fwd(const EotShape* shape, const float* patch, const float* scale, const float* images,
                             const float* boxes, const int32_t* box_offsets, const EotBoxParams* params,
                             const float* print_wb, float* out_images, float* out_masks, void* workspace,
                             size_t workspace_bytes, void* stream) {
  return forward_entry(shape, patch, scale, images, boxes, box_offsets, params, print_wb, out_images, out_masks, workspace,
                       workspace_bytes, stream);
}

extern "C" int eot_check_workspace(const EotShape* shape, const void* workspace, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  if (!workspace) { set_error("workspace is NULL"); return EOT_ERR_NULL_POINTER; }
  const Layout L = make_layout(normalised(*shape));
  int flag = 0;
  EOT_CHECK_CUDA(cudaMemcpyAsync(&flag, static_cast<const char*>(workspace) + L.off_counters + 2 * sizeof(int), sizeof(int),
                                 cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  EOT_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (flag) { set_error("a patch window did not fit the image (the reference would fail in tf.pad / scatter)"); return EOT_ERR_GEOMETRY; }
  return EOT_OK;
}

extern "C" int eot_brightness_match(const float* src, int64_t src_pixels, const float* tgt, int64_t tgt_pixels,
                                    float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!src || !tgt || !out || !workspace) { set_error("eot_brightness_match: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  if (src_pixels <= 0 || tgt_pixels <= 0) { set_error("eot_brightness_match: empty image"); return EOT_ERR_BAD_SHAPE; }
  if (workspace_bytes < 16 || ((uintptr_t)workspace & 7)) { set_error("eot_brightness_match: workspace needs 16 aligned bytes"); return EOT_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = static_cast<double*>(workspace);
  EOT_CHECK_CUDA(cudaMemsetAsync(sums, 0, 16, st));
  const int cap = sm_count() * 8;
  const int gs = (int)((src_pixels + kThreads - 1) / kThreads < cap ? (src_pixels + kThreads - 1) / kThreads : cap);
  const int gt = (int)((tgt_pixels + kThreads - 1) / kThreads < cap ? (tgt_pixels + kThreads - 1) / kThreads : cap);
  k_bm_ysum<<<gs, kThreads, 0, st>>>(src, src_pixels, sums);
  k_bm_ysum<<<gt, kThreads, 0, st>>>(tgt, tgt_pixels, sums + 1);
  k_bm_apply<<<gs, kThreads, 0, st>>>(src, src_pixels, tgt_pixels, sums, out);
  count_launches(3);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
