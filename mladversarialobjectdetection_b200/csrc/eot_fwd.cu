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
//   k_resize2         (eot_resize.cu) antialiased triangle resize + noise + brightness delta -> u_j (RGBX texels)
//   k_composite3      (eot_composite.cu) projective bilinear sample of the ring-padded u_j, `< -1` mask,
//                     sequential-paste resolution, clip, store, route bytes
#include "eot_common.cuh"
#include "eot_composite.cuh"
#include "eot_resize.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

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
      if (err_flag) atomicOr(err_flag, 1);
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
  pl.ia0 = fabsf(pl.T[0]) < 1e-6f ? 0.0f : 1.0f / pl.T[0];
  pl.ia3 = fabsf(pl.T[3]) < 1e-6f ? 0.0f : 1.0f / pl.T[3];
  pl.two_tap = 0;
  pl.rsv = 0;
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
// w_out: tap k at w_out[k * stride] (tap-major table: the columns pass reads it coalesced, the rows pass uniformly)
__device__ inline void span_row(int o, const SpanCfg& c, int in_size, int* start_out, float* __restrict__ w_out, int stride) {
  const float col_f = (float)o + 0.5f;
  const float sample_f = col_f * c.inv_scale + (-c.inv_scale * 0.0f);
  if (sample_f < 0.0f || sample_f > (float)in_size) {
    *start_out = 0;
    for (int k = 0; k < c.span; ++k) w_out[(size_t)k * stride] = 0.0f;
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
    w_out[(size_t)k * stride] = (k < n && ok) ? tri_weight((int)s0 + k, sample_f, c.one_over) * inv_total : 0.0f;
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
// Under a saturated memory system every dependent global round trip of this role costs ~1.5 us and the image pass cannot
// end before the role does, so the chain is kept short: the image of the box is found by all threads at once (no binary
// search over the CSR), and the span starts / tap weights the later steps read back live in shared memory (`scratch`,
// `scratch_bytes`; boxes whose tables do not fit read them back from global memory).
__device__ int geometry_block(const EotShape& s, const Layout& L, int j, const float* __restrict__ boxes,
                               const int32_t* __restrict__ offsets, const EotBoxParams* __restrict__ params,
                               const float* __restrict__ scale, char* ws, EotBoxGeometry* geom_out, void* scratch,
                               int scratch_bytes) {
  __shared__ BoxPlan spl;
  __shared__ int s_maxcnt;
  __shared__ int s_img[3];                                        // image of box j, its CSR range
  int* counters = ws ? reinterpret_cast<int*>(ws + L.off_counters) : nullptr;
  if (threadIdx.x == 0) { s_maxcnt = 0; s_img[0] = -1; s_img[1] = 0; s_img[2] = 0; }
  __syncthreads();
  for (int b = threadIdx.x; b < s.batch; b += blockDim.x) {       // image of box j: offsets[b] <= j < offsets[b + 1]
    const int lo = offsets[b], hi = offsets[b + 1];
    if (lo <= j && j < hi) { s_img[0] = b; s_img[1] = lo; s_img[2] = hi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // total_boxes is a CAPACITY: the boxes in use are [0, box_offsets[B]) (read here, on the device, so that a caller
    // whose box count comes out of a kernel -- the first pass' NMS -- needs no host read); the slots past them are
    // invalid boxes without work.  More boxes than the capacity: flagged (eot_check_workspace), the surplus is dropped.
    const int n_used = offsets[s.batch];
    if (n_used > s.total_boxes && counters) atomicOr(counters + 2, 8);
    const bool used = j < n_used && s_img[0] >= 0;
    const int a = used ? s_img[0] : s.batch - 1;
    const int first = used ? s_img[1] : offsets[a], last = used ? s_img[2] : offsets[a + 1];
    EotBoxParams prm = {};
    float bx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (used) {
      prm = params[j];
      for (int k = 0; k < 4; ++k) bx[k] = boxes[(size_t)j * 4 + k];
    }
    BoxPlan pl = make_plan(bx, *scale, prm, s, a, first, min(last, s.total_boxes), (int64_t)j * L.slot,
                           used && counters ? counters + 2 : nullptr);
    if (!used) pl.valid = 0;
    if (pl.valid) pl.span = span_cfg(pl.ps, s.patch_size).span;
    spl = pl;
    if (ws) {
      if (pl.valid)   // window work of the image (the backward hands out the heaviest images first)
        atomicAdd(reinterpret_cast<unsigned*>(ws + L.off_cost) + pl.image, min((unsigned)pl.ps * (unsigned)pl.ps, 1u << 22));
      reinterpret_cast<BoxPlan*>(ws + L.off_plans)[j] = pl;
      reinterpret_cast<int4*>(ws + L.off_cnt)[j] =
          pl.valid ? make_int4(fwd_strips(pl.ps, L.resize_rows), 0, (pl.ps + L.rb - 1) / L.rb, (pl.d + L.cr - 1) / L.cr)
                   : make_int4(0, 0, 0, 0);
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
  // shared-memory copies of the starts and of the tap-major weights, when they fit
  const bool staged = scratch && (size_t)ps * (cfg.span + 1) * 4 <= (size_t)scratch_bytes;
  int* s_starts = reinterpret_cast<int*>(scratch);
  float* s_w = reinterpret_cast<float*>(scratch) + ps;
  // Two-tap form: with span 3 (up-sampling / unit scale) the triangle kernel has at most two adjacent non-zero taps per
  // output index; when that holds for every index of the box (checked, not assumed) the resize reads (a, b, wa, wb).
  // Dropping a tap of weight 0 drops an addition of +-0: the sums are the oracle's.
  float4* tab2 = reinterpret_cast<float4*>(ws + L.off_tab2) + (size_t)j * L.lmin;
  bool two = cfg.span == 3;
  for (int o = threadIdx.x; o < ps; o += blockDim.x) {
    int st;
    if (staged) {
      span_row(o, cfg, P, &st, s_w + o, ps);
      s_starts[o] = st;
      starts[o] = st;
      for (int k = 0; k < cfg.span; ++k) weights[(size_t)k * ps + o] = s_w[(size_t)k * ps + o];
    } else {
      span_row(o, cfg, P, &st, weights + o, ps);
      starts[o] = st;
    }
    if (cfg.span == 3) {
      const float* wsrc = staged ? s_w : weights;
      const float w0 = wsrc[o], w1 = wsrc[ps + o], w2 = wsrc[2 * ps + o];
      int ia; float wa, wb;
      if (w2 == 0.0f) { ia = st; wa = w0; wb = w1; }
      else if (w0 == 0.0f) { ia = st + 1; wa = w1; wb = w2; }
      else { ia = st; wa = w0; wb = w1; two = false; }
      const int ib = min(ia + 1, P - 1);
      ia = min(ia, P - 1);
      tab2[o] = make_float4(__int_as_float(ia), __int_as_float(ib), wa, wb);
    }
  }
  const int all_two = __syncthreads_and(two ? 1 : 0);   // (also: starts[] of this box are complete)
  if (threadIdx.x == 0) reinterpret_cast<BoxPlan*>(ws + L.off_plans)[j].two_tap = all_two;
  const int* st_src = staged ? s_starts : starts;
  const float* w_src = staged ? s_w : weights;
  int2* inv = reinterpret_cast<int2*>(ws + L.off_inv) + (size_t)j * P;
  float* wt = reinterpret_cast<float*>(ws + L.off_wt) + (size_t)j * P * L.tcap;
  int2* stt = reinterpret_cast<int2*>(ws + L.off_stt) + (size_t)j * (P + 1);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const int2 rng = inverse_span_search(st_src, ps, cfg.span, i);
    inv[i] = rng;
    int cnt = rng.y - rng.x + 1;
    if (cnt < 0) cnt = 0;
    if (cnt > L.tcap) { cnt = L.tcap; if (counters) atomicOr(counters + 2, 4); }
    stt[i] = make_int2(min(max(rng.x, 0), ps - 1), cnt);
    atomicMax(&s_maxcnt, cnt);
  }
  __syncthreads();
  // transposed weights (ScaleAndTranslateGrad = scatter of the same weights): row i of W^T, dense from rng.x, row
  // stride = the largest tap count of this box
  const int tstride = max(s_maxcnt, 1);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const int2 rng = inverse_span_search(st_src, ps, cfg.span, i);    // (recomputed: cheaper than the global read-back)
    int cnt = rng.y - rng.x + 1;
    cnt = cnt < 0 ? 0 : (cnt > L.tcap ? L.tcap : cnt);
    for (int k0 = 0; k0 < tstride; k0 += 4) {                     // the loads of four taps travel together
      float w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + q, o = rng.x + k;
        w[q] = 0.0f;
        if (k < cnt) {
          const int kk = i - st_src[o];
          if (kk >= 0 && kk < cfg.span) w[q] = w_src[(size_t)kk * ps + o];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k0 + q < tstride) wt[(size_t)i * tstride + k0 + q] = w[q];
    }
  }
  if (threadIdx.x == 0) {
    const int rows = bwd_strip_rows(ps);
    stt[P] = make_int2(tstride, rows);
    reinterpret_cast<int4*>(ws + L.off_cnt)[j].y = (P + rows - 1) / rows;   // backward resize strips of this box
  }
  // per window row: the columns whose sample can touch the core (the composite's segment list is cut from these)
  int2* rowtab = reinterpret_cast<int2*>(ws + L.off_rowtab) + (size_t)j * (s.height < s.width ? s.height : s.width);
  for (int wy = threadIdx.x; wy < spl.d; wy += blockDim.x) {
    int xa, xb;
    row_core_range(spl, wy, &xa, &xb);
    rowtab[wy] = make_int2(xa, xb);
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

constexpr int kMaxOrderedImages = 2048;   // batches beyond this keep the identity order (the ranking is quadratic in one CTA)
constexpr int kSmallRolesScratch = kThreads * 16 + kMaxOrderedImages * 4;   // prefix-sum partials + image costs
constexpr int kMaxSortedBoxes = 1024;                              // key and two counts per box in the same scratch (12 KB)
static_assert(kMaxSortedBoxes * 12 <= kSmallRolesScratch && kMaxSortedBoxes % kThreads == 0, "sorted item lists: scratch");

// Exclusive prefix sums of the per-box work-item counts (one CTA; N is a few hundred to a few thousand).
__device__ __forceinline__ int4 add4(int4 a, int4 b) { return make_int4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// Every count is loaded once, all loads of a thread together (the block runs while the image pass saturates the memory
// system: each dependent round trip costs ~1.5 us); with `items` / `citems` the forward's item lists are written in box
// order from the registers that hold counts and positions.
constexpr int kScanPer = 4;                                        // boxes per thread held in registers (N <= 1024)
constexpr int kMaxListedBoxes = 1024;                              // boxes whose list positions fit the shared-memory scratch
__device__ __forceinline__ int pack_counts(int nz, int nw) { return (nz & 0xffff) | (nw << 16); }
// s_pos / s_n (may alias `part`): per box, where its resize / composite items start in the lists and how many there are
// (box order), for write_item_lists; only filled when N <= kMaxListedBoxes.
__device__ void scan_block(int N, const int4* cnt, int4* base, int4* part /* [blockDim.x] shared */, int2* s_pos, int* s_n) {
  const int T = blockDim.x;
  const int per = (N + T - 1) / T;
  const int j0 = threadIdx.x * per, j1 = min(N, j0 + per);
  const bool regs = per <= kScanPer;
  int4 c[kScanPer];
  int4 sum = make_int4(0, 0, 0, 0);
  if (regs) {
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) c[q] = (j0 + q < j1) ? __ldcg(cnt + j0 + q) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) sum = add4(sum, c[q]);
  } else {
    for (int j = j0; j < j1; ++j) sum = add4(sum, __ldcg(cnt + j));
  }
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < T; d <<= 1) {
    int4 v = make_int4(0, 0, 0, 0);
    if ((int)threadIdx.x >= d) v = part[threadIdx.x - d];
    __syncthreads();
    part[threadIdx.x] = add4(part[threadIdx.x], v);
    __syncthreads();
  }
  int4 run = threadIdx.x ? part[threadIdx.x - 1] : make_int4(0, 0, 0, 0);
  const int4 total = part[T - 1];
  __syncthreads();                                                // (s_pos / s_n may overlay the partial sums)
  if (regs) {
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) {
      const int j = j0 + q;
      if (j < j1) {
        base[j] = run;
        if (s_pos) { s_pos[j] = make_int2(run.z, run.w); s_n[j] = pack_counts(c[q].z, c[q].w); }
        run = add4(run, c[q]);
      }
    }
  } else {
    for (int j = j0; j < j1; ++j) {
      base[j] = run;
      run = add4(run, __ldcg(cnt + j));
    }
  }
  if ((int)threadIdx.x == T - 1) base[N] = total;
  __syncthreads();
}

// The forward's item lists from the per-box positions: one warp per box, lanes write consecutive items (one thread per
// box would issue ~40 k scattered 8-byte stores from a single SM: ~20 us of its store pipe while the image pass waits).
__device__ void write_item_lists(int N, const int2* s_pos, const int* s_n, int2* items, int2* citems) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < N; j += nwarps) {
    const int2 at = s_pos[j];
    const int n = s_n[j], nz = n & 0xffff, nw = (unsigned)n >> 16;
    for (int i = lane; i < nz; i += 32) items[at.x + i] = make_int2(j, i);
    for (int i = lane; i < nw; i += 32) citems[at.y + i] = make_int2(j, i);
  }
}

__global__ void __launch_bounds__(kThreads) k_geometry_only(EotShape s, Layout L, const float* __restrict__ boxes,
                                                            const int32_t* __restrict__ offsets,
                                                            const EotBoxParams* __restrict__ params,
                                                            const float* __restrict__ scale, EotBoxGeometry* geom_out) {
  geometry_block(s, L, blockIdx.x, boxes, offsets, params, scale, nullptr, geom_out, nullptr, 0);
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
#define EOT_PASS_PIX 8
#endif
#ifndef EOT_PASS_STREAM
#define EOT_PASS_STREAM 0
#endif
#if EOT_PASS_STREAM >= 1
#define EOT_PASS_LD(p) __ldcs(p)
#else
#define EOT_PASS_LD(p) __ldg(p)
#endif
#if EOT_PASS_STREAM >= 2
#define EOT_PASS_ST(p, v) __stcs(p, v)
#else
#define EOT_PASS_ST(p, v) (*(p) = (v))
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
          v[k][0] = EOT_PASS_LD(in4 + q);
          v[k][1] = EOT_PASS_LD(in4 + q + 1);
          v[k][2] = EOT_PASS_LD(in4 + q + 2);
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int pix = pix0 + ((half * 2 + k) * kThreads + threadIdx.x) * 4;
        if (pix < HW) {
          const int q = (pix >> 2) * 3;
          const float4 a = v[k][0], bb = v[k][1], c = v[k][2];
          if (o) { EOT_PASS_ST(o4 + q, a); EOT_PASS_ST(o4 + q + 1, bb); EOT_PASS_ST(o4 + q + 2, c); }
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

// The two small roles of the pre-pass (by block index): [0, n_geom) geometry of box j (the last one to finish also
// builds the prefix sums, the item lists and the image order), then n_stat_imgs * pchunks blocks of patch luma
// statistics.  0: neither; 1: this block had one of them; 2: ... and was the geometry block that built the tables.
__device__ int small_roles(const EotShape& s, const Layout& L, int blk, const float* __restrict__ patch,
                            const float* __restrict__ print_wb, const float* __restrict__ boxes,
                            const int32_t* __restrict__ offsets, const EotBoxParams* __restrict__ params,
                            const float* __restrict__ scale, char* ws, int n_geom, int n_stat_imgs, int pchunks, double* red,
                            void* scratch /* kSmallRolesScratch bytes of shared memory, 16-byte aligned */, int sort_items) {
  if (blk < n_geom) {
    __shared__ int s_last;
    int4* s_part = reinterpret_cast<int4*>(scratch);               // [kThreads]
    unsigned* s_cost = reinterpret_cast<unsigned*>(s_part + kThreads);   // [kMaxOrderedImages]
    geometry_block(s, L, blk, boxes, offsets, params, scale, ws, nullptr, scratch, kSmallRolesScratch);
    __threadfence();
    __syncthreads();
    int* counters = reinterpret_cast<int*>(ws + L.off_counters);
    if (threadIdx.x == 0) s_last = (atomicAdd(counters + 5, 1) == n_geom - 1);
    __syncthreads();

    if (s_last) {                                        // last geometry block: prefix sums of the work-item counts
      __threadfence();
      const int4* cnt = reinterpret_cast<const int4*>(ws + L.off_cnt);
      int4* base = reinterpret_cast<int4*>(ws + L.off_base);
      int2* items = reinterpret_cast<int2*>(ws + L.off_items);    // forward resize / composite items in ticket order
      int2* citems = reinterpret_cast<int2*>(ws + L.off_citems);
      const bool sorted = sort_items && n_geom <= kMaxSortedBoxes && n_geom <= kMaxListedBoxes;
      // images by decreasing window work (ties by index): the backward's per-image items are handed out heaviest first.
      // The costs were added up by the geometry blocks; their loads travel with the scan's.
      const int B = s.batch;
      int* order = reinterpret_cast<int*>(ws + L.off_order);
      const unsigned* cost = reinterpret_cast<const unsigned*>(ws + L.off_cost);
      const bool listed = n_geom <= kMaxListedBoxes;              // positions and counts of every box fit the scratch
      int2* s_pos = reinterpret_cast<int2*>(scratch);
      int* s_n = reinterpret_cast<int*>(s_pos + kMaxListedBoxes);
      unsigned my_cost[kMaxOrderedImages / kThreads];             // (the scan's scratch overlays s_cost)
#pragma unroll
      for (int q = 0; q < kMaxOrderedImages / kThreads; ++q) {
        const int b = threadIdx.x + q * kThreads;
        my_cost[q] = (b < B && B <= kMaxOrderedImages) ? __ldcg(cost + b) : 0u;
      }
      scan_block(n_geom, cnt, base, s_part, listed && !sorted ? s_pos : nullptr, s_n);
      if (listed && !sorted) {
        write_item_lists(n_geom, s_pos, s_n, items, citems);
        __syncthreads();
      } else if (!listed) {                                       // very many boxes: one thread per box, box order
        for (int j = threadIdx.x; j < n_geom; j += blockDim.x) {
          const int4 n = __ldcg(cnt + j), at = base[j];
          for (int i = 0; i < n.z; ++i) items[at.z + i] = make_int2(j, i);
          for (int i = 0; i < n.w; ++i) citems[at.w + i] = make_int2(j, i);
        }
        __syncthreads();
      }
#pragma unroll
      for (int q = 0; q < kMaxOrderedImages / kThreads; ++q) {
        const int b = threadIdx.x + q * kThreads;
        if (b < B && B <= kMaxOrderedImages) s_cost[b] = my_cost[q];
      }
      __syncthreads();
      if (B <= kMaxOrderedImages) {
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
          const unsigned mine = s_cost[b];
          int rank = 0;
          for (int o2 = 0; o2 < B; ++o2) rank += (s_cost[o2] > mine || (s_cost[o2] == mine && o2 < b)) ? 1 : 0;
          order[rank] = b;
        }
      } else {
        for (int b = threadIdx.x; b < B; b += blockDim.x) order[b] = b;
      }
      if (sorted) {
        // Largest boxes first (item cost grows with the patch side): the persistent kernels end on the cheapest items, so
        // the tail in which warps run dry is as short as an item can be.  Only valid when one call takes the whole list
        // (image groups and the fused kernel address the list by image and keep box order).
        int* s_key = reinterpret_cast<int*>(scratch);              // [N] resize items of box j (proportional to its side)
        int2* s_cnt = reinterpret_cast<int2*>(s_key + ((n_geom + 1) & ~1));   // [N] (resize, composite) counts in rank order -> prefix
        constexpr int kPer = kMaxSortedBoxes / kThreads;           // boxes per thread, counts and ranks in registers
        int4 nq[kPer];
        int rq[kPer];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int j = threadIdx.x + q * kThreads;
          nq[q] = j < n_geom ? __ldcg(cnt + j) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int j = threadIdx.x + q * kThreads;
          if (j < n_geom) s_key[j] = nq[q].z;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int j = threadIdx.x + q * kThreads;
          rq[q] = 0;
          if (j < n_geom) {
            const int mine = nq[q].z;
            int rank = 0;
            for (int k = 0; k < n_geom; ++k) rank += (s_key[k] > mine || (s_key[k] == mine && k < j)) ? 1 : 0;
            rq[q] = rank;
            s_cnt[rank] = make_int2(nq[q].z, nq[q].w);
          }
        }
        __syncthreads();
        if (threadIdx.x < 32) {                                    // exclusive prefix over the ranks, one warp
          int2 run = make_int2(0, 0);
          for (int r0 = 0; r0 < n_geom; r0 += 32) {
            const int r = r0 + threadIdx.x;
            const int2 v = r < n_geom ? s_cnt[r] : make_int2(0, 0);
            int2 inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int ax = __shfl_up_sync(0xffffffffu, inc.x, d), ay = __shfl_up_sync(0xffffffffu, inc.y, d);
              if ((int)threadIdx.x >= d) { inc.x += ax; inc.y += ay; }
            }
            if (r < n_geom) s_cnt[r] = make_int2(run.x + inc.x - v.x, run.y + inc.y - v.y);
            run.x += __shfl_sync(0xffffffffu, inc.x, 31);
            run.y += __shfl_sync(0xffffffffu, inc.y, 31);
          }
        }
        __syncthreads();
        int2 at[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) at[q] = (threadIdx.x + q * kThreads < n_geom) ? s_cnt[rq[q]] : make_int2(0, 0);
        __syncthreads();                                          // (s_pos / s_n overlay the keys and the prefix)
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int j = threadIdx.x + q * kThreads;
          if (j < n_geom) { s_pos[j] = at[q]; s_n[j] = pack_counts(nq[q].z, nq[q].w); }
        }
        __syncthreads();
        write_item_lists(n_geom, s_pos, s_n, items, citems);
      }
    }
    return s_last ? 2 : 1;
  }
  blk -= n_geom;
  if (blk < n_stat_imgs * pchunks) {
    const int b = blk / pchunks;
    patch_stats_block(s, b, blk - b * pchunks, pchunks, patch, print_wb, reinterpret_cast<double*>(ws + L.off_ysum_patch), red);
    if (threadIdx.x == 0) {                              // (thread 0 added the block's sum) finished statistics blocks
      __threadfence();
      atomicAdd(reinterpret_cast<int*>(ws + L.off_fused) + 2, 1);
    }
    if (blk == 0 && threadIdx.x == 0) {                  // CSR copy for the backward
      int32_t* off_copy = reinterpret_cast<int32_t*>(ws + L.off_offsets);
      for (int i = 0; i <= s.batch; ++i) off_copy[i] = min(offsets[i], s.total_boxes);
    }
    return 1;
  }
  return 0;
}

// One launch, three independent roles selected by the block index (they only meet at k_match):
//   [0, n_geom)                         geometry of box j                   (n_geom = N or 0)
//   [.., + n_stat_imgs*pchunks)         patch luma statistics               (n_stat_imgs = B or 0)
//   [.., + (b1-b0)*cpi)                 image pass (copy + luma sum) of images [b0,b1), the HBM-bound bulk
#ifndef EOT_PREPASS_MINB
#define EOT_PREPASS_MINB 6
#endif
template <bool kVec>
__global__ void __launch_bounds__(kThreads, EOT_PREPASS_MINB) k_prepass(EotShape s, Layout L, const float* __restrict__ patch,
                                                      const float* __restrict__ print_wb, const float* __restrict__ boxes,
                                                      const int32_t* __restrict__ offsets,
                                                      const EotBoxParams* __restrict__ params,
                                                      const float* __restrict__ scale, const float* __restrict__ images,
                                                      float* out, float* mask, char* ws, int n_geom, int n_stat_imgs,
                                                      int pchunks, int cpi, int b0, int sort_items) {
  __shared__ double red[32];
  __shared__ __align__(16) unsigned char s_scratch[kSmallRolesScratch];
  pdl_trigger();
  int blk = blockIdx.x;
  if (small_roles(s, L, blk, patch, print_wb, boxes, offsets, params, scale, ws, n_geom, n_stat_imgs, pchunks, red, s_scratch, sort_items)) return;
  blk -= n_geom;
  blk -= n_stat_imgs * pchunks;
  const int b = b0 + blk / cpi, chunk = blk % cpi;
  image_pass_block<kVec>(s.height * s.width, b, chunk, images, out, mask, reinterpret_cast<double*>(ws + L.off_ysum_img),
                         reinterpret_cast<int*>(ws + L.off_oor), red);
}

// ---- acquire / release plumbing between CTAs of one launch ----
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Warp-wide wait until *p >= v: lane 0 polls (acquire) with exponential back-off -- thousands of warps polling a handful of
// L2 lines at full rate starve the very atomics they wait for.  A wait that outlasts ~2 s (a lost signal: a bug, never
// load) raises error flag 16 and goes on, so that a fault shows as a failed call instead of a hung device.
__device__ __forceinline__ void wait_ge(const int* p, int v, int* err, int* spins, bool one_thread = false) {
  if (one_thread) {                                               // called by a single thread (no warp convergence)
    unsigned ns = 64, n = 0;
    while (ld_acquire(p) < v) {
      __nanosleep(ns);
      if (ns < 2048) ns *= 2;
      if (++n > (1u << 20)) { atomicOr(err, 16); break; }
    }
    return;
  }
  if ((threadIdx.x & 31) == 0 && ld_acquire(p) < v) {
    unsigned ns = 128, n = 0;
    while (ld_acquire(p) < v) {
      __nanosleep(ns);
      if (ns < 2048) ns *= 2;
      if (++n > (1u << 20)) { atomicOr(err, 16); break; }
    }
    if (spins) atomicAdd(spins, (int)n);
  }
  __syncwarp();
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void match_range(const EotShape& s, const Layout& L, const float* __restrict__ patch,
                                            const float* __restrict__ print_wb, char* ws, int b, int t0, int tstride);

// ------------------------------------------------------------------------------------------------
// Image pass as a bulk-copy pipeline (cp.async.bulk + mbarrier: the TMA engine moves the bytes, the threads only read
// the staged tile for the luma sum).  Persistent CTAs over contiguous ranges of 1024-pixel tiles (12 KB; 48 bytes per
// thread), kBulkStages buffers per CTA:
//   thread 0   posts the byte count on the stage's mbarrier and issues the global -> shared copy of a tile kBulkStages - 1
//              tiles ahead; after the CTA has read a tile it issues the shared -> global copy of the same buffer
//              (bulk group) and, before it refills a buffer, waits until the copy that last read it has done reading
//   all        wait on the stage's mbarrier, read their four pixels (3 x LDS.128), accumulate luma in float64 and the
//              out-of-range flag; per image one block tree reduction + one atomicAdd
// Used when out != images, no mask is written and the rows are 16-byte aligned; every other case takes k_prepass.
// ------------------------------------------------------------------------------------------------
#ifndef EOT_PREPASS_BULK
#define EOT_PREPASS_BULK 1
#endif
#ifndef EOT_BULK_QUADS
#define EOT_BULK_QUADS 1
#endif
#ifndef EOT_BULK_TILE_PIX
#define EOT_BULK_TILE_PIX (4 * EOT_BULK_QUADS * kThreads)          // 1024 pixels = 12288 bytes per quad
#endif
constexpr int kBulkTilePix = EOT_BULK_TILE_PIX;                    // a smaller tile leaves the last threads of a quad idle
static_assert(kBulkTilePix % 32 == 0 && kBulkTilePix <= 4 * EOT_BULK_QUADS * kThreads, "tile: 128-byte multiple, covered by the quads");
#ifndef EOT_BULK_STAGES
#define EOT_BULK_STAGES 2
#endif
#ifndef EOT_BULK_CTAS
#define EOT_BULK_CTAS 6
#endif
constexpr int kBulkStages = EOT_BULK_STAGES;
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "EOT_BULK_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra EOT_BULK_WAIT_%=;\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, EOT_BULK_CTAS) k_prepass_bulk(EotShape s, Layout L, const float* __restrict__ patch,
                                                              const float* __restrict__ print_wb, const float* __restrict__ boxes,
                                                              const int32_t* __restrict__ offsets,
                                                              const EotBoxParams* __restrict__ params,
                                                              const float* __restrict__ scale, const float* __restrict__ images,
                                                              float* out, char* ws, int n_geom, int n_stat_imgs, int pchunks,
                                                              int n_copy_ctas, int b0, int b1, int sort_items) {
  extern __shared__ __align__(128) unsigned char stage_mem[];     // kBulkStages x 12288 bytes
  __shared__ double red[32];
  __shared__ __align__(8) uint64_t s_full[kBulkStages];
  static_assert((size_t)EOT_BULK_STAGES * kBulkTilePix * 12 >= (size_t)kSmallRolesScratch, "the stage buffers double as the small roles' scratch");
  pdl_trigger();
  int blk = blockIdx.x;
  if (small_roles(s, L, blk, patch, print_wb, boxes, offsets, params, scale, ws, n_geom, n_stat_imgs, pchunks, red, stage_mem, sort_items)) return;
  blk -= n_geom + n_stat_imgs * pchunks;
  const int HW = s.height * s.width;
  const int tpi = (HW + kBulkTilePix - 1) / kBulkTilePix;         // tiles per image
  // CTA (image slot, part): `parts` CTAs sweep one image together, tile k of the image going to part k % parts, so an
  // image is read and written as one linear stream and every CTA sums the luma of ONE image at a time (block reduction
  // + one atomic per image, no contention); with more images than CTAs a CTA takes images slot, slot + slots, ...
  const int n_img = b1 - b0;
  const int parts = max(1, n_copy_ctas / n_img);
  const int slots = n_copy_ctas / parts;                          // image slots served concurrently
  const int slot = blk / parts, part = blk - slot * parts;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBulkStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tiles_mine = part < tpi ? (tpi - part + parts - 1) / parts : 0;      // per image
  const int imgs_mine = slot < slots ? (n_img - slot + slots - 1) / slots : 0;
  const int my_tiles = tiles_mine * imgs_mine;
  auto tile_img = [&](int it) { return b0 + slot + (it / tiles_mine) * slots; };
  auto tile_k = [&](int it) { return part + (it % tiles_mine) * parts; };
  auto tile_bytes = [&](int it) { return (uint32_t)(min(kBulkTilePix, HW - tile_k(it) * kBulkTilePix) * 12); };
  auto tile_off = [&](int it) { return ((size_t)tile_img(it) * HW + (size_t)tile_k(it) * kBulkTilePix) * 3; };
  if (threadIdx.x == 0)                                           // prologue: the first kBulkStages - 1 tiles
    for (int it = 0; it < my_tiles && it < kBulkStages - 1; ++it)
      bulk_load(stage_mem + (size_t)(it % kBulkStages) * kBulkTilePix * 12, images + tile_off(it), tile_bytes(it), &s_full[it % kBulkStages]);
  double* ysum_img = reinterpret_cast<double*>(ws + L.off_ysum_img);
  int* oor_flags = reinterpret_cast<int*>(ws + L.off_oor);
  double acc = 0.0;
  bool oor = false;
  for (int it = 0; it < my_tiles; ++it) {
    const int stg = it % kBulkStages;
    bulk_wait_parity(&s_full[stg], (uint32_t)((it / kBulkStages) & 1));
#pragma unroll
    for (int qd = 0; qd < EOT_BULK_QUADS; ++qd) {
      const int px = (qd * kThreads + threadIdx.x) * 4;           // first of this thread's four pixels inside the tile
      const float4* tp = reinterpret_cast<const float4*>(stage_mem + (size_t)stg * kBulkTilePix * 12) + (px >> 2) * 3;
      if (px < (int)(tile_bytes(it) / 12)) {
        const float4 a = tp[0], bb = tp[1], c = tp[2];
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
    __syncthreads();                                              // every thread has read the tile
    if (threadIdx.x == 0) {
      if (out) bulk_store(out + tile_off(it), stage_mem + (size_t)stg * kBulkTilePix * 12, tile_bytes(it));
      const int itn = it + kBulkStages - 1;                       // refill the buffer of the previous tile
      if (itn < my_tiles) {
        if (out) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // ... once its outgoing copy has done reading it
        const int sn = itn % kBulkStages;
        bulk_load(stage_mem + (size_t)sn * kBulkTilePix * 12, images + tile_off(itn), tile_bytes(itn), &s_full[sn]);
      }
    }
    if ((it + 1) % tiles_mine == 0) {                             // (uniform) last tile of this image for this CTA
      const int any = __syncthreads_or(oor ? 1 : 0);
      acc = block_sum(acc, red);
      const int b = tile_img(it);
      if (threadIdx.x == 0) {
        atomicAdd(ysum_img + b, acc);
        if (any) atomicOr(oor_flags + b, 1);
      }
      acc = 0.0; oor = false;
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// print adjust + brightness match of the patch for image b (attacker.py:372; brightness_matcher.py:43-73)
// ------------------------------------------------------------------------------------------------
// blocks [0, nb * pchunks): match the patch to image b0 + blk / pchunks (needs the finished pre-pass: mean luma)
__global__ void __launch_bounds__(kThreads) k_match(EotShape s, Layout L, const float* __restrict__ patch,
                                                    const float* __restrict__ print_wb, const int32_t* __restrict__ offsets,
                                                    char* ws, int b0, int pchunks) {
  pdl_wait();
  pdl_trigger();
  const int blk = blockIdx.x;
  const int b = b0 + blk / pchunks;
  // images the composite's common path does not take (values outside [-1,1], more than 32 boxes): tell its second kernel
  if (blk % pchunks == 0 && threadIdx.x == 0 &&
      (reinterpret_cast<const int*>(ws + L.off_oor)[b] != 0 || offsets[b + 1] - offsets[b] > 32))
    atomicOr(reinterpret_cast<int*>(ws + L.off_counters) + 6, 1);
  match_range(s, L, patch, print_wb, ws, b, (blk % pchunks) * blockDim.x + threadIdx.x, pchunks * blockDim.x);
}

// texels t0, t0 + tstride, ... of the matched patch of image b
__device__ __forceinline__ void match_range(const EotShape& s, const Layout& L, const float* __restrict__ patch,
                                            const float* __restrict__ print_wb, char* ws, int b, int t0, int tstride) {
  const int P = s.patch_size;
  const double* ysum_img = reinterpret_cast<const double*>(ws + L.off_ysum_img);
  const double* ysum_patch = reinterpret_cast<const double*>(ws + L.off_ysum_patch);
  const float mu_t = (float)(__ldcg(ysum_img + b) / (double)((size_t)s.height * s.width));
  const float mu_s = (float)(__ldcg(ysum_patch + b) / (double)((size_t)P * P));
  const float* wb = print_wb + (size_t)b * 6;
  const float* base = patch + (s.num_patches > 1 ? (int64_t)b * s.patch_stride_n : 0);
  float4* m = reinterpret_cast<float4*>(ws + L.off_match) + (size_t)b * P * P;
  float w6[6];                                                    // the image's print-adjust coefficients: loaded once
#pragma unroll
  for (int k = 0; k < 6; ++k) w6[k] = __ldg(wb + k);
  // four texels per round: their patch values are fetched together (the kernel is a chain of L2 round trips otherwise)
  for (int t4 = t0; t4 < P * P; t4 += 4 * tstride) {
    float v[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = min(t4 + q * tstride, P * P - 1);
      const int py = t / P, px = t - py * P;
      const float* p = base + (int64_t)py * s.patch_stride_y + (int64_t)px * s.patch_stride_x;
      v[q][0] = __ldg(p); v[q][1] = __ldg(p + 1); v[q][2] = __ldg(p + 2);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = t4 + q * tstride;
      if (t < P * P) {
        const TexelYuv y = texel_yuv(v[q][0], v[q][1], v[q][2], w6);
        const float yp = clampf((y.y - mu_s) + mu_t, 0.0f, 1.0f);
        // rgb = [Y',U,V] . K'  as ((Y'*K'0c + U*K'1c) + V*K'2c)
        const float r = (yp * 1.0f + y.u * EOT_I10) + y.v * EOT_I20;
        const float g = (yp * 1.0f + y.u * EOT_I11) + y.v * EOT_I21;
        const float bl = (yp * 1.0f + y.u * EOT_I12) + y.v * EOT_I22;
        m[t] = make_float4(clampf(r, 0.0f, 1.0f) * EOT_C255_127 - 1.0f, clampf(g, 0.0f, 1.0f) * EOT_C255_127 - 1.0f,
                           clampf(bl, 0.0f, 1.0f) * EOT_C255_127 - 1.0f, 0.0f);
      }
    }
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
// The whole forward as ONE persistent kernel (cooperative launch: every CTA resident).  The image pass is bandwidth
// bound and the window work (match, resize, composite) instruction bound; as separate kernels they run back to back
// (the window kernels hold every register of an SM, nothing co-resides) and pay five kernel boundaries.  Here
//   CTAs [0, n_copy)   stream the images through shared memory with the bulk-copy engine (as k_prepass_bulk, `slots`
//                      images at a time, in image order), then join the window work;
//   the other CTAs     first run the geometry / patch-statistics roles, the last geometry block also lays out the task
//                      queue, then every WARP draws tasks from one ordered queue.
// Task queue, by step k:  match parts of image k,  resize items of image k - skew,  composite items of image k - 2 skew.
// A task waits (acquire spin on a per-image counter) only for tasks EARLIER in the queue or for the copy CTAs, which
// wait for nobody: with every CTA resident and tickets drawn in order, the lowest unfinished task can always run.
//   match(b)      needs the image's luma sum (all tiles of b read) and the patch statistics
//   resize(b)     needs every match part of b
//   composite(b)  needs every resize item of b and the stored copy of image b (it overwrites pixels of it)
// Tail: once every image's composite items are done, all threads share the open-pixel list / general rows
// (composite_rest_body).
// ------------------------------------------------------------------------------------------------
#ifndef EOT_FUSED_STAGES
#define EOT_FUSED_STAGES 4
#endif
#ifndef EOT_FUSED_MINB
#define EOT_FUSED_MINB 4
#endif
constexpr int kFusedStages = EOT_FUSED_STAGES;
constexpr int kFusedTilePix = 4 * kThreads;                        // 1024 pixels = 12288 bytes
constexpr int kMatchTexels = 512;                                  // texels per match task (one warp)

struct FusedView {
  int* ctl;
  int* img;
  int* step_start;
  int4* step_info;
};
__device__ __forceinline__ FusedView fused_view(char* ws, const Layout& L, int B) {
  FusedView v;
  v.ctl = reinterpret_cast<int*>(ws + L.off_fused);
  v.img = v.ctl + 64;
  v.step_start = v.img + 8 * (size_t)B;
  v.step_info = reinterpret_cast<int4*>(v.step_start + fused_start_ints(B));
  return v;
}
// Completion of one task of a per-image counter; with the debug timeline on, the task that completes the image's stage
// (count `full`) stamps the time into `stamp`.
__device__ __forceinline__ void warp_signal(int* p, int lane, int full = 0, int* stamp = nullptr, unsigned long long t_base = 0) {
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    const int old = atomicAdd(p, 1);
    if (stamp && old == full - 1) *stamp = (int)(global_ns() - t_base);
  }
}

// Task queue layout, by the geometry block that finished last (all its threads).
__device__ void build_steps(const EotShape& s, const Layout& L, char* ws, const int32_t* __restrict__ offsets, int N, int skew,
                            int nM, int* scratch /* >= B + 2 skew ints of shared memory */) {
  const int B = s.batch, steps = B + 2 * skew;
  const FusedView fv = fused_view(ws, L, B);
  const int4* base = reinterpret_cast<const int4*>(ws + L.off_base);
  __syncthreads();
  for (int k = threadIdx.x; k < steps; k += blockDim.x) {
    const int m = k < B ? nM : 0;
    int nR = 0, rbase = 0, nC = 0, cbase = 0;
    const int bR = k - skew, bC = k - 2 * skew;
    if (bR >= 0 && bR < B) { rbase = base[min(offsets[bR], N)].z; nR = base[min(offsets[bR + 1], N)].z - rbase; }
    if (bC >= 0 && bC < B) { cbase = base[min(offsets[bC], N)].w; nC = base[min(offsets[bC + 1], N)].w - cbase; }
    fv.step_info[k] = make_int4(m, nR, rbase, cbase);
    scratch[k] = m + nR + nC;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, empty = 0;
    for (int k = 0; k < steps; ++k) {
      fv.step_start[k] = run;
      run += scratch[k];
      if (k >= 2 * skew && scratch[k] - fv.step_info[k].x - fv.step_info[k].y == 0) ++empty;   // image without composite items
    }
    fv.step_start[steps] = run;
    fv.ctl[1] = run;
    if (empty) atomicAdd(fv.ctl + 3, empty);
    __threadfence();
    atomicExch(fv.ctl, 1);
  }
}

__global__ void __launch_bounds__(kThreads, EOT_FUSED_MINB) k_forward_fused(
    EotShape s, Layout L, const float* __restrict__ patch, const float* __restrict__ print_wb, const float* __restrict__ boxes,
    const int32_t* __restrict__ offsets, const EotBoxParams* __restrict__ params, const float* __restrict__ scale,
    const float* __restrict__ images, float* out, char* ws, int n_geom, int pchunks, int n_copy, int slots, int skew, int nM,
    float one, int prefetch, int dbg) {
  extern __shared__ __align__(128) unsigned char fsmem[];         // copy role: stage buffers; window work: per-warp scratch
  __shared__ double red[32];
  __shared__ __align__(8) uint64_t s_full[kFusedStages];
  const int B = s.batch, HW = s.height * s.width;
  const FusedView fv = fused_view(ws, L, B);
  int* counters = reinterpret_cast<int*>(ws + L.off_counters);
  int* err = counters + 2;
  unsigned long long t_base = 0;
  if (dbg && threadIdx.x == 0) {                                   // debug timeline (EOT_KERNEL_TIMES=1): ns since the first CTA started
    const unsigned long long now = global_ns();
    const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(fv.ctl + 8), 0ull, now);
    t_base = old ? old : now;
  }
  if (dbg) {                                                       // (every lane 0 stamps)
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_base = t_base;
    __syncthreads();
    t_base = s_base;
  }
  const int tpi = (HW + kFusedTilePix - 1) / kFusedTilePix;       // tiles per image
  const int parts = n_copy / slots;                               // CTAs sweeping one image together
  const int pass_expected = min(parts, tpi);                      // CTAs that own tiles of an image

  if ((int)blockIdx.x < n_copy) {
    // ---- copy role --------------------------------------------------------------------------------------------------
    const int blk = blockIdx.x;
    const int slot = blk / parts, part = blk - slot * parts;
    if (threadIdx.x == 0) {
      for (int i = 0; i < kFusedStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_full[i])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tiles_mine = part < tpi ? (tpi - part + parts - 1) / parts : 0;      // per image
    const int imgs_mine = (B - slot + slots - 1) / slots;
    const int my_tiles = tiles_mine * imgs_mine;
    auto tile_img = [&](int it) { return slot + (it / tiles_mine) * slots; };
    auto tile_k = [&](int it) { return part + (it % tiles_mine) * parts; };
    auto tile_bytes = [&](int it) { return (uint32_t)(min(kFusedTilePix, HW - tile_k(it) * kFusedTilePix) * 12); };
    auto tile_off = [&](int it) { return ((size_t)tile_img(it) * HW + (size_t)tile_k(it) * kFusedTilePix) * 3; };
    auto stage_ptr = [&](int stg) { return fsmem + (size_t)stg * kFusedTilePix * 12; };
    if (threadIdx.x == 0)
      for (int it = 0; it < my_tiles && it < kFusedStages - 1; ++it)
        bulk_load(stage_ptr(it % kFusedStages), images + tile_off(it), tile_bytes(it), &s_full[it % kFusedStages]);
    double* ysum_img = reinterpret_cast<double*>(ws + L.off_ysum_img);
    int* oor_flags = reinterpret_cast<int*>(ws + L.off_oor);
    double acc = 0.0;
    bool oor = false;
    int sig_it = tiles_mine - 1;                                  // (thread 0) last tile of the next image to report as stored
    for (int it = 0; it < my_tiles; ++it) {
      const int stg = it % kFusedStages;
      bulk_wait_parity(&s_full[stg], (uint32_t)((it / kFusedStages) & 1));
      const int px = threadIdx.x * 4;                             // this thread's four pixels inside the tile
      const float4* tp = reinterpret_cast<const float4*>(stage_ptr(stg)) + threadIdx.x * 3;
      if (px < (int)(tile_bytes(it) / 12)) {
        const float4 a = tp[0], bb = tp[1], c = tp[2];
        acc += (double)luma_of(a.x, a.y, a.z);
        acc += (double)luma_of(a.w, bb.x, bb.y);
        acc += (double)luma_of(bb.z, bb.w, c.x);
        acc += (double)luma_of(c.y, c.z, c.w);
        const float m0 = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
        const float m1 = fmaxf(fmaxf(fabsf(bb.x), fabsf(bb.y)), fmaxf(fabsf(bb.z), fabsf(bb.w)));
        const float m2 = fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w)));
        const float sum = (a.x + a.y + a.z + a.w) + (bb.x + bb.y + bb.z + bb.w) + (c.x + c.y + c.z + c.w);
        oor = oor || !(fmaxf(fmaxf(m0, m1), m2) <= 1.0f) || (sum != sum);        // fmaxf drops NaN: test the sum too
      }
      __syncthreads();                                            // every thread has read the tile
      if (threadIdx.x == 0) {
        bulk_store(out + tile_off(it), stage_ptr(stg), tile_bytes(it));
        const int itn = it + kFusedStages - 1;                    // refill the buffer of the previous tile
        if (itn < my_tiles) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // ... once its outgoing copy has done reading it
          bulk_load(stage_ptr(itn % kFusedStages), images + tile_off(itn), tile_bytes(itn), &s_full[itn % kFusedStages]);
        }
        if (sig_it <= it - 2) {                                   // stores up to tile it - 2 complete: report finished images
          asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
          __threadfence();
          while (sig_it <= it - 2) { atomicAdd(fv.img + 8 * tile_img(sig_it) + 1, 1); sig_it += tiles_mine; }
        }
      }
      if ((it + 1) % tiles_mine == 0) {                           // (uniform) last tile of this image for this CTA
        const int any = __syncthreads_or(oor ? 1 : 0);
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) {
          const int b = tile_img(it);
          atomicAdd(ysum_img + b, acc);
          if (any) atomicOr(oor_flags + b, 1);
          __threadfence();
          const int oldc = atomicAdd(fv.img + 8 * b, 1);
          if (dbg && oldc == pass_expected - 1 && (b == 0 || b == 4 || b == 16 || b == 32)) fv.ctl[28 + (b == 0 ? 0 : b == 4 ? 1 : b == 16 ? 2 : 3)] = (int)(global_ns() - t_base);
        }
        acc = 0.0; oor = false;
      }
    }
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __threadfence();
      while (sig_it < my_tiles) { atomicAdd(fv.img + 8 * tile_img(sig_it) + 1, 1); sig_it += tiles_mine; }
    }
    if (dbg && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(fv.ctl + 10), global_ns() - t_base);
    __syncthreads();                                              // the stage buffers become the warps' scratch
  } else {
    // ---- geometry and patch statistics ------------------------------------------------------------------------------
    const int n_small = n_geom + B * pchunks, nwin = gridDim.x - n_copy;
    for (int blk = blockIdx.x - n_copy; blk < n_small; blk += nwin) {
      const int r = small_roles(s, L, blk, patch, print_wb, boxes, offsets, params, scale, ws, n_geom, B, pchunks, red, fsmem, 0);
      if (r == 2) {
        if (dbg && threadIdx.x == 0) fv.ctl[27] = (int)(global_ns() - t_base);
        build_steps(s, L, ws, offsets, n_geom, skew, nM, reinterpret_cast<int*>(fsmem));
        if (dbg && threadIdx.x == 0) fv.ctl[24] = (int)(global_ns() - t_base);
      }
      __syncthreads();
    }
    if (dbg && threadIdx.x == 0) atomicMax(fv.ctl + 26, (int)(global_ns() - t_base));
  }

  // ---- window work: one ordered task queue, one task per warp at a time ------------------------------------------------
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  wait_ge(fv.ctl, 1, err, nullptr);
  const int total = __ldcg(fv.ctl + 1), steps = B + 2 * skew;
  const int P = s.patch_size;
  const size_t per_warp = resize_warp_smem(s, L);
  float4* inter = reinterpret_cast<float4*>(fsmem + (size_t)warp * per_warp);
  uint32_t* words = reinterpret_cast<uint32_t*>(fsmem + (size_t)warp * per_warp + (size_t)L.rb * P * 16);
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  const int2* items = reinterpret_cast<const int2*>(ws + L.off_items);
  const int2* citems = reinterpret_cast<const int2*>(ws + L.off_citems);
  const float* ubuf = reinterpret_cast<const float*>(ws + L.off_u);
  uint8_t* routes = reinterpret_cast<uint8_t*>(ws + L.off_route);
  const int2* rowtab = reinterpret_cast<const int2*>(ws + L.off_rowtab);
  const int open_cap = (int)L.open_cap;
  int* open_count = counters + 8;
  int2* open_list = reinterpret_cast<int2*>(ws + L.off_open);
  const int n_stat_blocks = B * pchunks;
  WarpTickets tk;
  tk.init(ws, L.off_tickets, 0);
  int t = tk.item(tk.draw(lane));
  while (t < total) {
    const int nxt = prefetch ? tk.draw(lane) : 0;                 // next task's ticket travels while this one computes
    int lo = 0, hi = steps;                                       // step of task t: last k with step_start[k] <= t
    while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (fv.step_start[m] <= t) lo = m; else hi = m; }
    const int k = lo;
    int r = t - fv.step_start[k];
    const int4 info = fv.step_info[k];
    if (r < info.x) {
      // match part r of image k
      const int b = k;
      wait_ge(fv.ctl + 2, n_stat_blocks, err, dbg ? fv.ctl + 20 : nullptr);
      wait_ge(fv.img + 8 * b, pass_expected, err, dbg ? fv.ctl + 20 : nullptr);
      if (r == 0 && lane == 0 && (reinterpret_cast<const int*>(ws + L.off_oor)[b] != 0 || offsets[b + 1] - offsets[b] > 32))
        atomicOr(counters + 6, 1);                                // the tail has whole rows to redo
      match_range(s, L, patch, print_wb, ws, b, r * 32 + lane, info.x * 32);
      warp_signal(fv.img + 8 * b + 2, lane, nM, dbg ? fv.img + 8 * b + 5 : nullptr, t_base);
    } else if ((r -= info.x) < info.y) {
      // resize item of image k - skew
      const int b = k - skew;
      const int2 item = __ldcg(items + info.z + r);
      wait_ge(fv.img + 8 * b + 2, nM, err, dbg ? fv.ctl + 21 : nullptr);
      const BoxPlan* pl = plans + item.x;
      const int mode = pl->two_tap ? 2 : pl->span;
      if (mode == 2) resize_item<2>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
      else if (mode == 5) resize_item<5>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
      else if (mode == 7) resize_item<7>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
      else if (mode == 3) resize_item<3>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
      else if (mode == 9) resize_item<9>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
      else resize_item<0>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
      warp_signal(fv.img + 8 * b + 3, lane, info.y, dbg ? fv.img + 8 * b + 6 : nullptr, t_base);
    } else {
      // composite item of image k - 2 skew
      r -= info.y;
      const int b = k - 2 * skew;
      const int2 item = __ldcg(citems + info.w + r);
      const int n_resize = fv.step_info[b + skew].y;
      const int n_comp = fv.step_start[k + 1] - fv.step_start[k] - info.x - info.y;
      wait_ge(fv.img + 8 * b + 3, n_resize, err, dbg ? fv.ctl + 22 : nullptr);
      wait_ge(fv.img + 8 * b + 1, pass_expected, err, dbg ? fv.ctl + 23 : nullptr);
      composite_item_main<false>(s, L, ws, plans, ubuf, rowtab, routes, images, out, nullptr, item.x, item.y, open_count, open_list,
                                 open_cap, one, lane);
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        if (atomicAdd(fv.img + 8 * b + 4, 1) == n_comp - 1) {
          atomicAdd(fv.ctl + 3, 1);
          if (dbg) fv.img[8 * b + 7] = (int)(global_ns() - t_base);
        }
      }
    }
    t = tk.item(prefetch ? nxt : tk.draw(lane));
  }
  // ---- tail: open pixels and general rows, shared by every thread of the grid -------------------------------------------
  if (dbg && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(fv.ctl + 12), global_ns() - t_base);
  wait_ge(fv.ctl + 3, B, err, nullptr);
  if (dbg && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(fv.ctl + 14), global_ns() - t_base);
  composite_rest_body<false>(s, L, ws, images, out, nullptr, offsets, 0, B, 0, 1, blockIdx.x * blockDim.x + threadIdx.x,
                             gridDim.x * blockDim.x);
  if (dbg && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(fv.ctl + 16), global_ns() - t_base);
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

// Auxiliary stream + events of the calling host thread for the current device (created on first use, kept for the life
// of the thread): the window kernels of image group g run there while the caller's stream copies group g + 1.  Nothing
// but these handles is cached; every call forks from and joins back into the caller's stream, so the call remains
// asynchronous on that stream and can be captured into a CUDA graph (the fork / join become graph edges).
struct AuxStream {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t pre[kMaxGroups] = {};
  cudaEvent_t done = nullptr;
};
static int aux_stream(AuxStream** out) {
  static thread_local AuxStream aux[8];
  int dev = 0;
  EOT_CHECK_CUDA(cudaGetDevice(&dev));
  AuxStream* a = nullptr;
  for (auto& e : aux)
    if (e.device == dev || e.device < 0) { a = &e; break; }
  if (!a) { set_error("eot_apply_fwd: more than 8 devices used by one host thread"); return EOT_ERR_CUDA; }
  if (a->device < 0) {
    EOT_CHECK_CUDA(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
    for (auto& e : a->pre) EOT_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    EOT_CHECK_CUDA(cudaEventCreateWithFlags(&a->done, cudaEventDisableTiming));
    a->device = dev;
  }
  *out = a;
  return EOT_OK;
}

static int forward_groups(int B, int N) {
  static const int forced = [] { const char* e = getenv("EOT_FWD_GROUPS"); return e ? atoi(e) : 0; }();
  // default 1: measured on B200 (profiles/r02_forward.md), the persistent window kernels take every register of an SM,
  // so the image pass of the next group cannot co-reside and each extra group only adds launches (+40..50 us per group
  // under graph replay).  EOT_FWD_GROUPS keeps the A/B reproducible.
  int g = forced > 0 ? forced : 1;
  if (N == 0) g = 1;
  return g < 1 ? 1 : (g > kMaxGroups ? kMaxGroups : (g > B ? B : g));
}

// The fused forward (k_forward_fused): memset of the accumulators + ONE cooperative launch.  Shapes it does not take
// (mask output, in-place, unaligned rows, no boxes, image groups) stay on the kernel sequence of launch_forward.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}
static int launch_forward_fused(const EotShape& s, const Layout& L, const float* patch, const float* scale, const float* images,
                                const float* boxes, const int32_t* box_offsets, const EotBoxParams* params,
                                const float* print_wb, float* out_images, char* ws, cudaStream_t st, bool* taken) {
  *taken = false;
  const int B = s.batch, P = s.patch_size, N = s.total_boxes;
  const size_t smem_win = resize_warp_smem(s, L) * (kThreads / 32);
  const size_t smem_copy = (size_t)kFusedStages * kFusedTilePix * 12;
  const size_t smem = smem_win > smem_copy ? smem_win : smem_copy;
  if (smem > 200 * 1024 || (size_t)(B + 2 * kMaxSkew) * 4 > smem) return EOT_OK;   // (launch_forward reports the patch side)
  struct Cfg { size_t smem; int per_sm; int dev; };
  static thread_local Cfg cfg = {0, 0, -1};
  int dev = 0;
  EOT_CHECK_CUDA(cudaGetDevice(&dev));
  if (cfg.smem != smem || cfg.dev != dev) {
    EOT_CHECK_CUDA(cudaFuncSetAttribute(k_forward_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    EOT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_forward_fused, kThreads, smem));
    cfg = {smem, occ, dev};
  }
  if (cfg.per_sm < 2) return EOT_OK;
  static const int copy_per_sm = env_int("EOT_FUSED_COPY", 0), slots_env = env_int("EOT_FUSED_SLOTS", 4),
                   skew_env = env_int("EOT_FUSED_SKEW", 16), prefetch_env = env_int("EOT_FUSED_PREFETCH", 1),
                   dbg_env = env_int("EOT_KERNEL_TIMES", 0);
  const int nsm = sm_count();
  const int grid = nsm * cfg.per_sm;
  int slots = slots_env < 1 ? 1 : (slots_env > B ? B : slots_env);
  int cps = copy_per_sm > 0 ? copy_per_sm : cfg.per_sm / 2;
  if (cps > cfg.per_sm - 1) cps = cfg.per_sm - 1;
  int n_copy = nsm * cps / slots * slots;                         // `slots` images at a time, n_copy / slots CTAs each
  if (n_copy < slots) return EOT_OK;
  const int skew = skew_env < 1 ? 1 : (skew_env > kMaxSkew ? kMaxSkew : skew_env);
  int nM = (P * P + kMatchTexels - 1) / kMatchTexels;
  int pchunks = max(1, min((P * P + kThreads * 4 - 1) / (kThreads * 4), 64));
  int n_geom = N;
  float one = 1.0f;
  int prefetch = prefetch_env, dbg = dbg_env;
  EotShape sv = s;
  Layout Lv = L;
  EOT_CHECK_CUDA(cudaMemsetAsync(ws + L.off_ysum_img, 0, L.off_plans - L.off_ysum_img, st));
  void* args[] = {&sv, &Lv, &patch, &print_wb, &boxes, &box_offsets, &params, &scale, &images, &out_images, &ws,
                  &n_geom, &pchunks, &n_copy, &slots, const_cast<int*>(&skew), &nM, &one, &prefetch, &dbg};
  EOT_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)k_forward_fused, dim3(grid), dim3(kThreads), args, smem, st));
  count_launches(1);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  if (dbg) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
      int ctl[32];
      EOT_CHECK_CUDA(cudaMemcpyAsync(ctl, ws + L.off_fused, sizeof(ctl), cudaMemcpyDeviceToHost, st));
      EOT_CHECK_CUDA(cudaStreamSynchronize(st));
      const unsigned long long* tt = reinterpret_cast<const unsigned long long*>(ctl + 8);
      if (dbg > 1) {
        static int img[8 * 4096];
        const int nb = B < 4096 ? B : 4096;
        EOT_CHECK_CUDA(cudaMemcpy(img, ws + L.off_fused + 256, (size_t)nb * 32, cudaMemcpyDeviceToHost));
        for (int b = 0; b < nb; b += (nb > 32 ? nb / 32 : 1))
          fprintf(stderr, "[eot]   image %3d: match %6.1f resize %6.1f composite %6.1f us (items %d %d)\n", b, img[8 * b + 5] * 1e-3,
                  img[8 * b + 6] * 1e-3, img[8 * b + 7] * 1e-3, img[8 * b + 3], img[8 * b + 4]);
      }
      fprintf(stderr, "[eot] fused: last geometry block starts tables %.1f us, tables ready %.1f, statistics done %.1f, small roles end (max) %.1f; pass of image 0 / 4 / 16 / 32 done %.1f %.1f %.1f %.1f\n",
              ctl[27] * 1e-3, ctl[24] * 1e-3, ctl[25] * 1e-3, ctl[26] * 1e-3, ctl[28] * 1e-3, ctl[29] * 1e-3, ctl[30] * 1e-3, ctl[31] * 1e-3);
      fprintf(stderr, "[eot] fused: tasks %d, copy done %.1f us, queue done %.1f us, composites done %.1f us, end %.1f us; polls M %d R %d C %d store %d (grid %d, copy CTAs %d, slots %d, skew %d)\n",
              ctl[1], tt[1] * 1e-3, tt[2] * 1e-3, tt[3] * 1e-3, tt[4] * 1e-3, ctl[20], ctl[21], ctl[22], ctl[23], grid, n_copy, slots, skew);
    }
  }
  *taken = true;
  return EOT_OK;
}

// Enqueues the whole forward: memset of the small accumulators, then per image group the pre-pass (image pass; the
// first one also carries the geometry and patch-statistics roles) on the caller's stream and match -> resize ->
// composite on the auxiliary stream, so that the HBM-bound image pass of group g + 1 overlaps the L2-resident,
// instruction-bound window work of group g.
static int launch_forward(const EotShape& s, const Layout& L, const float* patch, const float* scale, const float* images,
                          const float* boxes, const int32_t* box_offsets, const EotBoxParams* params,
                          const float* print_wb, float* out_images, float* mask, char* ws, cudaStream_t st) {
  const int B = s.batch, P = s.patch_size, HW = s.height * s.width, N = s.total_boxes;
  {
    static const int fused_on = env_int("EOT_FWD_FUSED", 0);
    const bool vec0 = (HW % 4 == 0) && (((uintptr_t)images | (uintptr_t)out_images) & 15) == 0;
    if (fused_on && N > 0 && vec0 && !mask && out_images != images && forward_groups(B, N) == 1) {
      bool taken = false;
      StageTimer ftimer(st, "eot_apply_fwd (fused)");
      if (int rc = launch_forward_fused(s, L, patch, scale, images, boxes, box_offsets, params, print_wb, out_images, ws, st, &taken))
        return rc;
      if (taken) { ftimer.mark("all"); return EOT_OK; }
    }
  }
  StageTimer timer(st, "eot_apply_fwd");
  EOT_CHECK_CUDA(cudaMemsetAsync(ws + L.off_ysum_img, 0, L.off_plans - L.off_ysum_img, st));
  const int pchunks = max(1, min((P * P + kThreads * 4 - 1) / (kThreads * 4), 64));
  const int cpi = (HW + kPassPixPerBlock - 1) / kPassPixPerBlock;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)images | (uintptr_t)out_images | (uintptr_t)(mask ? mask : out_images)) & 15) == 0;
  const int G = forward_groups(B, N);
  static const int sort_env = env_int("EOT_SORT_ITEMS", 1);
  const int sort_items = (G == 1 && sort_env) ? 1 : 0;               // one call takes the whole item list: largest boxes first
  AuxStream* aux = nullptr;
  if (G > 1)
    if (int rc = aux_stream(&aux)) return rc;
  float* pass_out = out_images;
  for (int g = 0; g < G; ++g) {
    const int b0 = (int)((long long)B * g / G), b1 = (int)((long long)B * (g + 1) / G);
    const int n_geom = g == 0 ? N : 0, n_stat = g == 0 ? B : 0;
    const long long nblocks = (long long)n_geom + (long long)n_stat * pchunks + (long long)(b1 - b0) * cpi;
    if (nblocks >= (1ll << 31)) { set_error("eot_apply_fwd: grid too large"); return EOT_ERR_BAD_SHAPE; }
    // bulk-copy pipeline when the pass really copies (out of place, no mask) and tiles / rows are 16-byte aligned
    static const int bulk_env = env_int("EOT_PREPASS_BULK_ON", 1);
    const bool bulk = EOT_PREPASS_BULK && bulk_env && vec && !mask && out_images != images;
    if (bulk) {
      const size_t smem = (size_t)kBulkStages * kBulkTilePix * 12;
      static thread_local bool attr_set = false;
      if (!attr_set) {
        EOT_CHECK_CUDA(cudaFuncSetAttribute(k_prepass_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      const long long tiles = (long long)(b1 - b0) * ((HW + kBulkTilePix - 1) / kBulkTilePix);
      const int copy_ctas = (int)(tiles < (long long)EOT_BULK_CTAS * sm_count() ? tiles : (long long)EOT_BULK_CTAS * sm_count());
      k_prepass_bulk<<<(unsigned)(n_geom + n_stat * pchunks + copy_ctas), kThreads, smem, st>>>(
          s, L, patch, print_wb, boxes, box_offsets, params, scale, images, pass_out, ws, n_geom, n_stat, pchunks, copy_ctas, b0, b1, sort_items);
    } else if (vec)
      k_prepass<true><<<(unsigned)nblocks, kThreads, 0, st>>>(s, L, patch, print_wb, boxes, box_offsets, params, scale, images,
                                                              pass_out, mask, ws, n_geom, n_stat, pchunks, cpi, b0, sort_items);
    else
      k_prepass<false><<<(unsigned)nblocks, kThreads, 0, st>>>(s, L, patch, print_wb, boxes, box_offsets, params, scale, images,
                                                               pass_out, mask, ws, n_geom, n_stat, pchunks, cpi, b0, sort_items);
    count_launches(1);
    if (G == 1) timer.mark("prepass");
    if (N == 0) continue;
    cudaStream_t sw = st;
    if (G > 1) {
      EOT_CHECK_CUDA(cudaEventRecord(aux->pre[g], st));
      EOT_CHECK_CUDA(cudaStreamWaitEvent(aux->stream, aux->pre[g], 0));
      sw = aux->stream;
    }
    EOT_CHECK_CUDA(launch_pdl(k_match, dim3((b1 - b0) * pchunks), dim3(kThreads), 0, sw, s, L, patch, print_wb, box_offsets, ws, b0, pchunks));
    count_launches(1);
    if (G == 1) timer.mark("match");
    if (int rc = launch_resize2(s, L, ws, box_offsets, b0, b1, 2 * g, sw)) return rc;
    if (G == 1) timer.mark("resize");
    if (int rc = launch_composite3(s, L, ws, box_offsets, images, out_images, mask, b0, b1, g, G, sw)) return rc;
    if (G == 1) timer.mark("composite");
  }
  if (G > 1 && N > 0) {
    EOT_CHECK_CUDA(cudaEventRecord(aux->done, aux->stream));
    EOT_CHECK_CUDA(cudaStreamWaitEvent(st, aux->done, 0));
    timer.mark("all");
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
  if (flag & 8) { set_error("box_offsets[B] exceeds shape.total_boxes (the box capacity of the call): the surplus boxes were dropped"); return EOT_ERR_GEOMETRY; }
  if (flag & 4) { set_error("internal: the transposed resize-weight table overflowed its tap capacity"); return EOT_ERR_GEOMETRY; }
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
