// Backward of the EOT patch application: dL/d(out_images) -> dL/dpatch
// (reference: tape.gradient at attacker.py:217; chain of SURVEY.md section 3.2 / App. D).
//
//   k_bwd_image   one CTA per (image, strip of patch rows), looping over the image's boxes with a SHARED-MEMORY
//                 accumulator of the strip's patch gradient (no global atomics, no per-box partial buffers, no g_u
//                 round trip):
//                   A  dL/d(u) of the output rows the strip's taps reach: TF's registered gradient of
//                      ImageProjectiveTransformV3 -- the SAME bilinear warp applied to dL/d(window) with the inverted
//                      transform, fill 0 -- reading dL/d(window) through the route bytes the forward composite left
//                      (TensorScatterUpdate / SelectV2 / outer clip routing), then the inner clip mask of
//                      attacker.py:428 -> a shared-memory tile (RGBX texels)
//                   B  rows pass of the exact resize transpose (ScaleAndTranslateGrad) over the transposed weight
//                      tables the forward's geometry role wrote -> shared-memory intermediate rows
//                   C  columns pass, accumulated into the strip
//                 wide boxes are walked in column chunks sized to the shared-memory budget; after the last box: first half
//                 of the BrightnessMatcher backward (clip mask, K'^T) and the per-image sum of dL/dY (block tree reduction).
//   k_bwd_texel   second half: subtract the per-image mean of dL/dY, K^T, rescale, print-adjust clip mask and weights;
//                 partial sums over image groups, summed in a fixed order by the last block of each texel chunk.
#include "eot_common.cuh"

namespace eot {

constexpr int kBwdGroups = 16;   // image groups of k_bwd_texel

// gradient that reaches R_j at a window pixel of a box, 3 channels: the route byte says which channels of the pasted
// pixel came from this box (SelectV2), passed the outer clip and were not overwritten by a later paste
// (TensorScatterUpdate grad).  `px` = pixel index inside the image plane window (32-bit: H*W*3 < 2^31 is checked).
__device__ __forceinline__ void routed_grad3_bits(unsigned bits, const float* __restrict__ Gwin, int px, float g[3]) {
  const float* gp = Gwin + px * 3;
  g[0] = (bits & 1u) ? __ldg(gp) : 0.0f;
  g[1] = (bits & 2u) ? __ldg(gp + 1) : 0.0f;
  g[2] = (bits & 4u) ? __ldg(gp + 2) : 0.0f;
}
__device__ __forceinline__ void routed_grad3(const uint8_t* __restrict__ route, const float* __restrict__ Gwin, int rt, int px,
                                             float g[3]) {
  routed_grad3_bits(route[rt], Gwin, px, g);
}

// dL/d(u) of texel (ty, tx) of box `me`, before the inner clip mask: TF's registered gradient of
// ImageProjectiveTransformV3 = the same bilinear warp applied to dL/d(window) with the inverted transform, fill 0.
struct WindowCtx {
  const uint8_t* route;     // route map of the box (d x d bytes)
  const float* Gwin;        // dL/d(out) at the window's first pixel
  int D, W;
  float Dm1, Df;
  bool affine;
};
__device__ __forceinline__ void window_grad_texel(const BoxPlan& me, const WindowCtx& wc, int ty, int tx, float g[3]) {
  const float yf = (float)(ty + me.pad_lo), xf = (float)(tx + me.pad_lo);
  float ix = (me.Ti[0] * xf + me.Ti[1] * yf) + me.Ti[2];
  float iy = (me.Ti[3] * xf + me.Ti[4] * yf) + me.Ti[5];
  g[0] = g[1] = g[2] = 0.0f;
  if (!wc.affine) {
    const float proj = (me.Ti[6] * xf + me.Ti[7] * yf) + 1.0f;
    if (proj == 0.0f) return;
    ix = ix / proj;
    iy = iy / proj;
  }
  // floorf for |v| < 2^22 as two full-rate additions (v + 1.5 * 2^23 rounded down; the integer sits in the mantissa);
  // anything further out (or NaN) lands outside the window in the tests below, as it must
  const float mx = __fadd_rd(ix, 12582912.0f), my = __fadd_rd(iy, 12582912.0f);
  const float x0f = mx - 12582912.0f, y0f = my - 12582912.0f;
  const float wx0 = ix - x0f, wx1 = (x0f + 1.0f) - ix, wy0 = iy - y0f, wy1 = (y0f + 1.0f) - iy;
  const int D = wc.D, W = wc.W;
  float v00[3], v01[3], v10[3], v11[3];
  if (x0f >= 0.0f && x0f < wc.Dm1 && y0f >= 0.0f && y0f < wc.Dm1) {            // all four taps inside the window
    const int xi0 = __float_as_int(mx) - 0x4B400000, yi0 = __float_as_int(my) - 0x4B400000;
    const int rt = yi0 * D + xi0, px = yi0 * W + xi0;
    // the four route bytes first, then all (predicated) gradient loads: one round trip each, not two per row
    const unsigned r00 = wc.route[rt], r01 = wc.route[rt + 1], r10 = wc.route[rt + D], r11 = wc.route[rt + D + 1];
    routed_grad3_bits(r00, wc.Gwin, px, v00);
    routed_grad3_bits(r01, wc.Gwin, px + 1, v01);
    routed_grad3_bits(r10, wc.Gwin, px + W, v10);
    routed_grad3_bits(r11, wc.Gwin, px + W + 1, v11);
  } else {
    const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
    const bool bx0 = x0f >= 0.0f && x0f < wc.Df, bx1 = x1f >= 0.0f && x1f < wc.Df;
    const bool by0 = y0f >= 0.0f && y0f < wc.Df, by1 = y1f >= 0.0f && y1f < wc.Df;
#pragma unroll
    for (int c = 0; c < 3; ++c) v00[c] = v01[c] = v10[c] = v11[c] = 0.0f;
    if (!((bx0 || bx1) && (by0 || by1))) return;
    const int xi0 = __float_as_int(mx) - 0x4B400000, yi0 = __float_as_int(my) - 0x4B400000;
    const int rt = yi0 * D + xi0, px = yi0 * W + xi0;
    if (by0 && bx0) routed_grad3(wc.route, wc.Gwin, rt, px, v00);
    if (by0 && bx1) routed_grad3(wc.route, wc.Gwin, rt + 1, px + 1, v01);
    if (by1 && bx0) routed_grad3(wc.route, wc.Gwin, rt + D, px + W, v10);
    if (by1 && bx1) routed_grad3(wc.route, wc.Gwin, rt + D + 1, px + W + 1, v11);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) g[c] = wy1 * (wx1 * v00[c] + wx0 * v01[c]) + wy0 * (wx1 * v10[c] + wx0 * v11[c]);
}

// ---- resize adjoint passes over shared memory ---------------------------------------------------------------------------
// TK > 0: compile-time tap count (taps past a row's true count carry weight 0 and a clamped index); TK == 0: run-time.
//   rows pass     inter[r][c]  = sum_k wT[py0 + r][k] * tile[min(st(py) + k, oy_last) - oy_lo][c]
//   columns pass  acc[r][px]  += sum_{k: ox0 <= st(px) + k < ox0 + cw} wT[px][k] * inter[r][st(px) + k - ox0]
template <int TK>
__device__ __forceinline__ void adjoint_passes(int P, int oy_last, int tstride, int py0, int rows, int oy_lo, int ox0, int cw,
                                               int ncols_total,
                                               const float4* tile, float4* inter, const float* s_wt, const int2* s_st,
                                               float* acc) {
  const float inv_cw = 1.0f / (float)cw;
  for (int idx = threadIdx.x; idx < rows * cw; idx += blockDim.x) {
    int r = (int)(((float)idx + 0.5f) * inv_cw);
    int c = idx - r * cw;
    if (c < 0) { --r; c += cw; } else if (c >= cw) { ++r; c -= cw; }
    const int2 sc = s_st[py0 + r];
    const float* w = s_wt + (py0 + r) * tstride;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    if (TK > 0) {
#pragma unroll
      for (int k = 0; k < (TK > 0 ? TK : 1); ++k) {
        const float4 v = tile[(min(sc.x + k, oy_last) - oy_lo) * cw + c];
        const float wk = w[k];
        a0 += wk * v.x; a1 += wk * v.y; a2 += wk * v.z;
      }
    } else {
      for (int k = 0; k < sc.y; ++k) {
        const float4 v = tile[(sc.x + k - oy_lo) * cw + c];
        const float wk = w[k];
        a0 += wk * v.x; a1 += wk * v.y; a2 += wk * v.z;
      }
    }
    inter[idx] = make_float4(a0, a1, a2, 0.0f);
  }
  __syncthreads();
  const int P3 = P * 3;
  const float inv_P = 1.0f / (float)P;
  const bool whole = TK > 0 && ox0 == 0 && ncols_total == cw;     // the chunk is the whole box width
  for (int idx = threadIdx.x; idx < rows * P; idx += blockDim.x) {
    int r = (int)(((float)idx + 0.5f) * inv_P);
    int px = idx - r * P;
    if (px < 0) { --r; px += P; } else if (px >= P) { ++r; px -= P; }
    const int2 sc = s_st[px];
    const float* w = s_wt + px * tstride;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    if (whole) {                                                  // fixed tap count: zero-weight padded taps at a clamped column
      const float4* irow = inter + r * cw;
#pragma unroll
      for (int k = 0; k < (TK > 0 ? TK : 1); ++k) {
        const float4 v = irow[min(sc.x + k, cw - 1)];
        const float wk = w[k];
        a0 += wk * v.x; a1 += wk * v.y; a2 += wk * v.z;
      }
    } else {
      const int k0 = max(0, ox0 - sc.x), k1 = min(sc.y, ox0 + cw - sc.x);    // this chunk's share of the taps of px
      if (k0 >= k1) continue;
      const float4* irow = inter + r * cw - ox0;
      for (int k = k0; k < k1; ++k) {
        const float4 v = irow[sc.x + k];
        const float wk = w[k];
        a0 += wk * v.x; a1 += wk * v.y; a2 += wk * v.z;
      }
    }
    float* o = acc + r * P3 + px * 3;
    o[0] += a0; o[1] += a1; o[2] += a2;
  }
}

#ifndef EOT_BWD_MINB
#define EOT_BWD_MINB 3
#endif
#ifndef EOT_BWD_ROWS
#define EOT_BWD_ROWS 16
#endif
#ifndef EOT_BWD_SMEM_KB
#define EOT_BWD_SMEM_KB 72
#endif
// Shared memory of k_bwd_image: strip accumulator, the box's transposed tables, tile + intermediate rows.
struct BwdSmem {
  int rows;            // patch rows per strip
  int budget;          // texels of tile + intermediate rows
  size_t off_st, off_wt, off_buf, total;
};
__host__ __device__ inline BwdSmem bwd_smem_plan(const EotShape& s, const Layout& L) {
  BwdSmem m;
  const int P = s.patch_size;
  int rows = (24 * 1024) / (P * 12);
  rows = rows < 1 ? 1 : (rows > EOT_BWD_ROWS ? EOT_BWD_ROWS : rows);
  // enough (image, strip) items for the heaviest-first hand-out to balance (about 1.5 per resident CTA) -- but not
  // thinner than 4 rows: every strip recomputes the output rows its outermost taps share with its neighbours
  while (rows > 4 && 2ll * s.batch * ((P + rows - 1) / rows) < 3ll * 148 * EOT_BWD_MINB) rows = (rows + 1) / 2;
  m.rows = rows;
  const size_t acc = align_up((size_t)m.rows * P * 12, 16);
  m.off_st = acc;
  m.off_wt = align_up(m.off_st + (size_t)(P + 1) * 8, 16);
  m.off_buf = align_up(m.off_wt + (size_t)P * L.tcap * 4, 16);
  const size_t cap = (size_t)EOT_BWD_SMEM_KB * 1024;
  size_t buf = m.off_buf < cap ? cap - m.off_buf : 0;
  if (buf < 24 * 1024) buf = 24 * 1024;                           // large tables: fewer CTAs per SM instead of tiny chunks
  m.budget = (int)(buf / 16);
  m.total = m.off_buf + (size_t)m.budget * 16;
  return m;
}

__global__ void __launch_bounds__(kThreads, EOT_BWD_MINB) k_bwd_image(EotShape s, Layout L, char* ws, const float* __restrict__ G,
                                                                     const float* __restrict__ patch,
                                                                     const float* __restrict__ print_wb) {
  extern __shared__ __align__(16) unsigned char bsm[];
  __shared__ double red[32];
  __shared__ int s_item;
  pdl_trigger();
  const BwdSmem m = bwd_smem_plan(s, L);
  float* acc = reinterpret_cast<float*>(bsm);
  int2* s_st = reinterpret_cast<int2*>(bsm + m.off_st);
  float* s_wt = reinterpret_cast<float*>(bsm + m.off_wt);
  float4* buf = reinterpret_cast<float4*>(bsm + m.off_buf);
  const int P = s.patch_size, P3 = P * 3, H = s.height, W = s.width;
  const int nstrips = (P + m.rows - 1) / m.rows;
  const int n_items = s.batch * nstrips;
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  const int32_t* offsets = reinterpret_cast<const int32_t*>(ws + L.off_offsets);
  const float* ubuf = reinterpret_cast<const float*>(ws + L.off_u);
  int* ticket = reinterpret_cast<int*>(ws + L.off_bwd_cnt);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1);
    __syncthreads();
    const int it = s_item;
    if (it >= n_items) break;
    // heaviest images first; the strips of one image are consecutive tickets (its routes and gradients stay L2-hot)
    const int b = reinterpret_cast<const int*>(ws + L.off_order)[it / nstrips], py0 = (it % nstrips) * m.rows;
    const int rows = min(m.rows, P - py0);
    for (int i = threadIdx.x; i < rows * P3; i += blockDim.x) acc[i] = 0.0f;
    for (int j = offsets[b]; j < offsets[b + 1]; ++j) {
      const BoxPlan me = plans[j];
      if (!me.valid) continue;
      const int ps = me.ps;
      __syncthreads();                                            // the previous box is done with the tables
      const int2* stt = reinterpret_cast<const int2*>(ws + L.off_stt) + (size_t)j * (P + 1);
      for (int i = threadIdx.x; i <= P; i += blockDim.x) s_st[i] = stt[i];
      __syncthreads();
      const int tstride = s_st[P].x;
      const float* wt = reinterpret_cast<const float*>(ws + L.off_wt) + (size_t)j * P * L.tcap;
      for (int i = threadIdx.x; i < P * tstride; i += blockDim.x) s_wt[i] = wt[i];
      // output rows the strip's taps reach (the zero-weight padded taps of the fixed-count passes are clamped into it)
      const int oy_lo = s_st[py0].x;
      int oy_hi = 0;
      for (int r = 0; r < rows; ++r) oy_hi = max(oy_hi, s_st[py0 + r].x + s_st[py0 + r].y);
      oy_hi = min(oy_hi, ps);
      const int n_oy = oy_hi - oy_lo;
      if (n_oy <= 0) continue;                                    // (uniform) no output row samples this strip
      int cwmax = m.budget / (n_oy + rows);
      if (cwmax < 1) {                                            // cannot happen for shapes make_layout accepts; flagged, not silent
        if (threadIdx.x == 0) atomicOr(reinterpret_cast<int*>(ws + L.off_counters) + 2, 16);
        continue;
      }
      cwmax = min(cwmax, ps);
      WindowCtx wc;
      wc.route = reinterpret_cast<const uint8_t*>(ws + L.off_route) + (size_t)j * L.rslot;
      wc.Gwin = G + (((size_t)me.image * H + me.y0) * W + me.x0) * 3;
      wc.D = me.d; wc.W = W;
      wc.Dm1 = (float)(me.d - 1); wc.Df = (float)me.d;
      wc.affine = (me.Ti[6] == 0.0f && me.Ti[7] == 0.0f);
      const float4* u4 = reinterpret_cast<const float4*>(ubuf + me.u_off);
      const int S = u_stride(ps);
      const int nchunks = (ps + cwmax - 1) / cwmax;
      const int cwb = (ps + nchunks - 1) / nchunks;               // balanced chunks
      for (int ox0 = 0; ox0 < ps; ox0 += cwb) {
        const int cw = min(cwb, ps - ox0);
        float4* tile = buf;
        float4* inter = buf + n_oy * cw;
        __syncthreads();                                          // the previous chunk is done with tile / inter (and the tables are staged)
        // A: dL/d(u) tile
        const float inv_cw = 1.0f / (float)cw;
        for (int idx = threadIdx.x; idx < n_oy * cw; idx += blockDim.x) {
          int r = (int)(((float)idx + 0.5f) * inv_cw);
          int c = idx - r * cw;
          if (c < 0) { --r; c += cw; } else if (c >= cw) { ++r; c -= cw; }
          const int ty = oy_lo + r, tx = ox0 + c;
          const unsigned bits = __float_as_uint(__ldg(&u4[(ty + 2) * S + tx + 2].w));   // inner clip pass bits (attacker.py:428)
          float g[3];
          window_grad_texel(me, wc, ty, tx, g);
          tile[idx] = make_float4((bits & 1u) ? g[0] : 0.0f, (bits & 2u) ? g[1] : 0.0f, (bits & 4u) ? g[2] : 0.0f, 0.0f);
        }
        __syncthreads();
        // B + C
        if (tstride == 3) adjoint_passes<3>(P, oy_hi - 1, tstride, py0, rows, oy_lo, ox0, cw, ps, tile, inter, s_wt, s_st, acc);
        else if (tstride == 4) adjoint_passes<4>(P, oy_hi - 1, tstride, py0, rows, oy_lo, ox0, cw, ps, tile, inter, s_wt, s_st, acc);
        else if (tstride == 5) adjoint_passes<5>(P, oy_hi - 1, tstride, py0, rows, oy_lo, ox0, cw, ps, tile, inter, s_wt, s_st, acc);
        else if (tstride == 6) adjoint_passes<6>(P, oy_hi - 1, tstride, py0, rows, oy_lo, ox0, cw, ps, tile, inter, s_wt, s_st, acc);
        else adjoint_passes<0>(P, oy_hi - 1, tstride, py0, rows, oy_lo, ox0, cw, ps, tile, inter, s_wt, s_st, acc);
      }
    }
    __syncthreads();
    // BrightnessMatcher backward, first half (brightness_matcher.py:65-72 reversed)
    const double* ysum_img = reinterpret_cast<const double*>(ws + L.off_ysum_img);
    const double* ysum_patch = reinterpret_cast<const double*>(ws + L.off_ysum_patch);
    const float mu_t = (float)(ysum_img[b] / (double)((size_t)H * W));
    const float mu_s = (float)(ysum_patch[b] / (double)((size_t)P * P));
    const float* wb = print_wb + (size_t)b * 6;
    float* gm = reinterpret_cast<float*>(ws + L.off_gm) + (size_t)b * P * P3;
    double gy_acc = 0.0;
    for (int idx = threadIdx.x; idx < rows * P; idx += blockDim.x) {
      const int r = idx / P, px = idx - r * P;
      const int py = py0 + r;
      const float* p = patch + ((size_t)py * P + px) * 3;
      const TexelYuv y = texel_yuv(__ldg(p), __ldg(p + 1), __ldg(p + 2), wb);
      const float y_pre = (y.y - mu_s) + mu_t;
      const float yp = clampf(y_pre, 0.0f, 1.0f);
      const float rr = (yp * 1.0f + y.u * EOT_I10) + y.v * EOT_I20;
      const float gg = (yp * 1.0f + y.u * EOT_I11) + y.v * EOT_I21;
      const float bb = (yp * 1.0f + y.u * EOT_I12) + y.v * EOT_I22;
      const float* a = acc + r * P3 + px * 3;
      const float g0 = (rr >= 0.0f && rr <= 1.0f) ? a[0] * EOT_C255_127 : 0.0f;
      const float g1 = (gg >= 0.0f && gg <= 1.0f) ? a[1] * EOT_C255_127 : 0.0f;
      const float g2 = (bb >= 0.0f && bb <= 1.0f) ? a[2] * EOT_C255_127 : 0.0f;
      float gY = g0 + g1 + g2;                                     // K' row Y = (1,1,1)
      const float gU = g0 * EOT_I10 + g1 * EOT_I11 + g2 * EOT_I12;
      const float gV = g0 * EOT_I20 + g1 * EOT_I21 + g2 * EOT_I22;
      if (!(y_pre >= 0.0f && y_pre <= 1.0f)) gY = 0.0f;
      gy_acc += (double)gY;
      float* o = gm + ((size_t)py * P + px) * 3;
      o[0] = gY; o[1] = gU; o[2] = gV;
    }
    gy_acc = block_sum(gy_acc, red);
    if (threadIdx.x == 0) atomicAdd(reinterpret_cast<double*>(ws + L.off_gy_sum) + b, gy_acc);
  }
}

// Second half of the BrightnessMatcher / print-adjust backward, summed over the images: block (x, g) sums image group
// g for a chunk of texels into its partial; the last block of a chunk to finish adds the groups in a fixed order.
__global__ void __launch_bounds__(kThreads) k_bwd_texel(EotShape s, Layout L, char* ws, const float* __restrict__ patch,
                                                        const float* __restrict__ print_wb, int groups, float* grad_patch,
                                                        int accumulate) {
  __shared__ int s_last;
  pdl_wait();
  const int P = s.patch_size, PP = P * P;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  float* parts = reinterpret_cast<float*>(ws + L.off_gp_part);
  if (t < PP) {
    const float p0 = __ldg(patch + (size_t)t * 3), p1 = __ldg(patch + (size_t)t * 3 + 1), p2 = __ldg(patch + (size_t)t * 3 + 2);
    const float* gmb = reinterpret_cast<const float*>(ws + L.off_gm);
    const double* gy_sum = reinterpret_cast<const double*>(ws + L.off_gy_sum);
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int b = g; b < s.batch; b += groups) {
      const float* wb = print_wb + (size_t)b * 6;
      const float* gm = gmb + ((size_t)b * PP + t) * 3;
      const float mean_gy = (float)(gy_sum[b] / (double)PP);
      const float gYs = gm[0] - mean_gy;                            // d(-mean(Ys)) term
      const float gU = gm[1], gV = gm[2];
      // g_s = g_yuv . K^T ; g_q = g_s * 127/255 ; print adjust: clip mask and weight
      const float s0 = (gYs * EOT_K00 + gU * EOT_K01 + gV * EOT_K02) * EOT_C127_255;
      const float s1 = (gYs * EOT_K10 + gU * EOT_K11 + gV * EOT_K12) * EOT_C127_255;
      const float s2 = (gYs * EOT_K20 + gU * EOT_K21 + gV * EOT_K22) * EOT_C127_255;
      const float q0 = wb[0] * p0 + wb[3], q1 = wb[1] * p1 + wb[4], q2 = wb[2] * p2 + wb[5];
      if (q0 >= -1.0f && q0 <= 1.0f) a0 += s0 * wb[0];
      if (q1 >= -1.0f && q1 <= 1.0f) a1 += s1 * wb[1];
      if (q2 >= -1.0f && q2 <= 1.0f) a2 += s2 * wb[2];
    }
    float* part = parts + ((size_t)g * PP + t) * 3;
    part[0] = a0; part[1] = a1; part[2] = a2;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(reinterpret_cast<int*>(ws + L.off_bwd_cnt) + 1 + blockIdx.x, 1) == groups - 1);
  __syncthreads();
  if (!s_last || t >= PP) return;
  __threadfence();
  float a[3] = {0.0f, 0.0f, 0.0f};
  if (accumulate) { a[0] = grad_patch[(size_t)t * 3]; a[1] = grad_patch[(size_t)t * 3 + 1]; a[2] = grad_patch[(size_t)t * 3 + 2]; }
  for (int gg = 0; gg < groups; ++gg) {
    const float* part = parts + ((size_t)gg * PP + t) * 3;
    a[0] += __ldcg(part); a[1] += __ldcg(part + 1); a[2] += __ldcg(part + 2);
  }
  grad_patch[(size_t)t * 3] = a[0]; grad_patch[(size_t)t * 3 + 1] = a[1]; grad_patch[(size_t)t * 3 + 2] = a[2];
}

}  // namespace eot

using namespace eot;

extern "C" int eot_apply_bwd(const EotShape* shape, const float* patch, const float* print_wb, const float* grad_images,
                             void* workspace, size_t workspace_bytes, float* grad_patch, int accumulate, void* stream) {
  if (!shape) { set_error("shape is NULL"); return EOT_ERR_NULL_POINTER; }
  if (!patch || !print_wb || !grad_images || !workspace || !grad_patch) {
    set_error("eot_apply_bwd: NULL pointer");
    return EOT_ERR_NULL_POINTER;
  }
  if (shape->batch <= 0 || shape->patch_size <= 0 || shape->num_patches != 1) {
    set_error("eot_apply_bwd: needs the shared-patch shape of the forward call (num_patches == 1)");
    return EOT_ERR_BAD_SHAPE;
  }
  EotShape s = *shape;
  const Layout L = make_layout(s);
  if (workspace_bytes < L.total) { set_error("workspace too small: %zu < %zu", workspace_bytes, L.total); return EOT_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = static_cast<char*>(workspace);
  const int B = s.batch, P = s.patch_size, PP = P * P, n = PP * 3;
  if (s.total_boxes == 0) {
    if (!accumulate) EOT_CHECK_CUDA(cudaMemsetAsync(grad_patch, 0, (size_t)n * sizeof(float), st));
    return EOT_OK;
  }
  StageTimer timer(st, "eot_apply_bwd");
  // per-image dL/dY sums, the work ticket and the texel pass' chunk counters
  EOT_CHECK_CUDA(cudaMemsetAsync(ws + L.off_gy_sum, 0, L.off_oor - L.off_gy_sum, st));
  const BwdSmem m = bwd_smem_plan(s, L);
  if (m.total > 200 * 1024) { set_error("eot_apply_bwd: shared-memory tables too large (P=%d, L=%d)", P, L.lmin); return EOT_ERR_BAD_SHAPE; }
  static thread_local size_t attr_set = 0;                        // raise the dynamic shared-memory limit once per size
  if (m.total > 32 * 1024 && m.total > attr_set) {
    EOT_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_image, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.total));
    attr_set = m.total;
  }
  int per_sm = EOT_BWD_MINB;
  while (per_sm > 1 && (m.total + 1024) * per_sm > 220 * 1024) --per_sm;
  const int nstrips = (P + m.rows - 1) / m.rows;
  const long long items = (long long)B * nstrips;
  const int grid = (int)(items < (long long)sm_count() * per_sm ? items : (long long)sm_count() * per_sm);
  k_bwd_image<<<grid, kThreads, m.total, st>>>(s, L, ws, grad_images, patch, print_wb);
  timer.mark("image");
  const int groups = B < kBwdGroups ? B : kBwdGroups;
  EOT_CHECK_CUDA(launch_pdl(k_bwd_texel, dim3((PP + kThreads - 1) / kThreads, groups), dim3(kThreads), 0, st, s, L, ws, patch, print_wb,
                            groups, grad_patch, accumulate));
  timer.mark("texel");
  count_launches(2);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
