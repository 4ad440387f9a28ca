// Backward of the EOT patch application: dL/d(out_images) -> dL/dpatch
// (reference: tape.gradient at attacker.py:217; chain of SURVEY.md section 3.2 / App. D).
//
//   k_bwd_window   per transformed-patch texel (3 channels): TF's registered gradient of
//                  ImageProjectiveTransformV3 -- the SAME bilinear warp applied to the gradient with
//                  the inverted transform, fill 0 -- reading dL/d(window) through the route bytes the
//                  forward composite left (TensorScatterUpdate / SelectV2 / outer clip routing), then
//                  the inner clip mask of attacker.py:428.  -> g_u[box] (RGBX texels)
//   k_bwd_resize3  exact transpose of the antialiased resize (ScaleAndTranslateGrad) as the forward's
//                  two passes over the transposed weight tables the geometry role wrote; persistent
//                  CTAs over (box, strip of patch rows); per-box partial gradient, no atomics.
//   k_bwd_match    deterministic per-image sum of the boxes' partials, first half of the
//                  BrightnessMatcher backward (clip mask, K'^T), per-image sum of dL/dY.
//   k_bwd_resize   memory-lean serial variant of the two above (EOT_FLAG_SERIAL_ADJOINT): one CTA per
//                  (image, block of patch rows) looping over the image's boxes with a shared-memory
//                  patch-gradient accumulator and a warp-shuffle tree for the dL/dY sum.
//   k_bwd_texel    second half: subtract the per-image mean of dL/dY, K^T, rescale, print-adjust
//                  clip mask and weights; partial sums over image groups.
//   k_bwd_reduce   deterministic sum of the partials (+ optional accumulate).
#include "eot_common.cuh"

namespace eot {

constexpr int kBwdGroups = 16;   // image groups of k_bwd_texel
constexpr int kBwdChunk = 8;     // patch rows per tap fetch in k_bwd_resize

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

#ifndef EOT_BWDW_MINB
#define EOT_BWDW_MINB 4
#endif
#ifndef EOT_BWDR_MINB
#define EOT_BWDR_MINB 3
#endif
__global__ void __launch_bounds__(kThreads, EOT_BWDW_MINB) k_bwd_window(EotShape s, Layout L, char* ws, const float* __restrict__ G) {
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  __shared__ int s_base[kMaxBaseSmem];
  int bstride;
  const int* base = stage_base(reinterpret_cast<const int4*>(ws + L.off_base), kItemBwdWindow, s.total_boxes, s_base, &bstride);
  const int n_items = base[bstride * s.total_boxes];
  const float* ubuf = reinterpret_cast<const float*>(ws + L.off_u);
  float* gubuf = reinterpret_cast<float*>(ws + L.off_gu);
  const uint8_t* routes = reinterpret_cast<const uint8_t*>(ws + L.off_route);
  const int H = s.height, W = s.width;
  const int4* cnt = reinterpret_cast<const int4*>(ws + L.off_cnt);
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int2 item = find_item(base, bstride, s.total_boxes, it);
    const int j = item.x;
    const BoxPlan me = plans[j];
    const int ps = me.ps, D = me.d;
    const int RR = strip_rows(ps, cnt[j].x);
    const float4* u4 = reinterpret_cast<const float4*>(ubuf + me.u_off);
    float4* gu = reinterpret_cast<float4*>(gubuf + (size_t)j * L.gslot);
    const uint8_t* route = routes + (size_t)j * L.rslot;
    const float* Gwin = G + (((size_t)me.image * H + me.y0) * W + me.x0) * 3;
    const int oy0 = item.y * RR;
    const int rows = min(RR, ps - oy0);
    const bool affine = (me.Ti[6] == 0.0f && me.Ti[7] == 0.0f);
    const float Dm1 = (float)(D - 1);
    const int S = u_stride(ps);
    // the strip's texels flattened over the CTA (strips are sized to a multiple of its thread count)
    const float inv_ps = 1.0f / (float)ps;
    for (int t = threadIdx.x; t < rows * ps; t += kThreads) {
      int r = (int)(((float)t + 0.5f) * inv_ps);
      int tx = t - r * ps;
      if (tx < 0) { --r; tx += ps; } else if (tx >= ps) { ++r; tx -= ps; }
      const int ty = oy0 + r;
      const float yf = (float)(ty + me.pad_lo);
      const float cx = me.Ti[1] * yf, cy = me.Ti[4] * yf, cp = me.Ti[7] * yf;
      const float4* urow = u4 + (ty + 2) * S + 2;
      {
        const unsigned bits = __float_as_uint(__ldg(&urow[tx].w));         // inner clip pass bits (attacker.py:428), fetched early
        const float xf = (float)(tx + me.pad_lo);
        float g[3] = {0.0f, 0.0f, 0.0f};
        float ix = (me.Ti[0] * xf + cx) + me.Ti[2];
        float iy = (me.Ti[3] * xf + cy) + me.Ti[5];
        bool ok = true;
        if (!affine) {
          const float proj = (me.Ti[6] * xf + cp) + 1.0f;
          ok = proj != 0.0f;
          if (ok) { ix = ix / proj; iy = iy / proj; }
        }
        if (ok) {
          const float x0f = floorf(ix), y0f = floorf(iy);
          const float wx0 = ix - x0f, wx1 = (x0f + 1.0f) - ix, wy0 = iy - y0f, wy1 = (y0f + 1.0f) - iy;
          float v00[3], v01[3], v10[3], v11[3];
          if (x0f >= 0.0f && x0f < Dm1 && y0f >= 0.0f && y0f < Dm1) {        // all four taps inside the window
            const int xi0 = (int)x0f, yi0 = (int)y0f;
            const int rt = yi0 * D + xi0, px = yi0 * W + xi0;
            // the four route bytes first, then all (predicated) gradient loads: one round trip each, not two per row
            const unsigned r00 = route[rt], r01 = route[rt + 1], r10 = route[rt + D], r11 = route[rt + D + 1];
            routed_grad3_bits(r00, Gwin, px, v00);
            routed_grad3_bits(r01, Gwin, px + 1, v01);
            routed_grad3_bits(r10, Gwin, px + W, v10);
            routed_grad3_bits(r11, Gwin, px + W + 1, v11);
          } else {
            const float x1f = x0f + 1.0f, y1f = y0f + 1.0f, Df = (float)D;
            const bool bx0 = x0f >= 0.0f && x0f < Df, bx1 = x1f >= 0.0f && x1f < Df;
            const bool by0 = y0f >= 0.0f && y0f < Df, by1 = y1f >= 0.0f && y1f < Df;
#pragma unroll
            for (int c = 0; c < 3; ++c) v00[c] = v01[c] = v10[c] = v11[c] = 0.0f;
            if ((bx0 || bx1) && (by0 || by1)) {
              const int xi0 = (int)x0f, yi0 = (int)y0f;
              const int rt = yi0 * D + xi0, px = yi0 * W + xi0;
              if (by0 && bx0) routed_grad3(route, Gwin, rt, px, v00);
              if (by0 && bx1) routed_grad3(route, Gwin, rt + 1, px + 1, v01);
              if (by1 && bx0) routed_grad3(route, Gwin, rt + D, px + W, v10);
              if (by1 && bx1) routed_grad3(route, Gwin, rt + D + 1, px + W + 1, v11);
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) g[c] = wy1 * (wx1 * v00[c] + wx0 * v01[c]) + wy0 * (wx1 * v10[c] + wx0 * v11[c]);
        }
        gu[ty * ps + tx] = make_float4((bits & 1u) ? g[0] : 0.0f, (bits & 2u) ? g[1] : 0.0f, (bits & 4u) ? g[2] : 0.0f, 0.0f);
      }
    }
  }
}

// Exact transpose of the antialiased resize for the boxes of one image, restricted to a block of patch rows.
// Per box the span / inverse-span tables are staged in shared memory once; then, per chunk of patch rows,
//   rows:    tmp[r][f]      = sum_oy w[oy][py - start[oy]] * gu[oy][f]
//   columns: acc[r][px][c] += sum_ox w[ox][px - start[ox]] * tmp[r][ox][c]      (kBwdChunk rows per tap fetch)
__host__ __device__ inline size_t bwd_resize_smem_bytes(const EotShape& s, const Layout& L, int rows_per_cta) {
  return ((size_t)rows_per_cta * s.patch_size * 3 + (size_t)kBwdChunk * L.lmin * 3 + (size_t)L.wcap + (size_t)L.lmin +
          2 * (size_t)s.patch_size) * sizeof(float);
}

__global__ void __launch_bounds__(kThreads) k_bwd_resize(EotShape s, Layout L, char* ws, const float* __restrict__ patch,
                                                         const float* __restrict__ print_wb,
                                                         const int32_t* __restrict__ offsets, int rows_per_cta) {
  extern __shared__ float smem[];
  __shared__ double red[32];
  const int P = s.patch_size, P3 = P * 3;
  const int tstride = L.lmin * 3;
  float* acc_tile = smem;                                          // [rows_per_cta][P3]   patch-gradient accumulator
  float* tmp = acc_tile + (size_t)rows_per_cta * P3;               // [kBwdChunk][lmin*3]
  float* s_w = tmp + (size_t)kBwdChunk * tstride;                  // [ps][span]
  int* s_st = reinterpret_cast<int*>(s_w + L.wcap);                // [ps]
  int2* s_inv = reinterpret_cast<int2*>(s_st + L.lmin);            // [P]
  const int b = blockIdx.y;
  const int py0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, P - py0);
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  const float* gubuf = reinterpret_cast<const float*>(ws + L.off_gu);
  for (int i = threadIdx.x; i < rows_per_cta * P3; i += blockDim.x) acc_tile[i] = 0.0f;
  for (int j = offsets[b]; j < offsets[b + 1]; ++j) {
    const BoxPlan* pl = plans + j;
    if (!pl->valid) continue;
    const int ps = pl->ps, ps3 = ps * 3, span = pl->span;
    const int* starts = reinterpret_cast<const int*>(ws + L.off_starts) + (size_t)j * L.lmin;
    const float* wts = reinterpret_cast<const float*>(ws + L.off_weights) + (size_t)j * L.wcap;
    const int2* inv = reinterpret_cast<const int2*>(ws + L.off_inv) + (size_t)j * P;
    const float* gu = gubuf + (size_t)j * L.gslot;
    __syncthreads();                                               // previous box done with the tables
    for (int i = threadIdx.x; i < ps; i += blockDim.x) s_st[i] = starts[i];
    for (int i = threadIdx.x; i < ps * span; i += blockDim.x) s_w[i] = wts[i];
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_inv[i] = inv[i];
    __syncthreads();
    for (int c0 = 0; c0 < rows; c0 += kBwdChunk) {
      const int crow = min(kBwdChunk, rows - c0);
      for (int f = threadIdx.x; f < ps3; f += blockDim.x) {
        const int fo = (f / 3) * 4 + f % 3;                          // g_u texels are RGBX
        for (int r = 0; r < crow; ++r) {
          const int py = py0 + c0 + r;
          const int2 rng = s_inv[py];
          float a = 0.0f;
          for (int oy = rng.x; oy <= rng.y; ++oy) {
            const int kk = py - s_st[oy];
            if (kk >= 0 && kk < span) a += s_w[kk * ps + oy] * gu[oy * ps * 4 + fo];
          }
          tmp[r * tstride + f] = a;
        }
      }
      __syncthreads();
      for (int f = threadIdx.x; f < P3; f += blockDim.x) {
        const int px = f / 3, c = f - px * 3;
        const int2 rng = s_inv[px];
        float a[kBwdChunk];
#pragma unroll
        for (int r = 0; r < kBwdChunk; ++r) a[r] = 0.0f;
        for (int ox = rng.x; ox <= rng.y; ++ox) {
          const int kk = px - s_st[ox];
          if (kk < 0 || kk >= span) continue;
          const float w = s_w[kk * ps + ox];
          const float* tp = tmp + ox * 3 + c;
#pragma unroll
          for (int r = 0; r < kBwdChunk; ++r) a[r] += w * tp[r * tstride];
        }
#pragma unroll
        for (int r = 0; r < kBwdChunk; ++r)
          if (r < crow) acc_tile[(c0 + r) * P3 + f] += a[r];
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // BrightnessMatcher backward, first half (brightness_matcher.py:65-72 reversed)
  const double* ysum_img = reinterpret_cast<const double*>(ws + L.off_ysum_img);
  const double* ysum_patch = reinterpret_cast<const double*>(ws + L.off_ysum_patch);
  const float mu_t = (float)(ysum_img[b] / (double)((size_t)s.height * s.width));
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
    const float* a = acc_tile + r * P3 + px * 3;
    const float g0 = (rr >= 0.0f && rr <= 1.0f) ? a[0] * EOT_C255_127 : 0.0f;
    const float g1 = (gg >= 0.0f && gg <= 1.0f) ? a[1] * EOT_C255_127 : 0.0f;
    const float g2 = (bb >= 0.0f && bb <= 1.0f) ? a[2] * EOT_C255_127 : 0.0f;
    float gY = g0 + g1 + g2;                                       // K' row Y = (1,1,1)
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

// Fully parallel variant: the transpose of the antialiased resize has the same shape as the forward resize once the
// weights are stored transposed (the geometry role does that: off_wt / off_stt), so it runs the same two passes:
//   rows:    inter[r][ox]    = sum_k wT[py][k] * g_u[st(py)+k][ox]          (RGBX texels, 128-bit loads)
//   columns: gbox[py][px][c] = sum_k wT[px][k] * inter[r][st(px)+k][c]
// One CTA per (box, strip of patch rows) writes that box's share of dL/d(matched patch) to its own slot;
// k_bwd_match then sums an image's boxes in a fixed order (no atomics, deterministic).
__host__ __device__ inline size_t bwd_resize3_smem_bytes(const EotShape& s, const Layout& L) {
  return (size_t)2560 * 16 + (size_t)L.lmin * 16 + (size_t)s.patch_size * L.tcap * 4 + (size_t)(s.patch_size + 1) * 8;
}

// TK > 0: compile-time tap count (every load of a texel issued before the first use; taps past the true count carry
// weight 0 and a clamped index); TK == 0: run-time count.
template <int TK>
__device__ __forceinline__ void bwd_resize_passes(int P, int ps, int tcap, int py0, int rows, const float4* __restrict__ gu,
                                                  const float* s_wt, const int2* s_st, float4* inter, float* gbox) {
  constexpr int NT = TK > 0 ? TK : 1;
  // rows pass, flattened over the strip's (patch row, texel column) pairs; two pairs per thread and iteration so that
  // 2 x NT gradient texels are in flight before the first use (the loads come from L2 / DRAM: g_u is 56 MB)
  const int total = rows * ps;
  const float inv_ps = 1.0f / (float)ps;
  for (int idx = threadIdx.x; idx < total; idx += 2 * blockDim.x) {
    int id[2], rr[2], oxx[2];
    id[0] = idx;
    id[1] = idx + blockDim.x < total ? idx + blockDim.x : idx;      // tail: recompute the first pair (same value stored twice)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      int r = (int)(((float)id[u] + 0.5f) * inv_ps);
      int ox = id[u] - r * ps;
      if (ox < 0) { --r; ox += ps; } else if (ox >= ps) { ++r; ox -= ps; }
      rr[u] = r; oxx[u] = ox;
    }
    if (TK > 0) {
      float4 v[2][NT];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int st = s_st[py0 + rr[u]].x;
#pragma unroll
        for (int k = 0; k < NT; ++k) v[u][k] = gu[(size_t)min(st + k, ps - 1) * ps + oxx[u]];
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float* w = s_wt + (py0 + rr[u]) * tcap;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int k = 0; k < NT; ++k) { const float wk = w[k]; a0 += wk * v[u][k].x; a1 += wk * v[u][k].y; a2 += wk * v[u][k].z; }
        inter[rr[u] * ps + oxx[u]] = make_float4(a0, a1, a2, 0.0f);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int2 sc = s_st[py0 + rr[u]];
        const float* w = s_wt + (py0 + rr[u]) * tcap;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        for (int k = 0; k < sc.y; ++k) {
          const float4 v = gu[(size_t)(sc.x + k) * ps + oxx[u]];
          const float wk = w[k];
          a0 += wk * v.x; a1 += wk * v.y; a2 += wk * v.z;
        }
        inter[rr[u] * ps + oxx[u]] = make_float4(a0, a1, a2, 0.0f);
      }
    }
  }
  __syncthreads();
  const int P3 = P * 3;
  for (int idx = threadIdx.x; idx < rows * P; idx += blockDim.x) {  // columns pass
    const int r = idx / P, px = idx - r * P;
    const int2 sc = s_st[px];
    const float* w = s_wt + px * tcap;
    const float4* irow = inter + r * ps;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    if (TK > 0) {
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const float4 v = irow[min(sc.x + k, ps - 1)];
        const float wk = w[k];
        a0 += wk * v.x; a1 += wk * v.y; a2 += wk * v.z;
      }
    } else {
      for (int k = 0; k < sc.y; ++k) {
        const float4 v = irow[sc.x + k];
        const float wk = w[k];
        a0 += wk * v.x; a1 += wk * v.y; a2 += wk * v.z;
      }
    }
    float* o = gbox + (size_t)(py0 + r) * P3 + px * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
  }
}

__global__ void __launch_bounds__(kThreads, EOT_BWDR_MINB) k_bwd_resize3(EotShape s, Layout L, char* ws) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_base[kMaxBaseSmem];
  __shared__ int2 s_item;
  int bstride;
  const int P = s.patch_size, P3 = P * 3;
  const int inter_texels = max(2560, L.lmin);                      // rows * ps <= 2560 unless a single row is longer
  float4* inter = reinterpret_cast<float4*>(smem);                // [rows][ps]
  int2* s_st = reinterpret_cast<int2*>(smem + (size_t)inter_texels * 4);   // [P+1]
  float* s_wt = reinterpret_cast<float*>(s_st + (P + 1));          // [P][tstride]
  const int* base = stage_base(reinterpret_cast<const int4*>(ws + L.off_base), kItemBwdResize, s.total_boxes, s_base, &bstride);
  const int n_items = base[bstride * s.total_boxes];
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    if (threadIdx.x == 0) s_item = find_item(base, bstride, s.total_boxes, it);
    __syncthreads();
    const int j = s_item.x;
    const BoxPlan* pl = reinterpret_cast<const BoxPlan*>(ws + L.off_plans) + j;
    const int ps = pl->ps;
    const int2* stt = reinterpret_cast<const int2*>(ws + L.off_stt) + (size_t)j * (P + 1);
    for (int i = threadIdx.x; i <= P; i += blockDim.x) s_st[i] = stt[i];
    __syncthreads();
    const int tstride = s_st[P].x, RR = s_st[P].y;
    const int py0 = s_item.y * RR;
    const int rows = min(RR, P - py0);
    const float* wt = reinterpret_cast<const float*>(ws + L.off_wt) + (size_t)j * P * L.tcap;
    const float4* gu = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ws + L.off_gu) + (size_t)j * L.gslot);
    float* gbox = reinterpret_cast<float*>(ws + L.off_gbox) + (size_t)j * P * P3;
    for (int i = threadIdx.x; i < P * tstride; i += blockDim.x) s_wt[i] = wt[i];
    __syncthreads();
    if (tstride == 3) bwd_resize_passes<3>(P, ps, tstride, py0, rows, gu, s_wt, s_st, inter, gbox);
    else if (tstride == 4) bwd_resize_passes<4>(P, ps, tstride, py0, rows, gu, s_wt, s_st, inter, gbox);
    else if (tstride == 5) bwd_resize_passes<5>(P, ps, tstride, py0, rows, gu, s_wt, s_st, inter, gbox);
    else if (tstride == 6) bwd_resize_passes<6>(P, ps, tstride, py0, rows, gu, s_wt, s_st, inter, gbox);
    else bwd_resize_passes<0>(P, ps, tstride, py0, rows, gu, s_wt, s_st, inter, gbox);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) k_bwd_match(EotShape s, Layout L, char* ws, const float* __restrict__ patch,
                                                        const float* __restrict__ print_wb,
                                                        const int32_t* __restrict__ offsets) {
  __shared__ double red[32];
  const int P = s.patch_size, PP = P * P;
  const int b = blockIdx.y;
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  const float* gbox = reinterpret_cast<const float*>(ws + L.off_gbox);
  const double* ysum_img = reinterpret_cast<const double*>(ws + L.off_ysum_img);
  const double* ysum_patch = reinterpret_cast<const double*>(ws + L.off_ysum_patch);
  const float mu_t = (float)(ysum_img[b] / (double)((size_t)s.height * s.width));
  const float mu_s = (float)(ysum_patch[b] / (double)PP);
  const float* wb = print_wb + (size_t)b * 6;
  float* gm = reinterpret_cast<float*>(ws + L.off_gm) + (size_t)b * PP * 3;
  const int j0 = offsets[b], j1 = offsets[b + 1];
  double gy_acc = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < PP; t += gridDim.x * blockDim.x) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int j = j0; j < j1; j += 4) {                             // four boxes' loads in flight, summed in box order
      float g[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = j + u;
        const bool ok = jj < j1 && plans[min(jj, j1 - 1)].valid;
        const float* gp = gbox + ((size_t)min(jj, j1 - 1) * PP + t) * 3;
        g[u][0] = ok ? __ldcg(gp) : 0.0f;
        g[u][1] = ok ? __ldcg(gp + 1) : 0.0f;
        g[u][2] = ok ? __ldcg(gp + 2) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a0 += g[u][0]; a1 += g[u][1]; a2 += g[u][2]; }
    }
    const float* p = patch + (size_t)t * 3;
    const TexelYuv y = texel_yuv(__ldg(p), __ldg(p + 1), __ldg(p + 2), wb);
    const float y_pre = (y.y - mu_s) + mu_t;
    const float yp = clampf(y_pre, 0.0f, 1.0f);
    const float rr = (yp * 1.0f + y.u * EOT_I10) + y.v * EOT_I20;
    const float gg = (yp * 1.0f + y.u * EOT_I11) + y.v * EOT_I21;
    const float bb = (yp * 1.0f + y.u * EOT_I12) + y.v * EOT_I22;
    const float g0 = (rr >= 0.0f && rr <= 1.0f) ? a0 * EOT_C255_127 : 0.0f;
    const float g1 = (gg >= 0.0f && gg <= 1.0f) ? a1 * EOT_C255_127 : 0.0f;
    const float g2 = (bb >= 0.0f && bb <= 1.0f) ? a2 * EOT_C255_127 : 0.0f;
    float gY = g0 + g1 + g2;                                       // K' row Y = (1,1,1)
    const float gU = g0 * EOT_I10 + g1 * EOT_I11 + g2 * EOT_I12;
    const float gV = g0 * EOT_I20 + g1 * EOT_I21 + g2 * EOT_I22;
    if (!(y_pre >= 0.0f && y_pre <= 1.0f)) gY = 0.0f;
    gy_acc += (double)gY;
    float* o = gm + (size_t)t * 3;
    o[0] = gY; o[1] = gU; o[2] = gV;
  }
  gy_acc = block_sum(gy_acc, red);
  if (threadIdx.x == 0) atomicAdd(reinterpret_cast<double*>(ws + L.off_gy_sum) + b, gy_acc);
}

__global__ void __launch_bounds__(kThreads) k_bwd_texel(EotShape s, Layout L, char* ws, const float* __restrict__ patch,
                                                        const float* __restrict__ print_wb, int groups) {
  const int P = s.patch_size, PP = P * P;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= PP) return;
  const int g = blockIdx.y;
  const float p0 = __ldg(patch + (size_t)t * 3), p1 = __ldg(patch + (size_t)t * 3 + 1), p2 = __ldg(patch + (size_t)t * 3 + 2);
  const float* gmb = reinterpret_cast<const float*>(ws + L.off_gm);
  const double* gy_sum = reinterpret_cast<const double*>(ws + L.off_gy_sum);
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
  for (int b = g; b < s.batch; b += groups) {
    const float* wb = print_wb + (size_t)b * 6;
    const float* gm = gmb + ((size_t)b * PP + t) * 3;
    const float mean_gy = (float)(gy_sum[b] / (double)PP);
    const float gYs = gm[0] - mean_gy;                              // d(-mean(Ys)) term
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
  float* part = reinterpret_cast<float*>(ws + L.off_gp_part) + ((size_t)g * PP + t) * 3;
  part[0] = a0; part[1] = a1; part[2] = a2;
}

__global__ void __launch_bounds__(kThreads) k_bwd_reduce(const float* __restrict__ part, int n, int groups,
                                                         float* grad_patch, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = accumulate ? grad_patch[i] : 0.0f;
  for (int g = 0; g < groups; ++g) a += part[(size_t)g * n + i];
  grad_patch[i] = a;
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
  const int32_t* offsets = reinterpret_cast<const int32_t*>(ws + L.off_offsets);
  EOT_CHECK_CUDA(cudaMemsetAsync(ws + L.off_gy_sum, 0, (size_t)B * sizeof(double), st));
  const int nsm = sm_count();
  k_bwd_window<<<nsm * 8, kThreads, 0, st>>>(s, L, ws, grad_images);
  if (L.use_gbox) {
    const size_t smem2 = bwd_resize3_smem_bytes(s, L);
    if (smem2 > 200 * 1024) { set_error("eot_apply_bwd: shared-memory tile too large (L=%d)", L.lmin); return EOT_ERR_BAD_SHAPE; }
    if (smem2 > 32 * 1024) EOT_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_resize3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    k_bwd_resize3<<<nsm * EOT_BWDR_MINB, kThreads, smem2, st>>>(s, L, ws);
    const int mchunks = max(1, min((PP + kThreads - 1) / kThreads, 64));
    k_bwd_match<<<dim3(mchunks, B), kThreads, 0, st>>>(s, L, ws, patch, print_wb, offsets);
    count_launches(1);
  } else {
    // patch rows per CTA: enough CTAs to fill the GPU about twice, a multiple of the chunk
    const int pmax = ((P + kBwdChunk - 1) / kBwdChunk) * kBwdChunk;
    int rpc = (int)(((long long)B * P) / (2ll * nsm));
    rpc = rpc < kBwdChunk ? kBwdChunk : (rpc / kBwdChunk) * kBwdChunk;
    if (rpc > 64) rpc = 64;
    if (rpc > pmax) rpc = pmax;
    size_t smem = bwd_resize_smem_bytes(s, L, rpc);
    while (smem > 100 * 1024 && rpc > kBwdChunk) {
      rpc -= kBwdChunk;
      smem = bwd_resize_smem_bytes(s, L, rpc);
    }
    if (smem > 200 * 1024) { set_error("eot_apply_bwd: shared-memory tile too large (P=%d, L=%d)", P, L.lmin); return EOT_ERR_BAD_SHAPE; }
    if (smem > 48 * 1024) EOT_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_resize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bwd_resize<<<dim3((P + rpc - 1) / rpc, B), kThreads, smem, st>>>(s, L, ws, patch, print_wb, offsets, rpc);
  }
  const int groups = B < kBwdGroups ? B : kBwdGroups;
  k_bwd_texel<<<dim3((PP + kThreads - 1) / kThreads, groups), kThreads, 0, st>>>(s, L, ws, patch, print_wb, groups);
  k_bwd_reduce<<<(n + kThreads - 1) / kThreads, kThreads, 0, st>>>(reinterpret_cast<const float*>(ws + L.off_gp_part), n, groups,
                                                                    grad_patch, accumulate);
  count_launches(4);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
