// Device side of the forward composite (shared by k_composite3 / k_composite_rest and the fused forward kernel); see
// eot_composite.cu.
#pragma once
#include "eot_common.cuh"

namespace eot {

// What the sampling of a box needs (loaded per item; the plans of a few hundred boxes stay L1 / L2 resident).
struct SegBox {
  float t0, t1, t2, t3, t4, t5, t6, t7;
  float lo2, hi;              // floor coordinates with a tap inside the core lie strictly between (padded window coordinates)
  int org, S;                 // pad_lo - 2, ps + 4
  int y0, x0, d, image;
  const float4* u;
};

__device__ __forceinline__ SegBox load_segbox(const BoxPlan* __restrict__ o, const float* ubuf) {
  SegBox b;
  const int4 g0 = *reinterpret_cast<const int4*>(&o->y0);        // y0, x0, ps, d
  const float4 ta = *reinterpret_cast<const float4*>(&o->T[0]);
  const float4 tb = *reinterpret_cast<const float4*>(&o->T[4]);
  const int pad_lo = o->pad_lo;
  b.t0 = ta.x; b.t1 = ta.y; b.t2 = ta.z; b.t3 = ta.w; b.t4 = tb.x; b.t5 = tb.y; b.t6 = tb.z; b.t7 = tb.w;
  b.lo2 = (float)(pad_lo - 2);
  b.hi = (float)(pad_lo + g0.z);
  b.org = pad_lo - 2;
  b.S = g0.z + 4;
  b.y0 = g0.x; b.x0 = g0.y; b.d = g0.w;
  b.image = o->image;
  b.u = reinterpret_cast<const float4*>(ubuf + o->u_off);
  return b;
}

// floorf for |v| < 2^22 as two full-rate additions (FRND / F2I run on the quarter-rate pipe): v + 1.5 * 2^23 rounded
// down is floor(v) + 1.5 * 2^23 exactly (the ulp there is 1), the integer sits in the low mantissa bits.  Values outside
// the range (or NaN) come out far outside every window, which is all the range test below needs.
struct Floor { float f; int i; };
__device__ __forceinline__ Floor fast_floor(float v) {
  const float m = __fadd_rd(v, 12582912.0f);
  Floor r;
  r.f = m - 12582912.0f;
  r.i = __float_as_int(m) - 0x4B400000;
  return r;
}

// texel (yi, xi) of the padded window of a box at ub[yi * S + xi]
__device__ __forceinline__ const float4* padded_base(const SegBox& b) { return b.u - (b.org * b.S + b.org); }

// ImageProjectiveTransformV3 (BILINEAR, CONSTANT fill -2) of the ring-padded transformed patch at window pixel
// (xf, .) of a row whose row terms (t1 * y, t4 * y, t7 * y) are cx, cy, cp.  false: no tap can touch the core (the
// sample is exactly -2 on every channel).  For the reference's pure rotation the projective row is zero, proj == 1
// exactly and x / 1 == x, so the two divisions are skipped without changing a bit.
__device__ __forceinline__ bool sample_row_px(const SegBox& me, const float4* ub, float xf, float cx, float cy, float cp,
                                              bool proj_on, float R[3]) {
  float ix = (me.t0 * xf + cx) + me.t2;
  float iy = (me.t3 * xf + cy) + me.t5;
  if (proj_on) {
    const float proj = (me.t6 * xf + cp) + 1.0f;
    if (proj == 0.0f) return false;
    ix = ix / proj;
    iy = iy / proj;
  }
  const Floor fx = fast_floor(ix), fy = fast_floor(iy);
  // at least one tap inside the core <=> floor coordinate in [pad_lo - 1, pad_lo + ps - 1] on both axes
  if (!(fx.f > me.lo2 && fx.f < me.hi && fy.f > me.lo2 && fy.f < me.hi)) return false;
  const float4* p = ub + (fy.i * me.S + fx.i);
  const float4 v00 = p[0], v01 = p[1], v10 = p[me.S], v11 = p[me.S + 1];
  const float wx1 = (fx.f + 1.0f) - ix, wx0 = ix - fx.f, wy1 = (fy.f + 1.0f) - iy, wy0 = iy - fy.f;
  blend3(v00, v01, v10, v11, wx1, wx0, wy1, wy0, R);
  return true;
}

// channels a sample pastes: !(R < -1) (attacker.py:440)
__device__ __forceinline__ unsigned paste_bits(const float R[3]) {
  return (unsigned)!(R[0] < -1.0f) | ((unsigned)!(R[1] < -1.0f) << 1) | ((unsigned)!(R[2] < -1.0f) << 2);
}
// channels a sample passes through the outer clip with gradient 1 (attacker.py:441)
__device__ __forceinline__ unsigned clip_pass_bits(const float R[3]) {
  return (unsigned)(fabsf(R[0]) <= 1.0f) | ((unsigned)(fabsf(R[1]) <= 1.0f) << 1) | ((unsigned)(fabsf(R[2]) <= 1.0f) << 2);
}

// ---- the other windows of the image ----------------------------------------------------------------------------------
// Column range (image coordinates) of box q in image row gy: the row-table range of the rotated core, or the whole
// window row when the background is clipped; empty when the row misses the window or q is j itself / past the image.
struct ColRange { int a, b; };
__device__ __forceinline__ ColRange candidate_range(const BoxPlan* __restrict__ plans, const int2* __restrict__ rowtab, int lfull,
                                                    int q, int j, int last, int gy, bool clip_bg) {
  ColRange r = {1, 0};
  if (q < last && q != j) {
    const BoxPlan* o = plans + q;
    const int4 g = *reinterpret_cast<const int4*>(&o->y0);         // y0, x0, ps, d
    if (o->valid && gy >= g.x && gy < g.x + g.w) {
      const int2 sp = clip_bg ? make_int2(0, g.w - 1) : __ldg(rowtab + (size_t)q * lfull + (gy - g.x));
      if (sp.x <= sp.y) { r.a = g.y + sp.x; r.b = g.y + sp.y; }
    }
  }
  return r;
}

// ---- general row: window columns [xlo, xhi] of window row wy of box j, any number of other boxes -------------------
// Up to 32 boxes per image (`single`): lane l holds the column range of box first + l for the whole row, so the per-step
// ownership test is one ballot over registers; more: the ranges are recomputed per step and 32-box round.
template <bool kMask>
__device__ __noinline__ void composite_row_general(const BoxPlan* __restrict__ plans, const float* ubuf,
                                                   const int2* __restrict__ rowtab, uint8_t* routes, int64_t rslot, int lfull,
                                                   int W, int j, int wy, int xlo, int xhi, bool clip_bg,
                                                   const float* __restrict__ img, float* o_img, float* m_img, int lane) {
  const SegBox me = load_segbox(plans + j, ubuf);
  const int first = plans[j].first_box, last = plans[j].last_box;
  const bool proj_on = me.t6 != 0.0f || me.t7 != 0.0f;
  const float4* ub = padded_base(me);
  const bool single = last - first <= 32;
  const int ctop = first + ((last - 1 - first) / 32) * 32;        // first box of the last 32-box round
  const int gy = me.y0 + wy;
  int2 sp = clip_bg ? make_int2(0, me.d - 1) : __ldg(rowtab + (size_t)j * lfull + wy);
  sp.x = max(sp.x, xlo);                                          // window columns [xlo, xhi] of the row only
  sp.y = min(sp.y, xhi);
  const float yf = (float)wy;
  const float cx = me.t1 * yf, cy = me.t4 * yf, cp = me.t7 * yf;
  uint8_t* rrow = routes + (size_t)j * rslot + (size_t)wy * me.d;
  ColRange mine = {1, 0};
  if (single) mine = candidate_range(plans, rowtab, lfull, first + lane, j, last, gy, clip_bg);
  for (int xs = sp.x; xs <= sp.y; xs += 32) {
    const int x = xs + lane, gx = me.x0 + x;
    bool active = x <= sp.y;
    const int ga = me.x0 + xs, gb = ga + 31;
    // 1. a newer box whose range covers the pixel owns it
    for (int c = single ? first : (j + 1 - first) / 32 * 32 + first; c < last; c += 32) {
      if (!single) mine = candidate_range(plans, rowtab, lfull, c + lane, j, last, gy, clip_bg);
      unsigned m = __ballot_sync(0xffffffffu, c + lane > j && mine.a <= gb && mine.b >= ga);
      while (m) {                                                 // warp-uniform
        const int i = __ffs(m) - 1;
        m &= m - 1;
        const int qa = __shfl_sync(0xffffffffu, mine.a, i), qb = __shfl_sync(0xffffffffu, mine.b, i);
        if (gx >= qa && gx <= qb) active = false;
      }
    }
    if (!__any_sync(0xffffffffu, active)) continue;
    // 2. the owner's own sample
    float v[3] = {0.0f, 0.0f, 0.0f};
    unsigned found = 0;
    if (active) {
      float R[3];
      if (sample_row_px(me, ub, (float)x, cx, cy, cp, proj_on, R)) {
        found = paste_bits(R);
        v[0] = R[0]; v[1] = R[1]; v[2] = R[2];
        const unsigned route = found & clip_pass_bits(R);
        if (route) rrow[x] = (uint8_t)route;                      // the maps start all-zero
      }
    }
    // 3. channels it leaves open: the older boxes covering the pixel, newest first
    unsigned missing = active ? (7u & ~found) : 0u;
    if (__any_sync(0xffffffffu, missing != 0u)) {
      for (int c = single ? first : ctop; c >= first; c -= 32) {
        if (!single) mine = candidate_range(plans, rowtab, lfull, c + lane, j, last, gy, clip_bg);
        unsigned m = __ballot_sync(0xffffffffu, c + lane < j && mine.a <= gb && mine.b >= ga);
        while (m) {                                               // warp-uniform
          const int i = 31 - __clz(m);
          m &= ~(1u << i);
          const int qa = __shfl_sync(0xffffffffu, mine.a, i), qb = __shfl_sync(0xffffffffu, mine.b, i);
          const bool inq = missing != 0u && gx >= qa && gx <= qb;
          if (!__any_sync(0xffffffffu, inq)) continue;
          const int k = c + i;
          const SegBox ob = load_segbox(plans + k, ubuf);
          if (inq) {
            const int xk = gx - ob.x0, yk = gy - ob.y0;
            const float ykf = (float)yk;
            float R[3];
            if (sample_row_px(ob, padded_base(ob), (float)xk, ob.t1 * ykf, ob.t4 * ykf, ob.t7 * ykf,
                              ob.t6 != 0.0f || ob.t7 != 0.0f, R)) {
              const unsigned take = paste_bits(R) & missing;
              if (take) {
                if (take & 1u) v[0] = R[0];
                if (take & 2u) v[1] = R[1];
                if (take & 4u) v[2] = R[2];
                found |= take;
                missing &= ~take;
                const unsigned route = take & clip_pass_bits(R);
                if (route) routes[(size_t)k * rslot + (size_t)yk * ob.d + xk] = (uint8_t)route;
              }
            }
          }
        }
      }
    }
    // 4. store; with a clipped background the owner (= the last window covering the pixel) rewrites the open
    // channels with clip(original)
    const unsigned wr = active ? (clip_bg ? 7u : found) : 0u;
    if (!wr) continue;
    const int e = (gy * W + gx) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (wr & (1u << ch)) {
        const bool mineb = (found >> ch) & 1u;
        const float bg = (!mineb || kMask) ? __ldg(img + e + ch) : 0.0f;
        const float o = clampf(mineb ? v[ch] : bg, -1.0f, 1.0f);
        o_img[e + ch] = o;
        if (kMask) m_img[e + ch] = bg - o;                        // Masker: mask = original - pasted (attack_detection.py:429-430)
      }
  }
}

#ifndef EOT_COMP_MINB
#define EOT_COMP_MINB 4
#endif

// ---- main path ---------------------------------------------------------------------------------------------------------
// k_composite3 (images inside [-1,1], at most 32 boxes per image): per window row the newer boxes' ranges are cut out
// of the row's own range with ballots / warp reductions over the lane-held ranges, and the pixel loop runs over the
// remaining intervals: it carries no ownership test, no call and no shuffle.  An owned pixel that leaves a channel open
// (R < -1: the edge of the own core) inside the hull of the older boxes' ranges is appended to the `open pixel` list;
// k_composite_rest fills those from the older boxes, one lane per pixel, and runs whole rows of everything that is not
// the common case (out-of-range images, crowds of more than 32 boxes) through composite_row_general.
__host__ __device__ inline bool item_is_general(int clip_bg, int first, int last) { return clip_bg != 0 || last - first > 32; }

// Sampling coordinates of one window pixel: source position (ix, iy) in the padded window, its floor (as floats and as
// the mantissa bits of v + 1.5 * 2^23: floor + 0x4B400000), whether a tap can touch the core.
struct PixCoord { float ix, iy, fx, fy; int xi, yi; bool ok; };
template <bool kProj>
__device__ __forceinline__ PixCoord pix_coord(const SegBox& me, float xf, float cx, float cy, float cp, bool in) {
  PixCoord c;
  c.ix = (me.t0 * xf + cx) + me.t2;
  c.iy = (me.t3 * xf + cy) + me.t5;
  c.ok = in;
  if (kProj) {
    const float proj = (me.t6 * xf + cp) + 1.0f;
    c.ok = c.ok && proj != 0.0f;
    c.ix = c.ix / proj;
    c.iy = c.iy / proj;
  }
  // floorf for |v| < 2^22 as two full-rate additions (FRND / F2I run on the quarter-rate pipe)
  const float mx = __fadd_rd(c.ix, 12582912.0f), my = __fadd_rd(c.iy, 12582912.0f);
  c.fx = mx - 12582912.0f;
  c.fy = my - 12582912.0f;
  c.xi = __float_as_int(mx);
  c.yi = __float_as_int(my);
  // at least one tap inside the core <=> floor coordinate in [pad_lo - 1, pad_lo + ps - 1] on both axes
  c.ok = c.ok && c.fx > me.lo2 && c.fx < me.hi && c.fy > me.lo2 && c.fy < me.hi;
  return c;
}

// Window row wy of box j in the common case.  kProj: the box has a projective row (two divisions per pixel).
template <bool kMask, bool kProj>
__device__ __forceinline__ void composite_row_main(const SegBox& me, const BoxPlan* __restrict__ plans,
                                                   const int2* __restrict__ rowtab, uint8_t* route_j, int lfull, int W, int j,
                                                   int wy, const int2 sp, const ColRange mine, bool others, int first, int last,
                                                   const float* __restrict__ img, float* o_img, float* m_img, int* open_count,
                                                   int2* open_list, int open_cap, float one, int lane) {
  if (sp.x > sp.y) return;
  const int S = me.S;
  // texel (yi, xi) of the padded window sits at u[(yi - org) * S + (xi - org)]; with the floor coordinates taken
  // straight from the mantissa bits the constants fold into one 32-bit offset (wrapping arithmetic: the true index is
  // small)
  const int fold = (int)(0u - (0x4B400000u + (unsigned)me.org) * (unsigned)(S + 1));
  const int gy = me.y0 + wy;
  const bool newer = first + lane > j && mine.a <= mine.b;
  // hull of the older boxes' ranges in this image row, in window columns (empty: [1, 0])
  int oa = 1, ob = 0;
  if (others) {
    const bool older = first + lane < j && mine.a <= mine.b;
    oa = __reduce_min_sync(0xffffffffu, older ? mine.a : 0x3fffffff) - me.x0;
    ob = __reduce_max_sync(0xffffffffu, older ? mine.b : -0x3fffffff) - me.x0;
  }
  const float yf = (float)wy;
  const float cx = me.t1 * yf, cy = me.t4 * yf, cp = me.t7 * yf;
  const int erow = (gy * W + me.x0) * 3;
  float* orow = keep_ptr(o_img + erow);                            // (opaque: else re-derived from the kernel parameters per pixel)
  uint8_t* rrow = keep_ptr(route_j + (size_t)wy * me.d);
  int cur = sp.x;
  while (cur <= sp.y) {
    int e = sp.y;
    if (others) {                                                 // cut the newer boxes' ranges out of [cur, sp.y]
      const int gcur = me.x0 + cur;
      const unsigned cov = __ballot_sync(0xffffffffu, newer && mine.a <= gcur && gcur <= mine.b);
      if (cov) {
        cur = __reduce_max_sync(0xffffffffu, (cov >> lane) & 1u ? mine.b : -0x3fffffff) - me.x0 + 1;
        continue;
      }
      const int nxa = __reduce_min_sync(0xffffffffu, newer && mine.a > gcur ? mine.a : 0x3fffffff) - me.x0;
      e = min(e, nxa - 1);
    }
    // pixel loop over the owned columns [cur, e]
    float xf = (float)(cur + lane);
    for (int x = cur + lane; x - lane <= e; x += 32, xf += 32.0f) {
      const PixCoord pc = pix_coord<kProj>(me, xf, cx, cy, cp, x <= e);
      unsigned pasted = 0;
      if (pc.ok) {
        const float4* p = me.u + (pc.yi * S + pc.xi + fold);
        const float4 v00 = p[0], v01 = p[1], v10 = p[S], v11 = p[S + 1];
        const float ix = pc.ix, iy = pc.iy, fx = pc.fx, fy = pc.fy;
        const float wx1 = (fx + 1.0f) - ix, wx0 = ix - fx, wy1 = (fy + 1.0f) - iy, wy0 = iy - fy;
        float R[3];
#if EOT_PACKED_MATH
        blend3_packed(v00, v01, v10, v11, wx1, wx0, wy1, wy0, one, R);
#else
        blend3(v00, v01, v10, v11, wx1, wx0, wy1, wy0, R);
#endif
        const bool p0 = !(R[0] < -1.0f), p1 = !(R[1] < -1.0f), p2 = !(R[2] < -1.0f);   // attacker.py:440
        if (p0 || p1 || p2) {
          pasted = (unsigned)p0 | ((unsigned)p1 << 1) | ((unsigned)p2 << 2);
          const unsigned route = clip_pass_bits(R);               // |R| <= 1 implies R >= -1: a subset of the pasted channels
          if (route) rrow[x] = (uint8_t)route;                    // the maps start all-zero
          float* po = orow + x * 3;
          const float o0 = clampf(R[0], -1.0f, 1.0f), o1 = clampf(R[1], -1.0f, 1.0f), o2 = clampf(R[2], -1.0f, 1.0f);
          if (p0) po[0] = o0;
          if (p1) po[1] = o1;
          if (p2) po[2] = o2;
          if (kMask) {                                            // Masker: mask = original - pasted (attack_detection.py:429-430)
            const float* pi = img + erow + x * 3;
            float* pm = m_img + erow + x * 3;
            if (p0) pm[0] = __ldg(pi) - o0;
            if (p1) pm[1] = __ldg(pi + 1) - o1;
            if (p2) pm[2] = __ldg(pi + 2) - o2;
          }
        }
      }
      // an owned pixel inside the older boxes' hull with an open channel: an older box may fill it (k_composite_rest)
      if (oa <= ob) {
        const bool open = x <= e && x >= oa && x <= ob && pasted != 7u;
        const unsigned m = __ballot_sync(0xffffffffu, open);
        if (m) {
          int at = 0;
          if (lane == 0) at = atomicAdd(open_count, __popc(m));
          at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
          if (open && at < open_cap) open_list[at] = make_int2(j, (wy << 16) | (x << 3) | (int)(7u & ~pasted));
        }
      }
    }
    cur = e + 1;
  }
}

// One work item of the common case: window rows [item_block * kCompRows, ...) of box j.  Lane l looks after box first + l
// of the image (its geometry stays in registers for the item); the row-table entries of every row of the item -- the box's own
// and lane l's candidate's -- are loaded up front, together: one round trip per item instead of two dependent ones per row.
template <bool kMask>
__device__ __forceinline__ void composite_item_main(const EotShape& s, const Layout& L, char* ws, const BoxPlan* __restrict__ plans,
                                                    const float* ubuf, const int2* __restrict__ rowtab, uint8_t* routes,
                                                    const float* __restrict__ images, float* out, float* mask, int j,
                                                    int item_block, int* open_count, int2* open_list, int open_cap, float one,
                                                    int lane) {
  const int W = s.width, lfull = s.height < s.width ? s.height : s.width;
  const size_t img_elems = (size_t)s.height * s.width * 3;
  const int r0 = item_block * kCompRows;
  const SegBox me = load_segbox(plans + j, ubuf);
  const int r1 = min(r0 + kCompRows, me.d);
  const int first = plans[j].first_box, last = plans[j].last_box;
  if (item_is_general(reinterpret_cast<const int*>(ws + L.off_oor)[me.image], first, last)) return;   // k_composite_rest's
  // does the window of box first + lane meet the item's rows at all?
  bool cand = false;
  int4 g = make_int4(0, 0, 0, 0);                                 // y0, x0, ps, d of the lane's box
  const int q = first + lane;
  if (q < last && q != j) {
    const BoxPlan* o = plans + q;
    g = *reinterpret_cast<const int4*>(&o->y0);
    cand = o->valid && g.x < me.y0 + r1 && g.x + g.w > me.y0 + r0 && g.y < me.x0 + me.d && g.y + g.w > me.x0;
  }
  const bool others = __any_sync(0xffffffffu, cand);
  // (loading the box's own ranges earlier, beside its plan, was measured slower: +3.5 us of register pressure)
  int2 sp[kCompRows], csp[kCompRows];
#pragma unroll
  for (int r = 0; r < kCompRows; ++r) {
    const int wy = min(r0 + r, me.d - 1), gy = me.y0 + wy;
    sp[r] = __ldg(rowtab + (size_t)j * lfull + wy);
    csp[r] = make_int2(1, 0);
    if (cand && gy >= g.x && gy < g.x + g.w) csp[r] = __ldg(rowtab + (size_t)q * lfull + (gy - g.x));
  }
  const size_t img_off = (size_t)me.image * img_elems;
#pragma unroll
  for (int r = 0; r < kCompRows; ++r) {
    const int wy = r0 + r;
    if (wy >= r1) break;
    ColRange mine = {1, 0};
    if (csp[r].x <= csp[r].y) { mine.a = g.y + csp[r].x; mine.b = g.y + csp[r].y; }
    if (me.t6 != 0.0f || me.t7 != 0.0f)
      composite_row_main<kMask, true>(me, plans, rowtab, routes + (size_t)j * L.rslot, lfull, W, j, wy, sp[r], mine, others, first, last,
                                      images + img_off, out + img_off, kMask ? mask + img_off : nullptr, open_count, open_list,
                                      open_cap, one, lane);
    else
      composite_row_main<kMask, false>(me, plans, rowtab, routes + (size_t)j * L.rslot, lfull, W, j, wy, sp[r], mine, others, first, last,
                                       images + img_off, out + img_off, kMask ? mask + img_off : nullptr, open_count, open_list,
                                       open_cap, one, lane);
  }
}

// The rest: (1) the open pixels k_composite3 listed, one lane each: the older boxes of the image whose range covers the
// pixel are sampled newest first for the channels still open; (2) whole rows of the items k_composite3 skipped (and of
// every item, should the list have overflowed) through composite_row_general.
// gtid / gthreads: index of the calling thread among the threads sharing the work (whole warps) and their number.
template <bool kMask>
__device__ __forceinline__ void composite_rest_body(const EotShape& s, const Layout& L, char* ws, const float* __restrict__ images,
                                                    float* out, float* mask, const int32_t* __restrict__ offsets, int b0, int b1,
                                                    int group, int ngroups, int gtid, int gthreads) {
  const int lane = gtid & 31;
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  const float* ubuf = reinterpret_cast<const float*>(ws + L.off_u);
  uint8_t* routes = reinterpret_cast<uint8_t*>(ws + L.off_route);
  const int2* rowtab = reinterpret_cast<const int2*>(ws + L.off_rowtab);
  const int2* citems = reinterpret_cast<const int2*>(ws + L.off_citems);
  const int4* base = reinterpret_cast<const int4*>(ws + L.off_base);
  const int* oor = reinterpret_cast<const int*>(ws + L.off_oor);
  const int W = s.width, lfull = s.height < s.width ? s.height : s.width;
  const size_t img_elems = (size_t)s.height * s.width * 3;
  const int open_cap = (int)(L.open_cap / ngroups);
  const int n_open = reinterpret_cast<const int*>(ws + L.off_counters)[8 + group];
  const bool overflow = n_open > open_cap;
  const int2* open_list = reinterpret_cast<const int2*>(ws + L.off_open) + (size_t)group * open_cap;
  if (!overflow) {
    for (int i = gtid; i < n_open; i += gthreads) {
      const int2 rec = open_list[i];
      const int j = rec.x, wy = rec.y >> 16, x = (rec.y >> 3) & 0x1fff;
      unsigned missing = (unsigned)rec.y & 7u;
      const BoxPlan* pj = plans + j;
      const int gy = pj->y0 + wy, gx = pj->x0 + x;
      const size_t img_off = (size_t)pj->image * img_elems;
      const int e = (gy * W + gx) * 3;
      for (int k = j - 1; k >= pj->first_box && missing; --k) {   // older boxes, newest first
        const BoxPlan* o = plans + k;
        const SegBox ob = load_segbox(o, ubuf);                     // (all of the plan's fields in one round trip)
        const int4 g = make_int4(ob.y0, ob.x0, 0, ob.d);            // y0, x0, -, d
        if (!o->valid || gy < g.x || gy >= g.x + g.w) continue;
        const int2 sp = __ldg(rowtab + (size_t)k * lfull + (gy - g.x));
        const int xk = gx - g.y, yk = gy - g.x;
        if (xk < sp.x || xk > sp.y) continue;
        const float ykf = (float)yk;
        float R[3];
        if (!sample_row_px(ob, padded_base(ob), (float)xk, ob.t1 * ykf, ob.t4 * ykf, ob.t7 * ykf, ob.t6 != 0.0f || ob.t7 != 0.0f, R))
          continue;
        const unsigned take = paste_bits(R) & missing;
        if (!take) continue;
        missing &= ~take;
        const unsigned route = take & clip_pass_bits(R);
        if (route) routes[(size_t)k * L.rslot + (size_t)yk * g.w + xk] = (uint8_t)route;   // the maps start all-zero
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          if (take & (1u << ch)) {
            const float o1 = clampf(R[ch], -1.0f, 1.0f);
            out[img_off + e + ch] = o1;
            if (kMask) mask[img_off + e + ch] = __ldg(images + img_off + e + ch) - o1;
          }
      }
    }
  }
  // whole rows of the items outside the common case (k_match flags their presence)
  if (!overflow && reinterpret_cast<const int*>(ws + L.off_counters)[6] == 0) return;
  const int lo = base[min(offsets[b0], s.total_boxes)].w, hi = base[min(offsets[b1], s.total_boxes)].w;
  const int nw = gthreads >> 5;
  for (int it = lo + (gtid >> 5); it < hi; it += nw) {
    const int2 item = citems[it];
    const int j = item.x;
    const BoxPlan* pj = plans + j;
    const int clip_bg = oor[pj->image];
    if (!overflow && !item_is_general(clip_bg, pj->first_box, pj->last_box)) continue;
    const int r0 = item.y * kCompRows, r1 = min(r0 + kCompRows, pj->d);
    const size_t img_off = (size_t)pj->image * img_elems;
    for (int wy = r0; wy < r1; ++wy)
      composite_row_general<kMask>(plans, ubuf, rowtab, routes, L.rslot, lfull, W, j, wy, 0, pj->d - 1, clip_bg != 0,
                                   images + img_off, out + img_off, kMask ? mask + img_off : nullptr, lane);
  }
}


}  // namespace eot
