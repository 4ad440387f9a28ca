// Forward resize of the matched patch to every box's patch side, + noise + brightness delta + clip
// (attacker.py:425-428; tf.image.resize(antialias=True) = ScaleAndTranslate: GatherRows then GatherColumns, each a
// sequential float32 accumulation over the span).
//
// Work item = `rb` consecutive output rows of one box, owned by ONE WARP: the intermediate row of ScaleAndTranslate
// (output row oy over all P source columns) is only ever read by output row oy, so a warp produces it into its own
// shared-memory scratch and consumes it itself -- no CTA barrier, no strip staging, no cross-warp traffic.
//   rows pass     lane = source column (128-bit loads of the RGBX matched patch, coalesced); the row's span start and
//                 tap weights are warp-uniform registers
//   columns pass  lane = output texel of the item's flat texel range (rows may be shorter than a warp: the item is swept
//                 in 32-texel segments across its rows); per-texel tap table read coalesced; taps from the warp's scratch
//   noise         Philox4x32-10 words of a segment (96 elements = 24-25 counter values) are produced by the first 25
//                 lanes into the scratch and picked up by the texel lanes (stride 3 words: conflict-free)
// Items are handed out by an atomic ticket per warp (fetched one item ahead, together with its (box, block) record
// from the item table the geometry role wrote).
//
// Code paths, chosen per box: the two-tap table (up-sampling / unit scale: the bulk of the texels at the reference's
// scale .4), compile-time spans 3 / 5 / 7 / 9 (down-sampling) and a run-time span.
#include "eot_common.cuh"

namespace eot {

__host__ __device__ inline size_t resize_warp_smem(const EotShape& s, const Layout& L) {
  return (size_t)L.rb * s.patch_size * 16 + 32 * 16;     // intermediate rows + Philox words (25 x uint4, padded)
}

// SPAN == 2: two-tap table; SPAN > 2: compile-time span from the tap-major weight table; SPAN == 0: run-time span.
template <int SPAN>
__device__ __forceinline__ void resize_item(const EotShape& s, const Layout& L, char* ws, const BoxPlan* __restrict__ pl, int j,
                                            int blk, float4* inter, uint32_t* words, int lane) {
  const int P = s.patch_size;
  const int ps = pl->ps;
  const int span = SPAN > 2 ? SPAN : pl->span;
  const float delta = pl->delta;
  const uint32_t key0 = pl->key0, key1 = pl->key1;
  const float4* __restrict__ m4 = reinterpret_cast<const float4*>(ws + L.off_match) + (size_t)pl->image * P * P;
  const float4* __restrict__ tab2 = reinterpret_cast<const float4*>(ws + L.off_tab2) + (size_t)j * L.lmin;
  const int* __restrict__ starts = reinterpret_cast<const int*>(ws + L.off_starts) + (size_t)j * L.lmin;
  const float* __restrict__ wts = reinterpret_cast<const float*>(ws + L.off_weights) + (size_t)j * L.wcap;
  float4* u4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(ws + L.off_u) + pl->u_off);
  const int S = u_stride(ps);
  const int r0 = blk * L.rb;
  const int nrows = min(L.rb, ps - r0);
  // ---- rows pass: inter[r][x] = sum_k w[oy][k] * m[start(oy) + k][x] ----
  for (int r = 0; r < nrows; ++r) {
    const int oy = r0 + r;
    float4* irow = inter + r * P;
    if (SPAN == 2) {
      const float4 t = __ldg(tab2 + oy);
      const float4* __restrict__ ra = m4 + __float_as_int(t.x) * P;
      const float4* __restrict__ rb = m4 + __float_as_int(t.y) * P;
      const float wa = t.z, wb = t.w;
      for (int x = lane; x < P; x += 64) {
        const int x2 = min(x + 32, P - 1);                        // tail: recomputes the last column (same value)
        const float4 a0 = ra[x], b0 = rb[x], a1 = ra[x2], b1 = rb[x2];
        irow[x] = make_float4(wa * a0.x + wb * b0.x, wa * a0.y + wb * b0.y, wa * a0.z + wb * b0.z, 0.0f);
        irow[x2] = make_float4(wa * a1.x + wb * b1.x, wa * a1.y + wb * b1.y, wa * a1.z + wb * b1.z, 0.0f);
      }
    } else if (SPAN > 2) {
      const int st = __ldg(starts + oy);
      float w[SPAN > 2 ? SPAN : 1];
      int ro[SPAN > 2 ? SPAN : 1];
#pragma unroll
      for (int k = 0; k < SPAN; ++k) { w[k] = __ldg(wts + k * ps + oy); ro[k] = min(st + k, P - 1) * P; }
      for (int x = lane; x < P; x += 32) {
        float4 v[SPAN > 2 ? SPAN : 1];
#pragma unroll
        for (int k = 0; k < SPAN; ++k) v[k] = m4[ro[k] + x];
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int k = 0; k < SPAN; ++k) { a0 = a0 + w[k] * v[k].x; a1 = a1 + w[k] * v[k].y; a2 = a2 + w[k] * v[k].z; }
        irow[x] = make_float4(a0, a1, a2, 0.0f);
      }
    } else {
      const int st = __ldg(starts + oy);
      for (int x = lane; x < P; x += 32) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        for (int k = 0; k < span; ++k) {
          const float wk = __ldg(wts + k * ps + oy);
          const float4 v = m4[min(st + k, P - 1) * P + x];
          a0 = a0 + wk * v.x; a1 = a1 + wk * v.y; a2 = a2 + wk * v.z;
        }
        irow[x] = make_float4(a0, a1, a2, 0.0f);
      }
    }
  }
  __syncwarp();
  // ---- columns pass + noise + delta + clip over the item's flat texel range ----
  const int ntex = nrows * ps;
  const uint32_t e_base = (uint32_t)(r0 * ps) * 3u;              // noise element index of the item's first texel
  const float amp = s.noise_amp;
  const int rb = L.rb;
  for (int seg = 0; seg < ntex; seg += 32) {
    const uint32_t e0 = e_base + 3u * (uint32_t)seg;
    const uint32_t g0 = e0 >> 2;
    const int nel = 3 * min(32, ntex - seg);
    const int ngroups = (int)(((e0 + (uint32_t)nel - 1u) >> 2) - g0) + 1;
    if (lane < ngroups) reinterpret_cast<uint4*>(words)[lane] = philox4x32_10(g0 + (uint32_t)lane, key0, key1);
    __syncwarp();
    const int t = seg + lane;
    if (t < ntex) {
      int r = 0, ox = t;
      if (ox >= ps) { ox -= ps; r = 1; }
      if (rb > 2) {
        if (ox >= ps) { ox -= ps; r = 2; }
        if (ox >= ps) { ox -= ps; r = 3; }
      }
      const float4* irow = inter + r * P;
      float a0, a1, a2;
      if (SPAN == 2) {
        const float4 tt = __ldg(tab2 + ox);
        const float4 va = irow[__float_as_int(tt.x)], vb = irow[__float_as_int(tt.y)];
        a0 = tt.z * va.x + tt.w * vb.x;
        a1 = tt.z * va.y + tt.w * vb.y;
        a2 = tt.z * va.z + tt.w * vb.z;
      } else if (SPAN > 2) {
        const int st = __ldg(starts + ox);
        float w[SPAN > 2 ? SPAN : 1];
#pragma unroll
        for (int k = 0; k < SPAN; ++k) w[k] = __ldg(wts + k * ps + ox);
        a0 = 0.0f; a1 = 0.0f; a2 = 0.0f;
#pragma unroll
        for (int k = 0; k < SPAN; ++k) {
          const float4 v = irow[min(st + k, P - 1)];
          a0 = a0 + w[k] * v.x; a1 = a1 + w[k] * v.y; a2 = a2 + w[k] * v.z;
        }
      } else {
        const int st = __ldg(starts + ox);
        a0 = 0.0f; a1 = 0.0f; a2 = 0.0f;
        for (int k = 0; k < span; ++k) {
          const float wk = __ldg(wts + k * ps + ox);
          const float4 v = irow[min(st + k, P - 1)];
          a0 = a0 + wk * v.x; a1 = a1 + wk * v.y; a2 = a2 + wk * v.z;
        }
      }
      const uint32_t* wp = words + ((e0 & 3u) + 3u * (uint32_t)lane);
      const float v0 = (a0 + noise_from_word(wp[0], amp)) + delta;
      const float v1 = (a1 + noise_from_word(wp[1], amp)) + delta;
      const float v2 = (a2 + noise_from_word(wp[2], amp)) + delta;
      const unsigned bits = (unsigned)(fabsf(v0) <= 1.0f) | ((unsigned)(fabsf(v1) <= 1.0f) << 1) |
                            ((unsigned)(fabsf(v2) <= 1.0f) << 2);
      u4[(r0 + r + 2) * S + ox + 2] =
          make_float4(clampf(v0, -1.0f, 1.0f), clampf(v1, -1.0f, 1.0f), clampf(v2, -1.0f, 1.0f), __uint_as_float(bits));
    }
    __syncwarp();
  }
}

#ifndef EOT_RESIZE_MINB
#define EOT_RESIZE_MINB 4
#endif
__global__ void __launch_bounds__(kThreads, EOT_RESIZE_MINB) k_resize2(EotShape s, Layout L, char* ws,
                                                                      const int32_t* __restrict__ offsets, int b0, int b1,
                                                                      int ticket_slot) {
  extern __shared__ __align__(16) unsigned char resize_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per_warp = resize_warp_smem(s, L);
  float4* inter = reinterpret_cast<float4*>(resize_smem + (size_t)warp * per_warp);
  uint32_t* words = reinterpret_cast<uint32_t*>(resize_smem + (size_t)warp * per_warp + (size_t)L.rb * s.patch_size * 16);
  const int4* base = reinterpret_cast<const int4*>(ws + L.off_base);
  const int lo = base[min(offsets[b0], s.total_boxes)].z, hi = base[min(offsets[b1], s.total_boxes)].z;
  WarpTickets tk;
  tk.init(ws, L.off_tickets, ticket_slot);
  const int2* items = reinterpret_cast<const int2*>(ws + L.off_items);
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  int it = lo + tk.item(tk.draw(lane));
  int2 item = it < hi ? __ldcg(items + it) : make_int2(0, 0);
  while (it < hi) {
    const int nxt = tk.draw(lane);                    // next item's ticket travels while this one computes
    const BoxPlan* pl = plans + item.x;
    const int mode = pl->two_tap ? 2 : pl->span;
    if (mode == 2) resize_item<2>(s, L, ws, pl, item.x, item.y, inter, words, lane);
    else if (mode == 5) resize_item<5>(s, L, ws, pl, item.x, item.y, inter, words, lane);
    else if (mode == 7) resize_item<7>(s, L, ws, pl, item.x, item.y, inter, words, lane);
    else if (mode == 3) resize_item<3>(s, L, ws, pl, item.x, item.y, inter, words, lane);
    else if (mode == 9) resize_item<9>(s, L, ws, pl, item.x, item.y, inter, words, lane);
    else resize_item<0>(s, L, ws, pl, item.x, item.y, inter, words, lane);
    it = lo + tk.item(nxt);
    if (it < hi) item = __ldcg(items + it);
  }
}

int launch_resize2(const EotShape& s, const Layout& L, char* ws, const int32_t* offsets, int b0, int b1, int ticket_slot,
                   cudaStream_t st) {
  const size_t smem = resize_warp_smem(s, L) * (kThreads / 32);
  if (smem > 200 * 1024) { set_error("eot_apply_fwd: patch side %d needs %zu bytes of shared memory per CTA", s.patch_size, smem); return EOT_ERR_BAD_SHAPE; }
  static thread_local size_t attr_set = 0;                        // raise the dynamic shared-memory limit once per size
  if (smem > 32 * 1024 && smem > attr_set) {
    EOT_CHECK_CUDA(cudaFuncSetAttribute(k_resize2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  int per_sm = EOT_RESIZE_MINB;
  const size_t budget = 220 * 1024;
  while (per_sm > 1 && (smem + 1024) * per_sm > budget) --per_sm;
  k_resize2<<<sm_count() * per_sm, kThreads, smem, st>>>(s, L, ws, offsets, b0, b1, ticket_slot);
  count_launches(1);
  return EOT_OK;
}

}  // namespace eot
