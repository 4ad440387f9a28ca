// Device side of the forward resize (shared by k_resize2 and the fused forward kernel); see eot_resize.cu.
#pragma once
#include "eot_common.cuh"

namespace eot {


__host__ __device__ inline size_t resize_warp_smem(const EotShape& s, const Layout& L) {
  // intermediate rows + the Philox words of one chunk per row of the item
  return (size_t)L.rb * s.patch_size * 16 + (size_t)L.rb * kNoiseWords * 4;
}

// SPAN == 2: two-tap table; SPAN > 2: compile-time span from the tap-major weight table; SPAN == 0: run-time span.
template <int SPAN>
__device__ __forceinline__ void resize_item(const EotShape& s, const Layout& L, char* ws, const BoxPlan* __restrict__ pl, int j,
                                            int blk, float4* inter, uint32_t* words, float one, int lane) {
  constexpr int NS = SPAN > 2 ? SPAN : 1;
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
#if EOT_PACKED_MATH
        irow[x] = lerp2_texel(a0, b0, wa, wb, one);
        irow[x2] = lerp2_texel(a1, b1, wa, wb, one);
#else
        irow[x] = make_float4(wa * a0.x + wb * b0.x, wa * a0.y + wb * b0.y, wa * a0.z + wb * b0.z, 0.0f);
        irow[x2] = make_float4(wa * a1.x + wb * b1.x, wa * a1.y + wb * b1.y, wa * a1.z + wb * b1.z, 0.0f);
#endif
      }
    } else if (SPAN > 2) {
      const int st = __ldg(starts + oy);
      float w[NS];
      int ro[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) { w[k] = __ldg(wts + k * ps + oy); ro[k] = min(st + k, P - 1) * P; }
      for (int x = lane; x < P; x += 32) {
        float4 v[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) v[k] = m4[ro[k] + x];
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int k = 0; k < NS; ++k) { a0 = a0 + w[k] * v[k].x; a1 = a1 + w[k] * v[k].y; a2 = a2 + w[k] * v[k].z; }
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
  // ---- columns pass + noise + delta + clip, chunks of kNoiseChunk output columns over all rows of the item ----
  const float amp = s.noise_amp;
  const float nlo = -amp, nrng = amp - nlo;                         // TF random_uniform range map: u * (hi - lo) + lo
  for (int c0 = 0; c0 < ps; c0 += kNoiseChunk) {
    const int cn = min(kNoiseChunk, ps - c0);                       // columns of this chunk
    // Philox words of the chunk's elements, per row: element e = (row * ps + column) * 3 + channel sits in word e & 3 of
    // counter e >> 2; 96 counters (3 rounds of the full warp) serve 128 texels
    for (int r = 0; r < nrows; ++r) {
      const uint32_t e0 = (uint32_t)((r0 + r) * ps + c0) * 3u;
      const uint32_t g0 = e0 >> 2;
      const int nctr = (int)(((e0 + (uint32_t)cn * 3u - 1u) >> 2) - g0) + 1;
      uint4* wr = reinterpret_cast<uint4*>(words + r * kNoiseWords);
      for (int c = lane; c < nctr; c += 32) wr[c] = philox4x32_10(g0 + (uint32_t)c, key0, key1);
    }
    __syncwarp();
    for (int seg = 0; seg < cn; seg += 32) {
      const int oc = seg + lane;                                    // column inside the chunk
      if (oc < cn) {
        const int ox = c0 + oc;
        // the column's taps are shared by the rows of the item
        float4 tt;
        int st = 0;
        float w[NS];
        if (SPAN == 2) {
          tt = __ldg(tab2 + ox);
        } else {
          st = __ldg(starts + ox);
          if (SPAN > 2) {
#pragma unroll
            for (int k = 0; k < NS; ++k) w[k] = __ldg(wts + k * ps + ox);
          }
        }
        for (int r = 0; r < nrows; ++r) {
          const float4* irow = inter + r * P;
          float a0, a1, a2;
          if (SPAN == 2) {
            const float4 va = irow[__float_as_int(tt.x)], vb = irow[__float_as_int(tt.y)];
#if EOT_PACKED_MATH
            const float4 ab = lerp2_texel(va, vb, tt.z, tt.w, one);
            a0 = ab.x; a1 = ab.y; a2 = ab.z;
#else
            a0 = tt.z * va.x + tt.w * vb.x;
            a1 = tt.z * va.y + tt.w * vb.y;
            a2 = tt.z * va.z + tt.w * vb.z;
#endif
          } else if (SPAN > 2) {
            a0 = 0.0f; a1 = 0.0f; a2 = 0.0f;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
              const float4 v = irow[min(st + k, P - 1)];
              a0 = a0 + w[k] * v.x; a1 = a1 + w[k] * v.y; a2 = a2 + w[k] * v.z;
            }
          } else {
            a0 = 0.0f; a1 = 0.0f; a2 = 0.0f;
            for (int k = 0; k < span; ++k) {
              const float wk = __ldg(wts + k * ps + ox);
              const float4 v = irow[min(st + k, P - 1)];
              a0 = a0 + wk * v.x; a1 = a1 + wk * v.y; a2 = a2 + wk * v.z;
            }
          }
          const uint32_t e0 = (uint32_t)((r0 + r) * ps + c0) * 3u;
          const uint32_t* wp = words + r * kNoiseWords + ((e0 & 3u) + 3u * (uint32_t)oc);
          // TF Uint32ToFloat: mantissa bits -> [1,2) - 1
          const float n0 = (__uint_as_float((wp[0] & 0x7FFFFFu) | 0x3F800000u) - 1.0f) * nrng + nlo;
          const float n1 = (__uint_as_float((wp[1] & 0x7FFFFFu) | 0x3F800000u) - 1.0f) * nrng + nlo;
          const float n2 = (__uint_as_float((wp[2] & 0x7FFFFFu) | 0x3F800000u) - 1.0f) * nrng + nlo;
          const float v0 = (a0 + n0) + delta, v1 = (a1 + n1) + delta, v2 = (a2 + n2) + delta;
          const unsigned bits = (unsigned)(fabsf(v0) <= 1.0f) | ((unsigned)(fabsf(v1) <= 1.0f) << 1) |
                                ((unsigned)(fabsf(v2) <= 1.0f) << 2);
          u4[(r0 + r + 2) * S + ox + 2] =
              make_float4(clampf(v0, -1.0f, 1.0f), clampf(v1, -1.0f, 1.0f), clampf(v2, -1.0f, 1.0f), __uint_as_float(bits));
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace eot
