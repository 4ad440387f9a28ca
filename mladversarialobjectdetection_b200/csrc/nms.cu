// First-pass post-processing on the device: person candidates -> per-image (soft-)NMS -> clipped ragged boxes
// (reference: attacker.py:100-116,143-170; tf2/postprocess.py:159-205 -> tf.raw_ops.NonMaxSuppressionV5;
//  tf2/postprocess.py clip_boxes; tf2/anchors.py:44-58 decode).  SURVEY.md section 8 row f1.
//
// The reference runs this as a host-synchronous map_fn over the images with TF's CPU-only NMS kernel.  Here one CTA
// owns one image:
//   1. ordered compaction of the candidates the score kernel left in cand_score[b][:] (person arg-max & valid box,
//      >= score_floor, > NMS score threshold); anchor order is kept because NonMaxSuppressionV5 breaks score ties by
//      the lower input index
//   2. decode of the candidates' boxes from the regression levels and the anchor table
//   3. the selection loop of DoNonMaxSuppressionOp (tensorflow/core/kernels/image/non_max_suppression_op.cc), kept in
//      its LAZY form so that the float32 score products are taken in TF's order: pop the best (score, lower index)
//      candidate, rescore it against the boxes selected since its last visit (newest first), select it if its score
//      did not move, otherwise push it back.  The priority queue is a warp-wide arg-max over the live candidates
//      (tens to hundreds after the score threshold); the IoU weights of one visit are computed by the lanes in
//      parallel, only the product chain is serial.
//   4. clip_boxes + padded outputs; a second one-CTA kernel turns valid_len[B] into CSR row splits and compacts the
//      boxes -- the ragged tensor the patcher consumes, without a host round trip per image.
#include "eot_common.cuh"

#include <math.h>

namespace eot {

constexpr int kNmsThreads = 256;
constexpr int kNmsSmemCand = 1024;   // candidates per image kept in shared memory (more: the global workspace)
constexpr int kNmsMaxOut = 128;      // max_output_size supported (EfficientDet: 100)

struct NmsLayout {
  size_t off_idx;     // int32[B][K]  anchor index of candidate k
  size_t off_score;   // float[B][K]
  size_t off_begin;   // int32[B][K]  suppress_begin_index
  size_t off_box;     // float4[B][K]
  size_t total;
  int K;
};

__host__ __device__ inline NmsLayout nms_layout(const NmsShape& s) {
  NmsLayout L;
  const size_t B = (size_t)s.batch;
  const int K = (s.max_candidates > 0 && s.max_candidates < s.total_anchors) ? s.max_candidates : s.total_anchors;
  L.K = K;
  size_t o = 0;
  L.off_idx = o;   o = align_up(o + B * K * 4, 256);
  L.off_score = o; o = align_up(o + B * K * 4, 256);
  L.off_begin = o; o = align_up(o + B * K * 4, 256);
  L.off_box = o;   o = align_up(o + B * K * 16, 256);
  L.total = o;
  return L;
}

struct NmsLevels {
  const float* box[SCORE_MAX_LEVELS];
  int n_anchors[SCORE_MAX_LEVELS];
  int anchor_base[SCORE_MAX_LEVELS + 1];
};

// IOU of non_max_suppression_op.cc (corner order normalised, zero for empty boxes)
__device__ __forceinline__ float nms_iou(const float4 a, const float4 b) {
  const float ymin_a = fminf(a.x, a.z), xmin_a = fminf(a.y, a.w), ymax_a = fmaxf(a.x, a.z), xmax_a = fmaxf(a.y, a.w);
  const float ymin_b = fminf(b.x, b.z), xmin_b = fminf(b.y, b.w), ymax_b = fmaxf(b.x, b.z), xmax_b = fmaxf(b.y, b.w);
  const float area_a = (ymax_a - ymin_a) * (xmax_a - xmin_a);
  const float area_b = (ymax_b - ymin_b) * (xmax_b - xmin_b);
  if (area_a <= 0.0f || area_b <= 0.0f) return 0.0f;
  const float ih = fmaxf(fminf(ymax_a, ymax_b) - fmaxf(ymin_a, ymin_b), 0.0f);
  const float iw = fmaxf(fminf(xmax_a, xmax_b) - fmaxf(xmin_a, xmin_b), 0.0f);
  const float inter = ih * iw;
  return inter / ((area_a + area_b) - inter);
}

// float32 exp rounded from the float64 result (what a correctly rounded expf returns; the host oracle does the same)
__device__ __forceinline__ float exp_f32(float x) { return (float)exp((double)x); }

__global__ void __launch_bounds__(kNmsThreads) k_person_nms(NmsShape s, NmsLayout L, NmsLevels lv,
                                                            const float* __restrict__ cand_score,
                                                            const float* __restrict__ anchors, float* nms_boxes,
                                                            float* nms_scores, int32_t* valid_len, char* ws) {
  __shared__ int s_wcount[kNmsThreads / 32];
  __shared__ int s_n;
  __shared__ float s_score[kNmsSmemCand];
  __shared__ int s_begin[kNmsSmemCand];
  __shared__ float4 s_box[kNmsSmemCand];
  __shared__ float4 s_selbox[kNmsMaxOut];
  __shared__ float s_selscore[kNmsMaxOut];
  __shared__ float s_w[kNmsMaxOut];
  __shared__ int s_count;

  const int b = blockIdx.x, A = s.total_anchors, K = L.K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = kNmsThreads / 32;
  const float* cs = cand_score + (size_t)b * A;
  int* g_idx = reinterpret_cast<int*>(ws + L.off_idx) + (size_t)b * K;
  float* g_score = reinterpret_cast<float*>(ws + L.off_score) + (size_t)b * K;
  int* g_begin = reinterpret_cast<int*>(ws + L.off_begin) + (size_t)b * K;
  float4* g_box = reinterpret_cast<float4*>(ws + L.off_box) + (size_t)b * K;
  // candidate <=> passes attacker.py:87-88 (score >= score_thresh, done on the score the kernel stored) and the
  // priority-queue admission of NonMaxSuppressionV5 (score > score_threshold)
  const float floor_ = s.score_floor;
  const float thr = s.score_threshold;

  // ---- 1. ordered compaction: warp w owns the contiguous anchor range [w*seg, (w+1)*seg); a lane reads four
  //         consecutive scores per step (128-bit loads, no cross-lane traffic unless a step holds a candidate) ----
  const bool vec4 = (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(cs) & 15) == 0);
  const int step = vec4 ? 128 : 32;
  const int seg = ((A + NW - 1) / NW + step - 1) / step * step;
  const int a_lo = min(A, warp * seg), a_hi = min(A, a_lo + seg);
  int cnt = 0;
  if (vec4) {
#pragma unroll 4
    for (int a = a_lo + 4 * lane; a < a_hi; a += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(cs + a));
      cnt += (int)(v.x >= floor_ && v.x > thr) + (int)(v.y >= floor_ && v.y > thr) + (int)(v.z >= floor_ && v.z > thr) +
             (int)(v.w >= floor_ && v.w > thr);
    }
  } else {
    for (int a = a_lo + lane; a < a_hi; a += 32) {
      const float v = __ldg(cs + a);
      cnt += (int)(v >= floor_ && v > thr);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_wcount[warp] = cnt;
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) { if (w < warp) base += s_wcount[w]; total += s_wcount[w]; }
  if (threadIdx.x == 0) s_n = min(total, K);
  if (cnt > 0) {                                              // warp-uniform: this warp's range holds candidates
    for (int a0 = a_lo; a0 < a_hi; a0 += step) {
      float v[4] = {-1.0f, -1.0f, -1.0f, -1.0f};
      const int a = a0 + (vec4 ? 4 * lane : lane);
      if (a < a_hi) {
        if (vec4) { const float4 q = __ldg(reinterpret_cast<const float4*>(cs + a)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
        else v[0] = __ldg(cs + a);
      }
      int mine = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) mine += (int)(v[k] >= floor_ && v[k] > thr);
      if (!__any_sync(0xffffffffu, mine > 0)) continue;
      int incl = mine;                                        // inclusive prefix over the lanes = anchor order
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      int pos = base + incl - mine;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (v[k] >= floor_ && v[k] > thr) {
          if (pos < K) { g_idx[pos] = a + k; g_score[pos] = v[k]; }
          ++pos;
        }
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  const int n = s_n;
  const bool overflow = total > K;               // more candidates than the workspace holds: reported, not truncated silently
  const bool in_smem = n <= kNmsSmemCand;
  float* score = in_smem ? s_score : g_score;
  int* begin = in_smem ? s_begin : g_begin;
  float4* box = in_smem ? s_box : g_box;

  // ---- 2. decode (tf2/anchors.py:44-58) ----
  for (int k = threadIdx.x; k < n; k += kNmsThreads) {
    const int a = g_idx[k];
    int l = 0;
    while (l + 1 < s.num_levels && a >= lv.anchor_base[l + 1]) ++l;
    const int a_local = a - lv.anchor_base[l];
    const float4 tb = __ldg(reinterpret_cast<const float4*>(lv.box[l] + ((size_t)b * lv.n_anchors[l] + a_local) * 4));
    const float4 an = __ldg(reinterpret_cast<const float4*>(anchors + (size_t)a * 4));
    const float yca = (an.x + an.z) / 2.0f, xca = (an.y + an.w) / 2.0f;
    const float ha = an.z - an.x, wa = an.w - an.y;
    const float w = exp_f32(tb.w) * wa, h = exp_f32(tb.z) * ha;     // (ty,tx,th,tw) = (x,y,z,w)
    const float yc = tb.x * ha + yca, xc = tb.y * wa + xca;
    box[k] = make_float4(yc - h / 2.0f, xc - w / 2.0f, yc + h / 2.0f, xc + w / 2.0f);
    score[k] = g_score[k];
    begin[k] = 0;
  }
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  // ---- 3. selection loop (warp 0) ----
  if (warp == 0) {
    const bool soft = s.soft_nms_sigma > 0.0f;
    const float scale = soft ? -0.5f / s.soft_nms_sigma : 0.0f;
    const int max_out = min(s.max_output_size, kNmsMaxOut);
    int count = 0;
    while (count < max_out) {
      // priority queue top: highest score, lower index on ties; dead candidates carry score 0 (live scores are > thr >= 0)
      unsigned long long key = 0ull;
      for (int k = lane; k < n; k += 32) {
        const float sc = score[k];
        if (sc > 0.0f) {
          const unsigned long long kk = ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)k);
          key = kk > key ? kk : key;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      if (key == 0ull) break;                                   // queue empty
      const int i = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
      const float original = __uint_as_float((unsigned)(key >> 32));
      const int bg = begin[i];
      const float4 bi = box[i];
      // weights of the boxes selected since the candidate's last visit
      bool hard = false;
      for (int j = bg + lane; j < count; j += 32) {
        const float sim = nms_iou(bi, s_selbox[j]);
        // suppress_weight: exp(scale * sim * sim) when soft or sim <= threshold, else 0
        s_w[j] = (soft || sim <= s.iou_threshold) ? exp_f32((scale * sim) * sim) : 0.0f;
        if (!soft && sim > s.iou_threshold) hard = true;
      }
      hard = __any_sync(0xffffffffu, hard);
      __syncwarp();
      float sc = original;
      if (lane == 0) {
        for (int j = count - 1; j >= bg; --j) {                  // newest selection first, as TF
          sc = sc * s_w[j];
          if (sc <= thr) break;
        }
      }
      sc = __shfl_sync(0xffffffffu, sc, 0);
      __syncwarp();
      if (lane == 0) {
        if (!hard && sc == original) {
          s_selbox[count] = bi;
          s_selscore[count] = sc;
          score[i] = 0.0f;
        } else if (!hard && sc > thr) {
          score[i] = sc;
          begin[i] = count;
        } else {
          score[i] = 0.0f;
        }
      }
      if (!hard && sc == original) ++count;
      __syncwarp();
    }
    if (lane == 0) s_count = count;
  }
  __syncthreads();

  // ---- 4. outputs: gather + clip_boxes (tf2/postprocess.py), zero padding ----
  const int count = s_count;
  const int max_out = s.max_output_size;
  for (int k = threadIdx.x; k < max_out; k += kNmsThreads) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.0f;
    if (k < count) {
      const float4 v = s_selbox[k];
      o = make_float4(clampf(v.x, 0.0f, s.image_height), clampf(v.y, 0.0f, s.image_width),
                      clampf(v.z, 0.0f, s.image_height), clampf(v.w, 0.0f, s.image_width));
      sc = s_selscore[k];
    }
    reinterpret_cast<float4*>(nms_boxes)[(size_t)b * max_out + k] = o;
    nms_scores[(size_t)b * max_out + k] = sc;
  }
  if (threadIdx.x == 0) valid_len[b] = overflow ? -1 : count;
}

// valid_len[B] -> row_splits[B+1]; boxes/scores compacted in image order (one CTA; B is tens to a few thousand).
// A negative valid_len (candidate overflow) is propagated as row_splits[B] = -1.
__global__ void __launch_bounds__(kNmsThreads) k_nms_ragged(int B, int max_out, const int32_t* __restrict__ valid_len,
                                                            const float* __restrict__ nms_boxes,
                                                            const float* __restrict__ nms_scores, int32_t* row_splits,
                                                            float* ragged_boxes, float* ragged_scores) {
  __shared__ int s_part[kNmsThreads];
  __shared__ int s_bad;
  const int T = kNmsThreads;
  const int per = (B + T - 1) / T;
  const int b0 = min(B, (int)threadIdx.x * per), b1 = min(B, b0 + per);
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  int sum = 0;
  for (int b = b0; b < b1; ++b) { const int v = valid_len[b]; if (v < 0) s_bad = 1; sum += max(v, 0); }
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < T; d <<= 1) {
    const int v = (int)threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
  for (int b = b0; b < b1; ++b) {
    row_splits[b] = run;
    const int c = max(valid_len[b], 0);
    for (int k = 0; k < c; ++k) {
      reinterpret_cast<float4*>(ragged_boxes)[run + k] = reinterpret_cast<const float4*>(nms_boxes)[(size_t)b * max_out + k];
      if (ragged_scores) ragged_scores[run + k] = nms_scores[(size_t)b * max_out + k];
    }
    run += c;
  }
  __syncthreads();
  if (threadIdx.x == T - 1) row_splits[B] = s_bad ? -1 : s_part[T - 1];
}

static int check_nms_shape(const NmsShape* s) {
  if (!s) { set_error("NmsShape is NULL"); return EOT_ERR_NULL_POINTER; }
  if (s->batch <= 0 || s->total_anchors <= 0 || s->num_levels <= 0 || s->num_levels > SCORE_MAX_LEVELS ||
      s->max_output_size <= 0 || s->max_output_size > kNmsMaxOut) {
    set_error("bad NMS shape: batch=%d A=%d levels=%d max_output_size=%d (<= %d)", s->batch, s->total_anchors, s->num_levels,
              s->max_output_size, kNmsMaxOut);
    return EOT_ERR_BAD_SHAPE;
  }
  long long tot = 0;
  for (int l = 0; l < s->num_levels; ++l) tot += s->level_anchors[l];
  if (tot != s->total_anchors) { set_error("bad NMS shape: level_anchors sum to %lld, total_anchors=%d", tot, s->total_anchors); return EOT_ERR_BAD_SHAPE; }
  if (!(s->score_threshold >= 0.0f)) { set_error("person_nms needs score_threshold >= 0 (got %g)", (double)s->score_threshold); return EOT_ERR_BAD_SHAPE; }
  return EOT_OK;
}

}  // namespace eot

using namespace eot;

extern "C" int person_nms_workspace_bytes(const NmsShape* shape, size_t* bytes) {
  if (int rc = check_nms_shape(shape)) return rc;
  if (!bytes) { set_error("bytes is NULL"); return EOT_ERR_NULL_POINTER; }
  *bytes = nms_layout(*shape).total;
  return EOT_OK;
}

extern "C" int person_nms(const NmsShape* shape, const float* cand_score, const float* const* box_levels,
                          const float* anchors, float* nms_boxes, float* nms_scores, int32_t* valid_len,
                          int32_t* row_splits, float* ragged_boxes, float* ragged_scores, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (int rc = check_nms_shape(shape)) return rc;
  if (!cand_score || !box_levels || !anchors || !nms_boxes || !nms_scores || !valid_len || !row_splits || !ragged_boxes ||
      !workspace) {
    set_error("person_nms: NULL pointer");
    return EOT_ERR_NULL_POINTER;
  }
  const NmsLayout L = nms_layout(*shape);
  if (workspace_bytes < L.total) { set_error("workspace too small: %zu < %zu", workspace_bytes, L.total); return EOT_ERR_WORKSPACE_TOO_SMALL; }
  if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return EOT_ERR_MISALIGNED; }
  NmsLevels lv;
  int base = 0;
  for (int l = 0; l < SCORE_MAX_LEVELS; ++l) {
    const bool on = l < shape->num_levels;
    lv.box[l] = on ? box_levels[l] : nullptr;
    lv.n_anchors[l] = on ? shape->level_anchors[l] : 0;
    lv.anchor_base[l] = base;
    if (on && !box_levels[l]) { set_error("person_nms: box level %d is NULL", l); return EOT_ERR_NULL_POINTER; }
    base += lv.n_anchors[l];
  }
  lv.anchor_base[SCORE_MAX_LEVELS] = base;
  cudaStream_t st = (cudaStream_t)stream;
  k_person_nms<<<shape->batch, kNmsThreads, 0, st>>>(*shape, L, lv, cand_score, anchors, nms_boxes, nms_scores, valid_len,
                                                     static_cast<char*>(workspace));
  k_nms_ragged<<<1, kNmsThreads, 0, st>>>(shape->batch, shape->max_output_size, valid_len, nms_boxes, nms_scores, row_splits,
                                          ragged_boxes, ragged_scores);
  count_launches(2);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
