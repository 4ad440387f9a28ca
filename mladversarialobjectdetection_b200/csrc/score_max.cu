// Person-score objective of PatchAttacker.second_pass + loss, and its gradient
// (reference: attacker.py:69-89,118-141,190-193; tf2/postprocess.py:67-79,104-116,136-156;
//  tf2/anchors.py:30-58).
//
//   k_score_fwd       one warp per tile of 32 anchors: the tile's 32 x C logits are streamed into
//                     shared memory with coalesced 64-bit loads, then each lane scans one anchor
//                     (first arg-max class, max logit, tie count).  Anchors whose arg-max is class 0
//                     decode their box, apply the valid-box filter and the sigmoid; the per-image
//                     maximum is a packed 64-bit atomicMax (score bits | ~anchor), after a
//                     warp-shuffle reduction.  Reads every logit exactly once: 4*A*C + 16*A bytes.
//   k_score_finalize  unpack the keys -> max_scores / argmax_anchor
//   k_score_ties      list the anchors attaining the maximum (TF's UnsortedSegmentMax gradient
//                     splits equally among ties)
//   k_score_zero      dense zero fill of dL/dcls for all levels (128-bit stores): 4*A*C bytes
//   k_score_scatter   the few non-zeros: dM * sigmoid' / ties, split over tied classes
//   k_score_scalars   dL/dscale and the data term of the loss
#include "eot_common.cuh"

#include <math.h>

namespace eot {

constexpr int kScoreWarps = 8;
constexpr int kTile = 32;     // anchors per warp tile

struct ScoreLevels {
  const float* cls[SCORE_MAX_LEVELS];
  const float* box[SCORE_MAX_LEVELS];
  float* dcls[SCORE_MAX_LEVELS];
  int n_anchors[SCORE_MAX_LEVELS];     // anchors of the level per image (locs * anchors_per_loc)
  int anchor_base[SCORE_MAX_LEVELS];   // index of the level's first anchor in [0, A)
  int tile_base[SCORE_MAX_LEVELS + 1]; // prefix of tiles per image
};

struct ScoreLayout {
  size_t off_keys;      // uint64[B]
  size_t off_counts;    // int32[B] candidates, int32[B] ties
  size_t off_cand;      // float[B][A]   candidate score or -1
  size_t off_ties;      // int32[B][A]   anchors attaining the max
  size_t total;
};

__host__ __device__ inline ScoreLayout score_layout(const ScoreShape& s) {
  ScoreLayout L;
  const size_t B = (size_t)s.batch, A = (size_t)s.total_anchors;
  size_t o = 0;
  L.off_keys = o;   o = align_up(o + B * 8, 256);
  L.off_counts = o; o = align_up(o + 2 * B * 4, 256);
  L.off_cand = o;   o = align_up(o + B * A * 4, 256);
  L.off_ties = o;   o = align_up(o + B * A * 4, 256);
  L.total = o;
  return L;
}

struct TileRef { const float* src; int b, l, a0, n_valid; };

__device__ __forceinline__ TileRef locate_tile(const ScoreShape& s, const ScoreLevels& lv, int t, int tiles_per_image) {
  TileRef r;
  r.b = t / tiles_per_image;
  int tt = t - r.b * tiles_per_image;
  int l = 0;
  while (l + 1 < s.num_levels && tt >= lv.tile_base[l + 1]) ++l;
  tt -= lv.tile_base[l];
  r.l = l;
  r.a0 = tt * kTile;
  r.n_valid = min(kTile, lv.n_anchors[l] - r.a0);
  r.src = lv.cls[l] + ((size_t)r.b * lv.n_anchors[l] + r.a0) * s.num_classes;
  return r;
}

// asynchronous global -> shared copy of one tile (8-byte cp.async: any even class count, no 16-byte alignment needed)
__device__ __forceinline__ void stage_tile_async(float2* dst, const TileRef& r, int C2, int lane) {
  const float2* src = reinterpret_cast<const float2*>(r.src);
  const int nf2 = r.n_valid * C2;
  const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll 9
  for (int i = lane; i < nf2; i += 32)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + (unsigned)i * 8u), "l"(src + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// One warp owns two tile buffers: the copy of its next tile (32 anchors x C logits, 11.5 KB for C = 90) is in flight
// while the lanes scan the current one out of shared memory, one anchor per lane.
__global__ void __launch_bounds__(kScoreWarps * 32, 1) k_score_fwd(ScoreShape s, ScoreLevels lv, const float* __restrict__ anchors,
                                                                   unsigned long long* keys, int* ncand,
                                                                   float* __restrict__ cand_score, int total_tiles) {
  extern __shared__ __align__(16) float2 tile_smem[];
  const int C = s.num_classes, C2 = C >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* buf0 = tile_smem + (size_t)warp * 2 * kTile * C2;
  const int tiles_per_image = lv.tile_base[s.num_levels];
  const int stride = gridDim.x * kScoreWarps;
  int t = blockIdx.x * kScoreWarps + warp;
  int cur = 0;
  TileRef r;
  if (t < total_tiles) {
    r = locate_tile(s, lv, t, tiles_per_image);
    stage_tile_async(buf0, r, C2, lane);
  }
  for (; t < total_tiles; t += stride) {
    TileRef rn;
    const bool more = t + stride < total_tiles;
    if (more) {
      rn = locate_tile(s, lv, t + stride, tiles_per_image);
      stage_tile_async(buf0 + (size_t)(cur ^ 1) * kTile * C2, rn, C2, lane);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const float2* sm = buf0 + (size_t)cur * kTile * C2;
    const int b = r.b, l = r.l, a0 = r.a0, n_valid = r.n_valid, n_l = lv.n_anchors[l];
    unsigned long long key = 0ull;
    float cs = -1.0f;
    if (lane < n_valid) {
      // tf.argmax returns the FIRST maximal index: class 0 (person) is the arg-max iff x[0] >= max(x[1:]),
      // and then the max logit is x[0] itself -- no index tracking needed.  Two independent chains for ILP.
      const float2* mine = sm + (size_t)lane * C2;
      float2 v = mine[0];
      const float best = v.x;
      float m0 = v.y, m1 = -INFINITY;
      int k = 1;
#pragma unroll 11
      for (; k + 1 < C2; k += 2) {
        const float2 u = mine[k], w = mine[k + 1];
        m0 = fmaxf(fmaxf(m0, u.x), u.y);
        m1 = fmaxf(fmaxf(m1, w.x), w.y);
      }
      if (k < C2) { const float2 u = mine[k]; m0 = fmaxf(fmaxf(m0, u.x), u.y); }
      const float others = fmaxf(m0, m1);
      if (best >= others) {   // person is the arg-max class
        const int a_local = a0 + lane;
        const float4 tb = __ldg(reinterpret_cast<const float4*>(lv.box[l] + ((size_t)b * n_l + a_local) * 4));
        const float4 an = __ldg(reinterpret_cast<const float4*>(anchors + (size_t)(lv.anchor_base[l] + a_local) * 4));
        const float yca = (an.x + an.z) / 2.0f, xca = (an.y + an.w) / 2.0f;
        const float ha = an.z - an.x, wa = an.w - an.y;
        const float w = expf(tb.w) * wa, h = expf(tb.z) * ha;     // (ty,tx,th,tw) = (x,y,z,w)
        const float yc = tb.x * ha + yca, xc = tb.y * wa + xca;
        const float ymin = yc - h / 2.0f, xmin = xc - w / 2.0f, ymax = yc + h / 2.0f, xmax = xc + w / 2.0f;
        const float bh = ymax - ymin, bw = xmax - xmin;
        const bool ok = (bw / s.image_width <= 1.0f) && (bh / s.image_height <= 1.0f) && (bh * bw > s.min_area);
        if (ok) {
          cs = (float)(1.0 / (1.0 + exp(-(double)best)));
          key = ((unsigned long long)__float_as_uint(cs) << 32) |
                (unsigned long long)(0xFFFFFFFFu - (unsigned)(lv.anchor_base[l] + a_local));
        }
      }
      cand_score[(size_t)b * s.total_anchors + lv.anchor_base[l] + a0 + lane] = cs;
    }
    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, key != 0ull));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    if (lane == 0 && cnt) {
      atomicMax(keys + b, key);
      atomicAdd(ncand + b, (int)cnt);
    }
    __syncwarp();                                              // every lane is done with this buffer
    r = rn;
    cur ^= 1;
  }
}

__global__ void k_score_finalize(int B, const unsigned long long* __restrict__ keys, const int* __restrict__ ncand,
                                 float* max_scores, int32_t* argmax_anchor, int32_t* num_candidates) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const unsigned long long k = keys[b];
  // maximum(reduce_max(ragged scores), 0): an empty row reduces to float lowest -> 0 (attacker.py:190)
  max_scores[b] = k ? __uint_as_float((unsigned)(k >> 32)) : 0.0f;
  if (argmax_anchor) argmax_anchor[b] = k ? (int32_t)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)) : -1;
  if (num_candidates) num_candidates[b] = ncand[b];
}

__global__ void __launch_bounds__(kThreads) k_score_ties(int A, const unsigned long long* __restrict__ keys,
                                                         const float* __restrict__ cand_score, int* n_ties, int* tie_list) {
  const int b = blockIdx.y;
  const unsigned long long k = keys[b];
  if (!k) return;
  const float M = __uint_as_float((unsigned)(k >> 32));
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < A; a += gridDim.x * blockDim.x) {
    if (cand_score[(size_t)b * A + a] == M) tie_list[(size_t)b * A + atomicAdd(n_ties + b, 1)] = a;
  }
}

struct ZeroSegs {
  float* ptr[SCORE_MAX_LEVELS];
  long long n4[SCORE_MAX_LEVELS + 1];   // prefix of float4 counts
  long long tail_floats[SCORE_MAX_LEVELS];
  int n;
};

__global__ void __launch_bounds__(kThreads) k_score_zero(ZeroSegs z) {
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long total = z.n4[z.n];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < z.n && i >= z.n4[l + 1]) ++l;
    reinterpret_cast<float4*>(z.ptr[l])[i - z.n4[l]] = zero;
  }
  if (blockIdx.x == 0) {   // unaligned leftovers (level sizes that are not a multiple of 4 floats)
    for (int l = 0; l < z.n; ++l) {
      const long long done = (z.n4[l + 1] - z.n4[l]) * 4;
      for (long long i = threadIdx.x; i < z.tail_floats[l]; i += blockDim.x) z.ptr[l][done + i] = 0.0f;
    }
  }
}

__global__ void __launch_bounds__(128) k_score_scatter(ScoreShape s, ScoreLevels lv, const float* __restrict__ max_scores,
                                                       const float* __restrict__ scale, const unsigned long long* __restrict__ keys,
                                                       const int* __restrict__ n_ties, const int* __restrict__ tie_list) {
  const int b = blockIdx.x;
  if (!keys[b]) return;
  const int nt = n_ties[b];
  const int C = s.num_classes;
  const float M = max_scores[b];
  const float dM = 2.0f * M + 2.0f * (M - *scale);
  for (int i = threadIdx.x; i < nt; i += blockDim.x) {
    const int a = tie_list[(size_t)b * s.total_anchors + i];
    int l = 0;
    while (l + 1 < s.num_levels && a >= lv.anchor_base[l + 1]) ++l;
    const int a_local = a - lv.anchor_base[l];
    const float* x = lv.cls[l] + ((size_t)b * lv.n_anchors[l] + a_local) * C;
    float* d = lv.dcls[l] + ((size_t)b * lv.n_anchors[l] + a_local) * C;
    const float best = x[0];                      // class 0 is the (first) arg-max of a candidate
    int kc = 0;
    for (int c = 0; c < C; ++c) kc += (x[c] == best);
    const float dz = ((dM / (float)nt) * M) * (1.0f - M);    // SigmoidGrad: dy * y * (1 - y)
    const float per = dz / (float)kc;
    for (int c = 0; c < C; ++c)
      if (x[c] == best) d[c] = per;
  }
}

__global__ void __launch_bounds__(kThreads) k_score_scalars(int B, const float* __restrict__ max_scores,
                                                            const float* __restrict__ scale, float* dscale_out, float* loss_out) {
  __shared__ double red[32];
  const float sc = *scale;
  double ds = 0.0, ls = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float M = max_scores[b];
    const float diff = M - sc;
    ds += (double)(-(2.0f * diff));
    ls += (double)(M * M + diff * diff);
  }
  ds = block_sum(ds, red);
  ls = block_sum(ls, red);
  if (threadIdx.x == 0) {
    if (dscale_out) *dscale_out = (float)ds;
    if (loss_out) *loss_out = (float)ls;
  }
}

static int check_score_shape(const ScoreShape* s) {
  if (!s) { set_error("score shape is NULL"); return EOT_ERR_NULL_POINTER; }
  if (s->batch <= 0 || s->num_levels <= 0 || s->num_levels > SCORE_MAX_LEVELS || s->num_classes <= 0 ||
      (s->num_classes & 1) || s->anchors_per_loc <= 0 || s->total_anchors <= 0) {
    set_error("bad score shape: batch=%d levels=%d classes=%d (must be even) anchors/loc=%d A=%d", s->batch, s->num_levels,
              s->num_classes, s->anchors_per_loc, s->total_anchors);
    return EOT_ERR_BAD_SHAPE;
  }
  long long tot = 0;
  for (int l = 0; l < s->num_levels; ++l) {
    if (s->level_locs[l] <= 0) { set_error("bad score shape: level %d has %d locations", l, s->level_locs[l]); return EOT_ERR_BAD_SHAPE; }
    tot += (long long)s->level_locs[l] * s->anchors_per_loc;
  }
  if (tot != s->total_anchors) { set_error("bad score shape: total_anchors %d != %lld", s->total_anchors, tot); return EOT_ERR_BAD_SHAPE; }
  return EOT_OK;
}

static void fill_levels(const ScoreShape& s, const float* const* cls, const float* const* box, float* const* dcls, ScoreLevels* lv) {
  int base = 0, tbase = 0;
  for (int l = 0; l < s.num_levels; ++l) {
    lv->cls[l] = cls ? cls[l] : nullptr;
    lv->box[l] = box ? box[l] : nullptr;
    lv->dcls[l] = dcls ? dcls[l] : nullptr;
    lv->n_anchors[l] = s.level_locs[l] * s.anchors_per_loc;
    lv->anchor_base[l] = base;
    lv->tile_base[l] = tbase;
    base += lv->n_anchors[l];
    tbase += (lv->n_anchors[l] + kTile - 1) / kTile;
  }
  lv->tile_base[s.num_levels] = tbase;
  for (int l = s.num_levels; l < SCORE_MAX_LEVELS; ++l) lv->anchor_base[l] = base;
}

}  // namespace eot

using namespace eot;

extern "C" int score_workspace_bytes(const ScoreShape* shape, size_t* bytes) {
  if (int rc = check_score_shape(shape)) return rc;
  if (!bytes) { set_error("bytes is NULL"); return EOT_ERR_NULL_POINTER; }
  *bytes = score_layout(*shape).total;
  return EOT_OK;
}

extern "C" int score_candidate_offset(const ScoreShape* shape, size_t* offset) {
  if (int rc = check_score_shape(shape)) return rc;
  if (!offset) { set_error("offset is NULL"); return EOT_ERR_NULL_POINTER; }
  *offset = score_layout(*shape).off_cand;
  return EOT_OK;
}

extern "C" int score_max_fwd(const ScoreShape* shape, const float* const* cls_levels, const float* const* box_levels,
                             const float* anchors, float* max_scores, int32_t* argmax_anchor, int32_t* num_candidates,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_score_shape(shape)) return rc;
  if (!cls_levels || !box_levels || !anchors || !max_scores || !workspace) { set_error("score_max_fwd: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  const ScoreShape s = *shape;
  for (int l = 0; l < s.num_levels; ++l) {
    if (!cls_levels[l] || !box_levels[l]) { set_error("score_max_fwd: level %d pointer is NULL", l); return EOT_ERR_NULL_POINTER; }
    if (((uintptr_t)cls_levels[l] & 7) || ((uintptr_t)box_levels[l] & 15)) { set_error("score_max_fwd: level %d is misaligned", l); return EOT_ERR_MISALIGNED; }
  }
  if ((uintptr_t)anchors & 15) { set_error("score_max_fwd: anchors must be 16-byte aligned"); return EOT_ERR_MISALIGNED; }
  const ScoreLayout L = score_layout(s);
  if (workspace_bytes < L.total) { set_error("score workspace too small: %zu < %zu", workspace_bytes, L.total); return EOT_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = static_cast<char*>(workspace);
  ScoreLevels lv;
  fill_levels(s, cls_levels, box_levels, nullptr, &lv);
  EOT_CHECK_CUDA(cudaMemsetAsync(ws + L.off_keys, 0, L.off_cand - L.off_keys, st));
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + L.off_keys);
  int* ncand = reinterpret_cast<int*>(ws + L.off_counts);
  int* nties = ncand + s.batch;
  float* cand = reinterpret_cast<float*>(ws + L.off_cand);
  const long long total_tiles = (long long)lv.tile_base[s.num_levels] * s.batch;
  if (total_tiles >= (1ll << 31)) { set_error("score_max_fwd: too many anchor tiles"); return EOT_ERR_BAD_SHAPE; }
  const size_t smem = (size_t)kScoreWarps * 2 * kTile * s.num_classes * sizeof(float);   // two tile buffers per warp
  if (smem > 220 * 1024 || (s.num_classes & 1)) { set_error("score_max_fwd: num_classes %d not supported (even, <= 107)", s.num_classes); return EOT_ERR_BAD_SHAPE; }
  EOT_CHECK_CUDA(cudaFuncSetAttribute(k_score_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int need = (int)((total_tiles + kScoreWarps - 1) / kScoreWarps);
  const int grid = need < sm_count() ? need : sm_count();      // persistent: one CTA (8 warps x 2 tile buffers = 184 KB) per SM
  k_score_fwd<<<grid, kScoreWarps * 32, smem, st>>>(s, lv, anchors, keys, ncand, cand, (int)total_tiles);
  k_score_finalize<<<(s.batch + 127) / 128, 128, 0, st>>>(s.batch, keys, ncand, max_scores, argmax_anchor, num_candidates);
  const int tgrid = (s.total_anchors + kThreads * 8 - 1) / (kThreads * 8);
  k_score_ties<<<dim3(tgrid, s.batch), kThreads, 0, st>>>(s.total_anchors, keys, cand, nties, reinterpret_cast<int*>(ws + L.off_ties));
  count_launches(3);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}

extern "C" int score_max_bwd(const ScoreShape* shape, const float* const* cls_levels, const float* max_scores,
                             const float* scale, float* const* dcls_levels, float* dscale_out, float* loss_out,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_score_shape(shape)) return rc;
  if (!cls_levels || !max_scores || !scale || !dcls_levels || !workspace) { set_error("score_max_bwd: NULL pointer"); return EOT_ERR_NULL_POINTER; }
  const ScoreShape s = *shape;
  const ScoreLayout L = score_layout(s);
  if (workspace_bytes < L.total) { set_error("score workspace too small: %zu < %zu", workspace_bytes, L.total); return EOT_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = static_cast<char*>(workspace);
  ScoreLevels lv;
  fill_levels(s, cls_levels, nullptr, dcls_levels, &lv);
  ZeroSegs z;
  z.n = s.num_levels;
  long long acc = 0;
  for (int l = 0; l < s.num_levels; ++l) {
    if (!cls_levels[l] || !dcls_levels[l]) { set_error("score_max_bwd: level %d pointer is NULL", l); return EOT_ERR_NULL_POINTER; }
    if ((uintptr_t)dcls_levels[l] & 15) { set_error("score_max_bwd: dcls level %d is misaligned", l); return EOT_ERR_MISALIGNED; }
    const long long n = (long long)s.batch * lv.n_anchors[l] * s.num_classes;
    z.ptr[l] = dcls_levels[l];
    z.n4[l] = acc;
    acc += n / 4;
    z.tail_floats[l] = n % 4;
  }
  z.n4[s.num_levels] = acc;
  long long zb = (acc + kThreads * 4 - 1) / (kThreads * 4);
  const long long cap = (long long)sm_count() * 32;
  if (zb > cap) zb = cap;
  if (zb < 1) zb = 1;
  k_score_zero<<<(int)zb, kThreads, 0, st>>>(z);
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(ws + L.off_keys);
  const int* nties = reinterpret_cast<const int*>(ws + L.off_counts) + s.batch;
  k_score_scatter<<<s.batch, 128, 0, st>>>(s, lv, max_scores, scale, keys, nties, reinterpret_cast<const int*>(ws + L.off_ties));
  if (dscale_out || loss_out) k_score_scalars<<<1, kThreads, 0, st>>>(s.batch, max_scores, scale, dscale_out, loss_out);
  count_launches((dscale_out || loss_out) ? 3 : 2);
  EOT_CHECK_CUDA(cudaPeekAtLastError());
  return EOT_OK;
}
