// Forward resize of the matched patch to every box's patch side, + noise + brightness delta + clip
// (attacker.py:425-428; tf.image.resize(antialias=True) = ScaleAndTranslate: GatherRows then GatherColumns, each a
// sequential float32 accumulation over the span).
//
// Work item = `rb` consecutive output rows of one box, owned by ONE WARP: the intermediate row of ScaleAndTranslate
// (output row oy over all P source columns) is only ever read by output row oy, so a warp produces it into its own
// shared-memory scratch and consumes it itself -- no CTA barrier, no strip staging, no cross-warp traffic.
//   rows pass     lane = source column (128-bit loads of the RGBX matched patch, coalesced); the row's span start and
//                 tap weights are warp-uniform registers
//   columns pass  lane = output column; the column's taps are fetched once and used for every row of the item; taps from
//                 the warp's scratch
//   noise         Philox4x32-10 words of 128 columns of a row (384 elements = 96 counter values: three rounds of the
//                 full warp) are produced into the scratch and picked up by the texel lanes (stride 3 words: conflict-free)
// Items are handed out by an atomic ticket per warp (fetched one item ahead, together with its (box, block) record
// from the item table the geometry role wrote).
//
// Code paths, chosen per box: the two-tap table (up-sampling / unit scale: the bulk of the texels at the reference's
// scale .4), compile-time spans 3 / 5 / 7 / 9 (down-sampling) and a run-time span.
#include "eot_resize.cuh"

namespace eot {

#ifndef EOT_RESIZE_MINB
#define EOT_RESIZE_MINB 4
#endif
__global__ void __launch_bounds__(kThreads, EOT_RESIZE_MINB) k_resize2(EotShape s, Layout L, char* ws,
                                                                      const int32_t* __restrict__ offsets, int b0, int b1,
                                                                      int ticket_slot, float one) {
  extern __shared__ __align__(16) unsigned char resize_smem[];
  pdl_wait();
  pdl_trigger();
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
    if (mode == 2) resize_item<2>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
    else if (mode == 5) resize_item<5>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
    else if (mode == 7) resize_item<7>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
    else if (mode == 3) resize_item<3>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
    else if (mode == 9) resize_item<9>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
    else resize_item<0>(s, L, ws, pl, item.x, item.y, inter, words, one, lane);
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
  EOT_CHECK_CUDA(launch_pdl(k_resize2, dim3(sm_count() * per_sm), dim3(kThreads), smem, st, s, L, ws, offsets, b0, b1, ticket_slot, 1.0f));
  count_launches(1);
  return EOT_OK;
}

}  // namespace eot
