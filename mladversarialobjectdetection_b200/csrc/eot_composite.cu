// Forward composite (attacker.py:436-444): rotate / projective-warp the padded transformed patch, `< -1` mask against
// the background, clip, paste -- for all boxes of all images, with the reference's sequential-paste semantics.
//
// Sequential paste, restated per element: the final value of channel c at a pixel is clip(R_k[c]) of the LAST box k (in
// paste order) whose warped patch covers it with R_k[c] >= -1 (attacker.py:440), else the (clipped) original.
//
// Work item = kCompRows consecutive window rows of one box, owned by ONE WARP (atomic ticket, item table written by the
// geometry role).  Per row the warp sweeps, in 32-pixel steps, the column range in which the rotated ps x ps core can
// be touched (row table; everything else blends to exactly -2 as the padded image would, i.e. background).
//
// Ownership: a pixel belongs to the NEWEST box whose column range of that row covers it; only that box's item writes it
// (its own sample, and -- for the channels that leaves open, R < -1: the edge of its core -- the samples of the older
// boxes covering the pixel, newest first).  An older box drops the pixels a newer range covers without sampling.  Every
// (pixel, channel) has exactly one writer: no ordering between work items, no cap on the boxes per image, and one
// sample per pixel except along core edges inside overlaps.  Images holding values outside [-1,1] (the window-wide clip
// of attacker.py:441 is then not the identity on the background) use whole window rows as ranges and the owner rewrites
// all three channels, the open ones with clip(original).
// Channels nobody pastes keep the value the image pass copied.  Route byte (bits 0-2) in the map of the box a channel
// came from: channel c of the output pixel is that box's sample and passes the outer clip -- the backward's
// TensorScatterUpdate / SelectV2 / clip routing.
#include "eot_composite.cuh"

namespace eot {

template <bool kMask>
__global__ void __launch_bounds__(kThreads, EOT_COMP_MINB) k_composite3(EotShape s, Layout L, char* ws,
                                                                       const float* __restrict__ images, float* out, float* mask,
                                                                       const int32_t* __restrict__ offsets, int b0, int b1,
                                                                       int group, int ngroups, float one) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const BoxPlan* plans = reinterpret_cast<const BoxPlan*>(ws + L.off_plans);
  const float* ubuf = reinterpret_cast<const float*>(ws + L.off_u);
  uint8_t* routes = reinterpret_cast<uint8_t*>(ws + L.off_route);
  const int2* rowtab = reinterpret_cast<const int2*>(ws + L.off_rowtab);
  const int2* citems = reinterpret_cast<const int2*>(ws + L.off_citems);
  const int4* base = reinterpret_cast<const int4*>(ws + L.off_base);
  const int open_cap = (int)(L.open_cap / ngroups);               // the group's share of the open-pixel list
  int* open_count = reinterpret_cast<int*>(ws + L.off_counters) + 8 + group;
  int2* open_list = reinterpret_cast<int2*>(ws + L.off_open) + (size_t)group * open_cap;
  const int lo = base[min(offsets[b0], s.total_boxes)].w, hi = base[min(offsets[b1], s.total_boxes)].w;
  WarpTickets tk;
  tk.init(ws, L.off_tickets, 2 * group + 1);
  int it = lo + tk.item(tk.draw(lane));
  int2 item = it < hi ? __ldcg(citems + it) : make_int2(0, 0);
  while (it < hi) {
    const int nxt = tk.draw(lane);                                // next item's ticket travels while this one computes
    composite_item_main<kMask>(s, L, ws, plans, ubuf, rowtab, routes, images, out, mask, item.x, item.y, open_count, open_list, open_cap,
                               one, lane);
    it = lo + tk.item(nxt);
    if (it < hi) item = __ldcg(citems + it);
  }
}

#ifndef EOT_REST_MINB
#define EOT_REST_MINB 2
#endif
template <bool kMask>
__global__ void __launch_bounds__(kThreads, EOT_REST_MINB) k_composite_rest(EotShape s, Layout L, char* ws, const float* __restrict__ images,
                                                             float* out, float* mask, const int32_t* __restrict__ offsets, int b0,
                                                             int b1, int group, int ngroups) {
  pdl_wait();
  composite_rest_body<kMask>(s, L, ws, images, out, mask, offsets, b0, b1, group, ngroups, blockIdx.x * blockDim.x + threadIdx.x,
                             gridDim.x * blockDim.x);
}

int launch_composite3(const EotShape& s, const Layout& L, char* ws, const int32_t* offsets, const float* images, float* out,
                      float* mask, int b0, int b1, int group, int ngroups, cudaStream_t st) {
  const int nsm = sm_count();
  if (mask) {
    EOT_CHECK_CUDA(launch_pdl(k_composite3<true>, dim3(nsm * EOT_COMP_MINB), dim3(kThreads), 0, st, s, L, ws, images, out, mask, offsets, b0, b1, group, ngroups, 1.0f));
    EOT_CHECK_CUDA(launch_pdl(k_composite_rest<true>, dim3(nsm * EOT_REST_MINB), dim3(kThreads), 0, st, s, L, ws, images, out, mask, offsets, b0, b1, group, ngroups));
  } else {
    EOT_CHECK_CUDA(launch_pdl(k_composite3<false>, dim3(nsm * EOT_COMP_MINB), dim3(kThreads), 0, st, s, L, ws, images, out, mask, offsets, b0, b1, group, ngroups, 1.0f));
    EOT_CHECK_CUDA(launch_pdl(k_composite_rest<false>, dim3(nsm * EOT_REST_MINB), dim3(kThreads), 0, st, s, L, ws, images, out, mask, offsets, b0, b1, group, ngroups));
  }
  count_launches(2);
  return EOT_OK;
}

}  // namespace eot
