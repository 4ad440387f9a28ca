// Shared device/host definitions for libeotpatch (sm_100a).
//
// Parity note: this translation-unit family is compiled with -fmad=false.  The reference runs every
// arithmetic op as its own float32 TF kernel (one rounding per op), and the `< -1` mask decision of
// attacker.py:440 depends on the last bit, so no multiply-add may be contracted.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eotpatch.h"

namespace eot {

// ---- constants of the restated TF ops (tensorflow/python/ops/image_ops_impl.py) -------------------
#define EOT_K00 0.299f
#define EOT_K10 0.587f
#define EOT_K20 0.114f
#define EOT_K01 (-0.14714119f)
#define EOT_K11 (-0.28886916f)
#define EOT_K21 0.43601035f
#define EOT_K02 0.61497538f
#define EOT_K12 (-0.51496512f)
#define EOT_K22 (-0.10001026f)
// yuv_to_rgb kernel rows: Y -> (1,1,1); U -> (0, -0.394642334, 2.03206185); V -> (1.13988303, -0.58062185, 0)
#define EOT_I10 0.0f
#define EOT_I11 (-0.394642334f)
#define EOT_I12 2.03206185f
#define EOT_I20 1.13988303f
#define EOT_I21 (-0.58062185f)
#define EOT_I22 0.0f
#define EOT_C127_255 ((float)(127.0 / 255.0))   // brightness_matcher.py:32
#define EOT_C255_127 ((float)(255.0 / 127.0))   // brightness_matcher.py:41
#define EOT_SQRT2 1.41421354f                   // float32(2. ** .5), attacker.py:470

#ifndef EOT_COMP_ROWS
#define EOT_COMP_ROWS 2
#endif
#ifndef EOT_RESIZE_RB
#define EOT_RESIZE_RB 2
#endif
#ifndef EOT_RESIZE_ROWS_CAP
#define EOT_RESIZE_ROWS_CAP 12
#endif
#ifndef EOT_BWD_ROWS_CAP
#define EOT_BWD_ROWS_CAP 32
#endif
constexpr int kCompRows = EOT_COMP_ROWS;  // window rows per composite work item (one warp)
// forward resize: texels per Philox round trip (96 counters = 3 full warps) and the words of one such chunk of a row
constexpr int kNoiseChunk = 128;
constexpr int kNoiseWords = (kNoiseChunk * 3 / 4 + 1) * 4;          // (+ one counter when the chunk starts unaligned)
constexpr int kThreads = 256;
// Work tickets.  One atomic counter handing out every item of a kernel is the bottleneck of warp-granular items (a
// same-address atomic completes every ~2.3 ns on B200: 27 k items = 60 us).  Each kernel therefore owns kTicketLanes
// counters, 256 bytes apart; counter c hands out the items c, c + kTicketLanes, c + 2 kTicketLanes, ...; a warp draws
// from the counter (global warp index mod kTicketLanes) only.
constexpr int kTicketSlots = 16;
constexpr int kTicketLanes = 16;

// Per-box plan written by the geometry kernel; 144 bytes.
struct __align__(16) BoxPlan {
  int32_t y0, x0, ps, d;
  int32_t pad_lo, pad_hi, valid, span;
  float T[8];     // output->input transform of tfa.image.rotate (+ projective row)
  float Ti[8];    // its inverse, normalised (gradient warp)
  float delta;
  uint32_t key0, key1;
  int32_t image;
  int64_t u_off;  // float offset of this box's transformed-patch buffer
  int32_t first_box, last_box;   // CSR range of the owning image
  float ia0, ia3;                // 1/T[0], 1/T[3] (0 when the coefficient is ~0): column range of the core per window row
  int32_t two_tap;               // 1: the resize of this box runs from the two-tap table
  int32_t rsv;
};
static_assert(sizeof(BoxPlan) == 144, "BoxPlan must stay 144 bytes");

// Workspace layout (byte offsets); identical on host and device.
struct Layout {
  size_t off_ysum_img;     // double[B]
  size_t off_ysum_patch;   // double[B]
  size_t off_gy_sum;       // double[B]   (backward: sum of dL/dY per image)
  size_t off_bwd_cnt;      // int32[1 + ceil(P*P/256)] backward: work ticket, finished image groups per texel chunk
  size_t off_oor;          // int32[B]    image b holds a value outside [-1,1] (then clip(background) is not the identity)
  size_t off_counters;     // int32[32]: 2 error flag, 5 finished geometry blocks, 6 an image needs the composite's general
                           //            path (out-of-range values or more than 32 boxes), 8 + g open pixels listed by the composite
                           //            of image group g
  size_t off_tickets;      // int32[kTicketSlots][kTicketLanes][64]  work tickets: slot = kernel (x image group), lane = one of
                           //            the interleaved sub-queues (own 256-byte line each: same-address atomics serialise)
  size_t off_fused;        // int32[...] fused forward kernel: readiness flag, task total, per-image progress counters, step table
                           //            (fused_ints(); zeroed with the accumulators by the call's memset)
  size_t off_cost;         // uint32[B] window work of image b (sum of ps^2 over its valid boxes), added up by the geometry blocks
  size_t off_plans;        // BoxPlan[N]
  size_t off_starts;       // int32[N][Lmin]
  size_t off_weights;      // float[N][wcap]  tap-major: weight of tap k of output index o at [k * ps + o]
  size_t off_tab2;         // float4[N][Lmin] two-tap boxes (up-sampling / unit scale: at most two adjacent non-zero taps
                           //            per output index): (source index a, source index b, weight a, weight b)
  size_t off_match;        // float4[B][P*P] matched patch per image, RGBX texels
  size_t off_u;            // float4[N][slot/4]: (ps+4)^2 texels per box = clipped (r,g,b) of the transformed patch +
                           //            inner-clip pass bits, inside a two-texel ring of the -2 pad / fill value
  size_t off_cnt;          // int4[N]    work items of box j: (backward window strips, backward resize strips, forward resize
                           //            row blocks, forward composite row blocks); 0 when invalid
  size_t off_base;         // int4[N+1]  exclusive prefix sums of off_cnt (box order == image order)
  size_t off_items;        // int2[N * ceil(Lmin / rb)] forward resize work items in ticket order: (box, row block)
  size_t off_citems;       // int2[N * ceil(min(H,W) / kCompRows)] forward composite work items: (box, window row block)
  size_t off_rowtab;       // int2[N][min(H,W)] per window row: first / last column whose sample can touch the core
  size_t off_open;         // int2[open_cap] composite: owned pixels with an open channel inside an older box's range:
                           //            (box, row << 16 | column << 3 | open channels)
  size_t off_inv;          // int2[N][P]  for patch index i: first/last output index whose span holds i
  size_t off_wt;           // float[N][P*tcap] transposed resize weights, row stride = the box's own max tap count:
                           //              tap k of patch index i = weight of output index st+k
  size_t off_stt;          // int2[N][P+1] (first output index, tap count) of patch index i (clamped into [0, ps));
                           //              entry P = (max tap count, patch rows per backward strip)
  size_t off_route;        // uint8[N][rslot] per window pixel: bit c = channel c of the output came from this box (and passes the clip)
  size_t off_gm;           // float[B][P*P*3] backward: dL/d(matched patch) per image
  size_t off_gu;           // float[N][gslot]  backward: dL/d(u) per box
  size_t off_gp_part;      // float[16][P*P*3] backward: partial dL/dpatch per image group
  size_t off_gbox;         // float[N][P*P*3]  backward: dL/d(matched patch) per box (0 bytes when it would exceed gbox_cap)
  size_t off_offsets;      // int32[B+1] copy of the CSR row splits (the backward has no other source)
  size_t off_order;        // int32[B]   images by decreasing window work (the backward starts the heavy images first)
  size_t total;
  int64_t slot;            // floats per u slot (4 per texel)
  int64_t gslot;           // floats per g_u slot (RGBX: 4 per texel)
  int32_t tcap;            // taps per patch index in the transposed weight table
  int32_t wcap;            // floats per weight table
  int32_t lmin;            // max patch side
  int32_t resize_rows;     // output rows per backward window strip
  int32_t rb;              // output rows per forward resize item (one warp; shared memory: rb * P * 16 B per warp)
  int32_t cr;              // window rows per forward composite item (one warp)
  int64_t open_cap;        // entries of the open-pixel list (overflow: the composite redoes every row the slow way)
  uint32_t p3_magic;       // ceil(2^32 / (3P)): idx / (3P) == umulhi(idx, p3_magic) for idx < 2^16
  int64_t rslot;           // bytes per route map
  int32_t use_gbox;        // per-box partial gradients fit: fully parallel resize adjoint
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Control block of the fused forward kernel (eot_fwd.cu: k_forward_fused), int32 units:
//   [0] tables ready   [1] tasks in the queue   [2] finished patch-statistics blocks   [3] images whose composite is done
//   [64 + 8 b + i]     image b: 0 tiles read (luma summed), 1 tiles stored, 2 match parts, 3 resize items, 4 composite items done
//   step_start[steps + 1], step_info int4[steps]   with steps = B + 2 * skew (skew <= kMaxSkew)
constexpr int kMaxSkew = 48;
__host__ __device__ inline size_t fused_steps_cap(int B) { return (size_t)B + 2 * kMaxSkew; }
__host__ __device__ inline size_t fused_start_ints(int B) { return (fused_steps_cap(B) + 1 + 3) / 4 * 4; }
__host__ __device__ inline size_t fused_ints(int B) { return 64 + 8 * (size_t)B + fused_start_ints(B) + 4 * fused_steps_cap(B); }

__host__ __device__ inline Layout make_layout(const EotShape& s) {
  Layout L;
  const size_t B = (size_t)s.batch, N = (size_t)(s.total_boxes > 0 ? s.total_boxes : 0);
  const int lfull = s.height < s.width ? s.height : s.width;
  float ms = (s.max_scale > 0.f && s.max_scale < 1.f) ? s.max_scale : 1.f;
  int lmin = (int)((float)(s.height > s.width ? s.height : s.width) * ms) + 1;
  if (lmin > lfull) lmin = lfull;
  L.lmin = lmin;
  L.wcap = 2 * s.patch_size + 3 * lmin + 8;
  L.slot = (int64_t)align_up((size_t)(lmin + 4) * (lmin + 4) * 4, 32);
  L.gslot = (int64_t)align_up((size_t)lmin * lmin * 4, 32);
  L.tcap = 3 * ((lmin + s.patch_size - 1) / s.patch_size) + 3;
  int rr = 2560 / s.patch_size;              // <= 40 KB of RGBX float32 intermediate rows
  L.resize_rows = rr > EOT_RESIZE_ROWS_CAP ? EOT_RESIZE_ROWS_CAP : (rr < 1 ? 1 : rr);
  // rows per forward resize item: two when four CTAs of eight such warps still fit an SM's shared memory (P <= 110),
  // else one -- measured at P = 300: 4 CTAs per SM with one row per item beat 2 CTAs with two rows (137 -> 131 us)
  L.rb = EOT_RESIZE_RB;
  while (L.rb > 1 && 8 * (size_t)L.rb * ((size_t)s.patch_size * 16 + kNoiseWords * 4) > 53 * 1024) --L.rb;
  L.cr = EOT_COMP_ROWS;
  L.p3_magic = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)(s.patch_size * 3) - 1) / (uint64_t)(s.patch_size * 3));
  const size_t PP3 = (size_t)s.patch_size * s.patch_size * 3;
  size_t o = 0;
  L.off_ysum_img = o;     o = align_up(o + B * sizeof(double), 256);
  L.off_ysum_patch = o;   o = align_up(o + B * sizeof(double), 256);
  L.off_gy_sum = o;       o = align_up(o + B * sizeof(double), 256);
  L.off_bwd_cnt = o;      o = align_up(o + (1 + ((size_t)s.patch_size * s.patch_size + 255) / 256) * sizeof(int32_t), 256);
  L.off_oor = o;          o = align_up(o + B * sizeof(int32_t), 256);
  L.off_counters = o;     o = align_up(o + 32 * sizeof(int32_t), 256);
  L.off_tickets = o;      o = align_up(o + (size_t)kTicketSlots * kTicketLanes * 256, 256);
  L.off_fused = o;        o = align_up(o + fused_ints((int)B) * sizeof(int32_t), 256);
  L.off_cost = o;         o = align_up(o + B * sizeof(uint32_t), 256);
  L.off_plans = o;        o = align_up(o + N * sizeof(BoxPlan), 256);
  L.off_starts = o;       o = align_up(o + N * (size_t)lmin * sizeof(int32_t), 256);
  L.off_weights = o;      o = align_up(o + N * (size_t)L.wcap * sizeof(float), 256);
  L.off_tab2 = o;         o = align_up(o + N * (size_t)lmin * 16, 256);
  L.off_match = o;        o = align_up(o + B * (size_t)s.patch_size * s.patch_size * 16, 256);
  L.off_u = o;            o = align_up(o + N * (size_t)L.slot * sizeof(float), 256);
  L.off_cnt = o;          o = align_up(o + N * 16, 256);
  L.off_base = o;         o = align_up(o + (N + 1) * 16, 256);
  L.off_items = o;        o = align_up(o + N * (size_t)((lmin + L.rb - 1) / L.rb) * 8, 256);
  L.off_citems = o;       o = align_up(o + N * (size_t)((lfull + L.cr - 1) / L.cr) * 8, 256);
  L.off_rowtab = o;       o = align_up(o + N * (size_t)lfull * 8, 256);
  L.open_cap = (int64_t)N * lfull * 4 < ((int64_t)1 << 30) ? (int64_t)N * lfull * 4 : ((int64_t)1 << 30);
  L.off_open = o;         o = align_up(o + (size_t)L.open_cap * 8, 256);
  L.off_inv = o;          o = align_up(o + N * (size_t)s.patch_size * 8, 256);
  L.off_wt = o;           o = align_up(o + N * (size_t)s.patch_size * L.tcap * sizeof(float), 256);
  L.off_stt = o;          o = align_up(o + N * (size_t)(s.patch_size + 1) * 8, 256);
  L.rslot = (int64_t)align_up((size_t)lfull * lfull, 32);
  L.off_route = o;        o = align_up(o + N * (size_t)L.rslot, 256);
  L.off_gm = o;           o = align_up(o + B * PP3 * sizeof(float), 256);
  L.off_gu = o;           // (the backward keeps dL/d(u) in shared memory)
  L.off_gp_part = o;      o = align_up(o + 16 * PP3 * sizeof(float), 256);
  L.use_gbox = (N * PP3 * sizeof(float) <= ((size_t)1 << 30) && !(s.flags & EOT_FLAG_SERIAL_ADJOINT)) ? 1 : 0;
  L.off_gbox = o;         // (the backward accumulates per image in shared memory: no per-box partials)
  L.off_offsets = o;      o = align_up(o + (B + 1) * sizeof(int32_t), 256);
  L.off_order = o;        o = align_up(o + B * sizeof(int32_t), 256);
  L.total = o;
  return L;
}

// Forward resize strips of a box (also the items of the backward's window kernel): at most `cap` output rows each,
// balanced (rows = ceil(ps / strips)).  Sizing the strips to multiples of the CTA's thread count was measured and did
// not pay (profiles/r01_launches.md).
__host__ __device__ inline int fwd_strips(int ps, int cap) {
  if (ps <= 0) return 0;
  const int r = cap < ps ? cap : ps;
  return (ps + r - 1) / r;
}
__host__ __device__ inline int strip_rows(int ps, int strips) { return strips > 0 ? (ps + strips - 1) / strips : 1; }

// patch rows per strip of the backward resize adjoint: <= 40 KB of RGBX intermediate rows of ps texels
__host__ __device__ inline int bwd_strip_rows(int ps) {
  int rr = 2560 / (ps > 0 ? ps : 1);
  return rr > EOT_BWD_ROWS_CAP ? EOT_BWD_ROWS_CAP : (rr < 1 ? 1 : rr);
}

// ---- error plumbing (host) -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
void count_launches(int n);   // kernels launched by this library since load (bench.py reports the delta)

// forward window kernels (eot_resize.cu, eot_composite.cu): images [b0,b1), atomic work-ticket slot in the workspace
int launch_resize2(const EotShape& s, const struct Layout& L, char* ws, const int32_t* offsets, int b0, int b1,
                   int ticket_slot, cudaStream_t st);
#ifndef EOT_PACKED_MATH
#define EOT_PACKED_MATH 1
#endif
constexpr int kMaxGroups = 8;   // image groups of one forward call (own work tickets and open-pixel list each)
int launch_composite3(const EotShape& s, const struct Layout& L, char* ws, const int32_t* offsets, const float* images,
                      float* out, float* mask, int b0, int b1, int group, int ngroups, cudaStream_t st);

// Debug aid, off unless the environment holds EOT_KERNEL_TIMES=1: CUDA events between the stages of one call; the
// destructor synchronises the stream and prints the stage times to stderr (warm caches, real launch gaps -- what the
// ncu launch list cannot show).  Never active in a captured stream.
struct StageTimer {
  StageTimer(cudaStream_t st, const char* what);
  ~StageTimer();
  void mark(const char* stage);
  bool on;
  cudaStream_t st;
  const char* what;
  int n;
  cudaEvent_t ev[16];
  const char* names[16];
};

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains (launch latency
// and block ramp-up overlap the predecessor's tail); it calls pdl_wait() before touching anything an earlier kernel
// wrote.  Every kernel of a chain waits, so completion stays transitive.  EOT_PDL=0 launches plainly.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // measured on B200 (profiles/r02_fused_forward.md): eager launches gain (config 4: 135 -> 125 us, config 2 unchanged), replays of
  // a captured graph lose (config 2: 191 -> 199 us) -- graph edges are cheap already -- so captures launch plainly
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
  cfg.numAttrs = (pdl_enabled() && !capturing) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define EOT_CHECK_CUDA(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::eot::cuda_fail(_e, #expr); \
  } while (0)

#ifdef __CUDACC__
// ---- small device helpers ---------------------------------------------------------------------------
// (no-op for a plain launch)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel of the stream (if launched with launch_pdl) become resident as this grid's CTAs retire; it still
// waits in pdl_wait() for this grid to complete.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Makes a pointer opaque to the optimiser: under register pressure ptxas otherwise re-derives per-row base pointers
// from the kernel parameters (a dozen 64-bit instructions) at every use inside the pixel loops.
template <typename T>
__device__ __forceinline__ T* keep_ptr(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
  if (wid == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// Philox4x32-10, counter (c0,0,0,0), key (k0,k1).
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t k0, uint32_t k1) {
  uint32_t c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;            // one IMAD.WIDE each
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ (k0 + (uint32_t)r * 0x9E3779B9u);
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ (k1 + (uint32_t)r * 0xBB67AE85u);
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Ticket dispenser of one warp: next() returns the warp's next item index in [0, n) or a value >= n when its sub-queue
// is exhausted (lane 0 draws, all lanes get the value).
struct WarpTickets {
  int* counter;
  int sub;
  __device__ __forceinline__ void init(char* ws, size_t off_tickets, int slot) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    sub = gw % kTicketLanes;
    counter = reinterpret_cast<int*>(ws + off_tickets) + ((size_t)slot * kTicketLanes + sub) * 64;
  }
  // raw draw by lane 0 (other lanes: 0); combine with item() after a shuffle so that the atomic's latency can be hidden
  __device__ __forceinline__ int draw(int lane) const { return lane == 0 ? atomicAdd(counter, 1) : 0; }
  __device__ __forceinline__ int item(int drawn) const { return __shfl_sync(0xffffffffu, drawn, 0) * kTicketLanes + sub; }
};

// Work-item kinds of the per-box count / prefix tables (components of the int4 entries).
enum ItemKind { kItemBwdWindow = 0, kItemBwdResize = 1, kItemResize = 2, kItemComposite = 3 };

// Work item i of kind `which` -> (box, index inside the box).  base[] holds the exclusive prefix sums of the per-box
// item counts, in box (== image) order; `stride` = ints between consecutive entries (4 in the workspace table, 1 in a
// staged single-kind copy).
__device__ __forceinline__ int2 find_item(const int* __restrict__ b, int stride, int N, int i) {
  int lo = 0, hi = N;                      // last j with base[j] <= i
  while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (b[stride * m] <= i) lo = m; else hi = m; }
  return make_int2(lo, i - b[stride * lo]);
}

// Copies one kind of the prefix table to shared memory when it fits (the binary search then costs no global latency).
// Returns the table to search and sets *stride.
constexpr int kMaxBaseSmem = 2049;
__device__ __forceinline__ const int* stage_base(const int4* base, int which, int N, int* smem, int* stride) {
  if (N + 1 > kMaxBaseSmem) { *stride = 4; return reinterpret_cast<const int*>(base) + which; }
  for (int i = threadIdx.x; i <= N; i += blockDim.x) smem[i] = reinterpret_cast<const int*>(base)[4 * i + which];
  __syncthreads();
  *stride = 1;
  return smem;
}

// TF Uint32ToFloat + random_uniform range map: u*(hi-lo)+lo with lo=-amp, hi=amp.
__device__ __forceinline__ float noise_from_word(uint32_t w, float amp) {
  const float u = __uint_as_float((w & 0x7FFFFFu) | 0x3F800000u) - 1.0f;
  const float lo = -amp;
  const float rng = amp - lo;
  return u * rng + lo;
}

// print adjust + rescale + Y of one patch texel (attacker.py:372, brightness_matcher.py:32,58).
struct TexelYuv { float y, u, v, q0, q1, q2; };
__device__ __forceinline__ TexelYuv texel_yuv(float p0, float p1, float p2, const float* __restrict__ wb) {
  TexelYuv t;
  t.q0 = wb[0] * p0 + wb[3];
  t.q1 = wb[1] * p1 + wb[4];
  t.q2 = wb[2] * p2 + wb[5];
  const float s0 = (clampf(t.q0, -1.f, 1.f) + 1.0f) * EOT_C127_255;
  const float s1 = (clampf(t.q1, -1.f, 1.f) + 1.0f) * EOT_C127_255;
  const float s2 = (clampf(t.q2, -1.f, 1.f) + 1.0f) * EOT_C127_255;
  t.y = (s0 * EOT_K00 + s1 * EOT_K10) + s2 * EOT_K20;
  t.u = (s0 * EOT_K01 + s1 * EOT_K11) + s2 * EOT_K21;
  t.v = (s0 * EOT_K02 + s1 * EOT_K12) + s2 * EOT_K22;
  return t;
}

// ImageProjectiveTransformV3 BILINEAR / CONSTANT sampling of the (virtually) padded transformed
// patch of one box at a window pixel, three channels at once.  u holds clip((resize + noise) + delta)
// (attacker.py:425-428) as RGBX texels in a (ps+4)^2 buffer whose two-texel ring holds -2: the pad ring of
// attacker.py:435 and the fill of :437 both read as -2, so clamping the floor coordinate to
// [pad_lo-2, pad_lo+ps] makes all four taps unconditional loads at base, base+1, base+S, base+S+1.
// For the reference's pure rotation the projective row is zero, proj == 1 exactly and x / 1 == x, so
// the two divisions are skipped without changing a bit.
struct Sampler {
  float t0, t1, t2, t3, t4, t5, t6, t7;
  float lo2, hi;         // clamp range of the floor coordinates (padded window coordinates)
  int org;               // pad_lo - 2: padded coordinate of ring texel 0
  int S;                 // ps + 4: row stride of the ringed texel buffer
  const float4* u;
  bool affine;
};

__device__ __forceinline__ int u_stride(int ps) { return ps + 4; }
__device__ __forceinline__ int u_index(int ps, int ty, int tx) { return (ty + 2) * (ps + 4) + (tx + 2); }

// Conservative range of window columns x in window row wy whose sample can touch the ps x ps core: outside it all four
// taps are pad / fill and the box contributes nothing.  Safety margins of a pixel and more; the per-pixel core test
// decides exactly.  Affine T: l < T0 x + c0 < h and l < T3 x + c1 < h.  Projective T (denominator D(x) = T6 x + q positive
// over the row, else the whole row): (T0 x + c0) / D(x) > L  <=>  (T0 - L T6) x > L q - c0, a half-line per bound.
__device__ __forceinline__ void half_line(float a, float b, bool greater, float* lo, float* hi) {
  // a * x > b (greater) or a * x < b, intersected into [lo, hi] with a one-pixel margin
  if (fabsf(a) < 1e-6f) {
    const bool holds = greater ? (0.0f > b - 1.0f) : (0.0f < b + 1.0f);   // a ~ 0: a * x ~ 0 for every column of the row
    if (!holds) { *lo = 1.0f; *hi = 0.0f; }
    return;
  }
  const float x0 = b / a;
  if ((a > 0.0f) == greater) *lo = fmaxf(*lo, x0 - 1.0f);
  else *hi = fminf(*hi, x0 + 1.0f);
}
__device__ __forceinline__ void row_core_range(const BoxPlan& pl, int wy, int* xa, int* xb) {
  const float yf = (float)wy;
  float lo = 0.0f, hi = (float)(pl.d - 1);
  const float clo = (float)pl.pad_lo, chi = (float)(pl.pad_lo + pl.ps);   // core bounds in padded window coordinates
  const float c0 = pl.T[1] * yf + pl.T[2], c1 = pl.T[4] * yf + pl.T[5];
  if (pl.T[6] != 0.0f || pl.T[7] != 0.0f) {
    const float q = pl.T[7] * yf + 1.0f;
    const float d0 = q, d1 = pl.T[6] * hi + q;
    if (!(d0 > 0.05f && d1 > 0.05f)) { *xa = 0; *xb = pl.d - 1; return; }   // denominator near / through zero: whole row
    const float L = clo - 2.0f, U = chi + 1.0f;                    // half a pixel more slack than the affine bounds
    half_line(pl.T[0] - L * pl.T[6], L * q - c0, true, &lo, &hi);
    half_line(pl.T[0] - U * pl.T[6], U * q - c0, false, &lo, &hi);
    half_line(pl.T[3] - L * pl.T[6], L * q - c1, true, &lo, &hi);
    half_line(pl.T[3] - U * pl.T[6], U * q - c1, false, &lo, &hi);
    *xa = max((int)floorf(lo) - 1, 0);
    *xb = min((int)ceilf(hi) + 1, pl.d - 1);
    return;
  }
  {
    const float l = clo - 1.5f - c0, h = chi + 0.5f - c0;      // need l < T0*x < h
    if (pl.ia0 == 0.0f) { if (!(0.0f > l - 1.0f && 0.0f < h + 1.0f)) { lo = 1.0f; hi = 0.0f; } }
    else { const float x1 = l * pl.ia0, x2 = h * pl.ia0; lo = fmaxf(lo, fminf(x1, x2) - 1.0f); hi = fminf(hi, fmaxf(x1, x2) + 1.0f); }
  }
  {
    const float l = clo - 1.5f - c1, h = chi + 0.5f - c1;
    if (pl.ia3 == 0.0f) { if (!(0.0f > l - 1.0f && 0.0f < h + 1.0f)) { lo = 1.0f; hi = 0.0f; } }
    else { const float x1 = l * pl.ia3, x2 = h * pl.ia3; lo = fmaxf(lo, fminf(x1, x2) - 1.0f); hi = fminf(hi, fmaxf(x1, x2) + 1.0f); }
  }
  *xa = max((int)floorf(lo), 0);
  *xb = min((int)ceilf(hi), pl.d - 1);
}

__device__ __forceinline__ Sampler make_sampler(const BoxPlan& pl, const float* ubuf) {
  Sampler S;
  S.t0 = pl.T[0]; S.t1 = pl.T[1]; S.t2 = pl.T[2]; S.t3 = pl.T[3];
  S.t4 = pl.T[4]; S.t5 = pl.T[5]; S.t6 = pl.T[6]; S.t7 = pl.T[7];
  S.lo2 = (float)(pl.pad_lo - 2);
  S.hi = (float)(pl.pad_lo + pl.ps);
  S.org = pl.pad_lo - 2;
  S.S = pl.ps + 4;
  S.u = reinterpret_cast<const float4*>(ubuf + pl.u_off);
  S.affine = (pl.T[6] == 0.0f && pl.T[7] == 0.0f);
  return S;
}

__device__ __forceinline__ void blend3(const float4 v00, const float4 v01, const float4 v10, const float4 v11, float wx1,
                                       float wx0, float wy1, float wy0, float R[3]) {
  R[0] = wy1 * (wx1 * v00.x + wx0 * v01.x) + wy0 * (wx1 * v10.x + wx0 * v11.x);
  R[1] = wy1 * (wx1 * v00.y + wx0 * v01.y) + wy0 * (wx1 * v10.y + wx0 * v11.y);
  R[2] = wy1 * (wx1 * v00.z + wx0 * v01.z) + wy0 * (wx1 * v10.z + wx0 * v11.z);
}

// ---- packed float32 pairs (sm_100: FMUL2 / FFMA2 work on 64-bit register pairs, two IEEE results per instruction) -----
// The reference rounds after every multiply and every add.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even
// under --fmad false (seen in the SASS), so the add is issued as fma(a, one, b) with a RUN-TIME 1.0 (a kernel argument):
// exact for the sum, and a multiplier the assembler does not know cannot be folded into the preceding multiply.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b, f32x2 one) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(one), "l"(b));
  return r;
}
// wa * a + wb * b on RGBX texels (each product and the sum rounded once, like the scalar form)
__device__ __forceinline__ float4 lerp2_texel(float4 a, float4 b, float wa, float wb, float one) {
  const f32x2 WA = pack2(wa, wa), WB = pack2(wb, wb), ONE = pack2(one, one);
  const f32x2 lo = add2(mul2(pack2(a.x, a.y), WA), mul2(pack2(b.x, b.y), WB), ONE);
  const f32x2 hi = add2(mul2(pack2(a.z, a.w), WA), mul2(pack2(b.z, b.w), WB), ONE);
  float4 r;
  unpack2(lo, r.x, r.y);
  unpack2(hi, r.z, r.w);
  return r;
}
// blend3 on packed pairs: 12 FMUL2 + 6 FFMA2 instead of 18 FMUL + 9 FADD
__device__ __forceinline__ void blend3_packed(const float4 v00, const float4 v01, const float4 v10, const float4 v11, float wx1,
                                              float wx0, float wy1, float wy0, float one, float R[3]) {
  const float4 top = lerp2_texel(v00, v01, wx1, wx0, one), bot = lerp2_texel(v10, v11, wx1, wx0, one);
  const float4 r = lerp2_texel(top, bot, wy1, wy0, one);
  R[0] = r.x; R[1] = r.y; R[2] = r.z;
}

// The four taps and bilinear weights of one window pixel; loading is separated from blending so that the loads of
// several pixels can be in flight together.  (ix, iy): source coordinates in the padded window.
struct Taps { float4 v00, v01, v10, v11; float wx1, wx0, wy1, wy0; };
__device__ __forceinline__ Taps load_taps(const Sampler& S, float ix, float iy) {
  Taps t;
  const float x0f = floorf(ix), y0f = floorf(iy);
  t.wx1 = (x0f + 1.0f) - ix; t.wx0 = ix - x0f; t.wy1 = (y0f + 1.0f) - iy; t.wy0 = iy - y0f;
  const int xi = (int)fminf(fmaxf(x0f, S.lo2), S.hi) - S.org;
  const int yi = (int)fminf(fmaxf(y0f, S.lo2), S.hi) - S.org;
  const float4* p = S.u + (yi * S.S + xi);
  t.v00 = p[0]; t.v01 = p[1]; t.v10 = p[S.S]; t.v11 = p[S.S + 1];
  return t;
}

// (ix, iy): source coordinates in the padded window (already divided by the projective term).
__device__ __forceinline__ void sample_at(const Sampler& S, float ix, float iy, float R[3]) {
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = (x0f + 1.0f) - ix, wx0 = ix - x0f, wy1 = (y0f + 1.0f) - iy, wy0 = iy - y0f;
  const int xi = (int)fminf(fmaxf(x0f, S.lo2), S.hi) - S.org;
  const int yi = (int)fminf(fmaxf(y0f, S.lo2), S.hi) - S.org;
  const float4* p = S.u + (yi * S.S + xi);
  blend3(p[0], p[1], p[S.S], p[S.S + 1], wx1, wx0, wy1, wy0, R);
}

#endif  // __CUDACC__

}  // namespace eot
