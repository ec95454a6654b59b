// common.cuh -- shared host/device helpers of librmnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rmnet_b200.h"

namespace rmnet {

// thread-local error string + launch counter (the only mutable state of the library)
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define RMNET_CHECK_ARG(cond, ...)                 \
  do {                                             \
    if (!(cond)) {                                 \
      ::rmnet::set_error(__VA_ARGS__);             \
      return RMNET_E_INVALID;                      \
    }                                              \
  } while (0)

#define RMNET_CUDA(call)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::rmnet::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                 \
                         cudaGetErrorString(_e));                                             \
      return RMNET_E_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define RMNET_LAUNCH_CHECK()                       \
  do {                                             \
    ::rmnet::count_launch();                       \
    RMNET_CUDA(cudaGetLastError());                \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// The kernels of one frame step form a chain on one stream; each is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that its launch latency and its independent prologue overlap
// the tail of its predecessor.  Device side: pdl_wait() blocks until the predecessor grid has completed and its
// writes are visible; pdl_trigger() lets the successor start launching.  Rule used throughout: a kernel triggers only
// AFTER its own wait, so when a successor's pre-wait code runs, everything two or more links up the chain is complete.
// Both are no-ops for a kernel launched without the attribute.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();  // runtime.cu: false when the environment sets RMNET_DISABLE_PDL=1 (debugging aid)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        bool pdl, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// internal cross-file entry points of the frame-step chain (att_map.cu, bank.cu)
int frame_regions_chain_head(const float *prev_mask, const float *flow, int B, int K, int H, int W, int sampler,
                             float prob_threshold, int n_pts_threshold, int n_bbox_loose_pixels, int pad_l, int pad_r, int pad_t,
                             int pad_b, int k_scan, int *mem_bboxes, int *mem_rects, int *cur_bboxes, int *cur_rects,
                             void *workspace, size_t workspace_bytes, float *clear, int n_clear, void *stream);
// Query side of one read (models/rmnet.py:355-358, :163), produced by roles of the pack kernel (bank.cu):
//   qhi/qlo [n_obj][nq_pad][128]: region-compacted 16-bit hi/lo planes of k4e * att16 (rows up to the next 128 zeroed),
//   mem_val[:, 512:1024] = q_val * att16.
struct QuerySide {
  const float *q_key, *q_val;
  long long q_key_obj_stride, q_val_obj_stride;  // floats; 0 = one query frame shared by all objects (:332-333)
  const int *q_rects;                            // [n_obj,4] cell rectangles, nullptr = dense
  uint16_t *qhi, *qlo;
  int nq_pad;
  int vec4;                                      // 128-bit accesses allowed for q_val / mem_val
  float *mem_val;
  int *range_flag;                               // optional: set to 1 when a query key saturated the fp16 planes (bank meta, slot 0)
};
int bank_memorize_impl(void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *k4, long long k_obj_stride,
                       long long k_ch_stride, const float *v4, long long v_obj_stride, long long v_ch_stride, const int *rects,
                       int n_obj, int h, int w, int elem_format, int commit, bool chained, const QuerySide *query_side,
                       void *stream);
int launch_query_side(const QuerySide &qs, int n_obj, int h, int w, int elem_format, cudaStream_t st);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- memory-bank blob layout (see bank.cu) -------------------------------------------------
// [meta: n_slots x 8 i32][vsum: 2 x n_slots x 512 i64 fixed point][K hi][K lo][V hi][V lo]; sections 1024 B aligned.
struct BankLayout {
  int n_slots, cap;
  size_t off_meta, off_vsum, off_khi, off_klo, off_vhi, off_vlo, total;
};
static inline BankLayout bank_layout(int n_slots, int cap) {
  BankLayout L;
  L.n_slots = n_slots;
  L.cap = cap;
  size_t o = 0;
  L.off_meta = o; o = align_up(o + (size_t)n_slots * 8 * sizeof(int), 1024);
  L.off_vsum = o; o = align_up(o + (size_t)n_slots * 2 * RMNET_CV * sizeof(long long), 1024);
  size_t kplane = (size_t)n_slots * cap * RMNET_CK * 2;
  size_t vplane = (size_t)n_slots * cap * RMNET_CV * 2;
  L.off_khi = o; o = align_up(o + kplane, 1024);
  L.off_klo = o; o = align_up(o + kplane, 1024);
  L.off_vhi = o; o = align_up(o + vplane, 1024);
  L.off_vlo = o; o = align_up(o + vplane, 1024);
  L.total = o;
  return L;
}
// meta ints per slot
enum { META_CELLS_C = 0, META_CELLS_T = 1, META_ZEROS_C = 2, META_ZEROS_T = 3, META_FRAMES_C = 4, META_FRAMES_T = 5, META_OVERFLOW = 6, META_RANGE = 7 };

constexpr float VSUM_SCALE = 16777216.0f;          // 2^24
constexpr float VSUM_INV_SCALE = 1.0f / 16777216.0f;
// Device view of a bank, passed by value to kernels.
struct BankView {
  int *meta;            // [n_slots][8]
  // [2][n_slots][512]  (0 = committed frames, 1 = temporary frame): sum of the stored V per channel, in 2^-24 FIXED POINT.
  // The per-CTA partial sums (fp32, computed in a fixed order) are accumulated with integer atomics, which are
  // associative: the totals -- and with them the uniform rows of mem_val -- are bit-reproducible from run to run, which
  // float atomics are not.  Range: |sum| < 2^39 (fp16-range values of a whole bank stay below 2^31).
  long long *vsum;
  uint16_t *khi, *klo;  // [n_slots][cap][128]
  uint16_t *vhi, *vlo;  // [n_slots][512][cap]
  int n_slots, cap;
};
static inline BankView bank_view(void *bank, int n_slots, int cap) {
  BankLayout L = bank_layout(n_slots, cap);
  char *b = (char *)bank;
  BankView v;
  v.meta = (int *)(b + L.off_meta);
  v.vsum = (long long *)(b + L.off_vsum);
  v.khi = (uint16_t *)(b + L.off_khi);
  v.klo = (uint16_t *)(b + L.off_klo);
  v.vhi = (uint16_t *)(b + L.off_vhi);
  v.vlo = (uint16_t *)(b + L.off_vlo);
  v.n_slots = n_slots;
  v.cap = cap;
  return v;
}

// ---- split-KV partial-result workspace (see memory_read_*.cu, merge.cu) ---------------------
// opart [n_splits][n_obj][512][nq_pad] f32 (unnormalised numerators), ml [n_splits][n_obj][2 halves][nq_pad][2] f32
struct ReadWorkspace {
  float *opart, *ml;
  int *sched;           // [SCHED_MAX_OBJ] partial slots per object, written by the tcgen05 kernel for merge.cu
  uint16_t *qhi, *qlo;  // [n_obj][nq_pad][128] packed query keys (QuerySide)
  int n_splits, nq_pad;
  size_t total;
};
enum { READ_MAX_SPLITS = 16, KV_TILE = 64, MAX_TILES_PER_SPLIT = 64 };
static inline ReadWorkspace read_workspace(void *ws, int n_obj, int N, int n_splits) {
  ReadWorkspace W;
  W.nq_pad = cdiv(N, 128) * 128;
  W.n_splits = n_splits;
  size_t o = 0;
  W.opart = (float *)((char *)ws + o);
  o = align_up(o + (size_t)W.n_splits * n_obj * RMNET_CV * W.nq_pad * sizeof(float), 1024);
  W.ml = (float *)((char *)ws + o);
  o = align_up(o + (size_t)W.n_splits * n_obj * 2 * W.nq_pad * 2 * sizeof(float), 1024);
  W.sched = (int *)((char *)ws + o);
  o = align_up(o + 64 * sizeof(int), 1024);
  W.qhi = (uint16_t *)((char *)ws + o);
  o = align_up(o + (size_t)n_obj * W.nq_pad * RMNET_CK * sizeof(uint16_t), 1024);
  W.qlo = (uint16_t *)((char *)ws + o);
  o = align_up(o + (size_t)n_obj * W.nq_pad * RMNET_CK * sizeof(uint16_t), 1024);
  W.total = o;
  return W;
}
// How many ways the KV axis of every (query tile, object, Cv half) MAY be split (grid.z / 2).  The cell counts live
// on the device, so the host policy is a function of the bank CAPACITY and only an upper bound: on the device a
// split covers  per = max(MIN_TILES_PER_SPLIT, ceil(n_tiles / n_splits))  tiles and surplus CTAs exit at once.
//   * chain bound: at most MAX_TILES_PER_SPLIT tiles are accumulated back to back on the tensor core.  Its fp32
//     accumulator truncates instead of rounding, which biases long same-sign sums by ~3e-8 per accumulation
//     (measured: 1.3e-4 relative over 507 tiles); the split partials are combined by merge.cu in RN fp32.
//   * wave efficiency: CTAs are one-per-SM heavy, so among split counts that keep >= 16 tiles per CTA at full
//     capacity pick the one whose CTA total fills whole waves of 148 SMs best.
enum { MIN_TILES_PER_SPLIT = 8 };
static inline int pick_splits(int n_obj, int N, int q_tile, int cap) {
  const int base = cdiv(N, q_tile) * n_obj * 2;
  const int tiles_max = cdiv(cap, KV_TILE);
  int s_min = cdiv(tiles_max, MAX_TILES_PER_SPLIT);
  if (s_min < 1) s_min = 1;
  if (s_min > READ_MAX_SPLITS) s_min = READ_MAX_SPLITS;
  int s_max = tiles_max / 16;
  if (s_max < s_min) s_max = s_min;
  if (s_max > READ_MAX_SPLITS) s_max = READ_MAX_SPLITS;
  int best = s_min;
  double best_eff = 0.0;
  for (int s = s_min; s <= s_max; ++s) {
    const long long units = (long long)base * s;
    const long long waves = (units + 147) / 148;
    const double eff = (double)units / (double)(waves * 148);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  return best;
}

#ifdef __CUDACC__
// 16-bit hi/lo split of an fp32 value.  fmt 0 = bf16 (8 + 8 mantissa bits: |err| <= 2^-17 |x|, fp32's range),
// fmt 1 = fp16 (11 + 11 bits: |err| <= 2^-23 |x| for |x| in [2^-3, 65504]; absolute error <= 2^-25 below that).
// fp16 planes saturate at +-65504: the return value tells the caller that x was out of range (reported through the
// bank's META_RANGE flag).  NaN passes through as NaN in both formats.
__device__ __forceinline__ bool split16(float x, int fmt, uint16_t &hi, uint16_t &lo) {
  bool sat = false;
  if (fmt == 0) {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    float r = x - __bfloat162float(h);
    __nv_bfloat16 l = __float2bfloat16_rn(r);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
  } else {
    sat = fabsf(x) > 65504.f;
    if (sat) x = copysignf(65504.f, x);
    __half h = __float2half_rn(x);
    float r = x - __half2float(h);
    __half l = __float2half_rn(r);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
  }
  return sat;
}
__device__ __forceinline__ float join16(uint16_t hi, uint16_t lo, int fmt) {
  if (fmt == 0) return __bfloat162float(__ushort_as_bfloat16(hi)) + __bfloat162float(__ushort_as_bfloat16(lo));
  return __half2float(__ushort_as_half(hi)) + __half2float(__ushort_as_half(lo));
}
__device__ __forceinline__ float cvt16(uint16_t v, int fmt) {
  return fmt == 0 ? __bfloat162float(__ushort_as_bfloat16(v)) : __half2float(__ushort_as_half(v));
}
// compact index -> cell position inside an inclusive cell rectangle (cx0,cx1,cy0,cy1) on an h x w grid
__device__ __forceinline__ int rect_cells(const int4 r) {
  int rw = r.y - r.x + 1, rh = r.w - r.z + 1;
  return (rw > 0 && rh > 0) ? rw * rh : 0;
}
__device__ __forceinline__ int rect_pos(const int4 r, int i, int w) {
  int rw = r.y - r.x + 1;
  int cy = r.z + i / rw, cx = r.x + i % rw;
  return cy * w + cx;
}
// ---- work schedule of the persistent tcgen05 kernel (device side, from the actual cell counts) ----------------------
// item = (object o, KV chunk j of ns(o), Cv half, query tile qt): a balanced 1/ns(o) share of the object's nt(o) KV
// tiles for 128 compact queries.  Items are ordered (o, j, half, qt) with qt fastest and dealt round-robin to the
// persistent CTAs, so CTAs that run side by side stream the SAME key/value tiles (one DRAM fetch, L2 hits for the rest).
// ns(o) = ceil(nt(o) / c); the chunk length c is chosen among max_nt / k, k = 1..READ_MAX_SPLITS, to minimise
// rounds(c) * c  (rounds = ceil(#items / #CTAs)), subject to the accumulation-chain bound c <= MAX_TILES_PER_SPLIT.
enum { SCHED_MAX_OBJ = 64, UMMA_QT = 128 };
struct SchedTable {
  int nt[SCHED_MAX_OBJ];         // KV tiles of object o
  int nqt[SCHED_MAX_OBJ];        // query tiles of object o
  int ns[SCHED_MAX_OBJ];         // KV chunks (= partial slots) of object o
  int count[SCHED_MAX_OBJ];      // stored cells of object o (committed + temporary frame)
  int stable[SCHED_MAX_OBJ];     // cells [0, stable) are not being written by a concurrently running pack kernel
  int ibase[SCHED_MAX_OBJ + 1];  // first item of object o
};
// ceil(a / b) for 0 < b, a < 2^20 via one float multiply and a fix-up (a 32-bit integer division costs ~25 instructions)
__device__ __forceinline__ unsigned ceil_div_small(unsigned a, unsigned b, float rcp_b) {
  unsigned q = (unsigned)((float)a * rcp_b);
  q += (q * b < a) ? 1u : 0u;
  q += (q * b < a) ? 1u : 0u;
  q -= (q > 0 && (q - 1u) * b >= a) ? 1u : 0u;
  return q;
}
// Called by ONE FULL WARP (all 32 lanes); lane 0 writes the table.  G = number of persistent CTAs.
// temp_rects != nullptr (rmnet_frame_step without commit): the temporary frame's cell count is derived from its cell
// rectangle exactly as bank_pack_kernel derives it, so the table can be built BEFORE the pack kernel has finished
// (the committed counters do not change during such a step).
__device__ __forceinline__ void sched_build(SchedTable &T, const int *__restrict__ bank_meta, const int *__restrict__ q_rects,
                                            const int *__restrict__ temp_rects, int cap, int n_obj, int h, int w, int G) {
  const int lane = threadIdx.x & 31;
  // per-object tile counts (lanes over objects), also kept in registers for the candidate evaluation
  int max_nt = 0;
  for (int o0 = 0; o0 < n_obj; o0 += 32) {
    const int o = o0 + lane;
    int nt = 0, nqt = 0;
    if (o < n_obj) {
      const int4 qr = q_rects ? __ldg(reinterpret_cast<const int4 *>(q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
      int count;
      if (temp_rects) {
        const int base = bank_meta[o * 8 + META_CELLS_C];
        const int r = rect_cells(__ldg(reinterpret_cast<const int4 *>(temp_rects) + o));
        count = base + (base + r > cap ? 0 : r);  // bank_pack_kernel drops a frame that would overflow the bank
        T.stable[o] = base;
      } else {
        count = bank_meta[o * 8 + META_CELLS_C] + bank_meta[o * 8 + META_CELLS_T];
        T.stable[o] = count;
      }
      nt = (count + KV_TILE - 1) / KV_TILE;
      nqt = (rect_cells(qr) + UMMA_QT - 1) / UMMA_QT;
      T.nt[o] = nt;
      T.nqt[o] = nqt;
      T.count[o] = count;
    }
    max_nt = max(max_nt, __reduce_max_sync(0xffffffffu, nqt > 0 ? nt : 0));
  }
  __syncwarp();
  // candidates: every chunk length c in [c_min, 64] (two per lane), c_min from the partial-slot bound.
  // 32-bit unsigned arithmetic only: 64-bit integer division is emulated with hundreds of instructions.
  const unsigned c_min = max(1u, ((unsigned)max_nt + READ_MAX_SPLITS - 1) / READ_MAX_SPLITS);
  // Both candidates of a lane are evaluated in ONE walk over the objects (independent chains: the walk is latency-
  // bound integer arithmetic, and in a stand-alone launch it sits on the kernel's critical path).
  const unsigned cand[2] = {(unsigned)lane + 1u, (unsigned)lane + 33u};
  const float rcp[2] = {__frcp_rn((float)cand[0]), __frcp_rn((float)cand[1])};
  unsigned items[2] = {0u, 0u}, longest[2] = {0u, 0u};
#pragma unroll 2
  for (int o = 0; o < n_obj; ++o) {
    const unsigned nt = T.nt[o], w2 = 2u * (unsigned)T.nqt[o];
    if (nt > 0 && w2 > 0) {
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const unsigned ns = ceil_div_small(nt, cand[rep], rcp[rep]);
        items[rep] += ns * w2;
        longest[rep] = max(longest[rep], ceil_div_small(nt, ns, __frcp_rn((float)ns)));  // balanced chunks: the longest actual chunk
      }
    }
  }
  unsigned best = 0xffffffffu, best_c = c_min;
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const unsigned c = cand[rep];
    if (c >= c_min && c <= MAX_TILES_PER_SPLIT && c <= (unsigned)max(max_nt, 1)) {
      const unsigned rounds = (items[rep] + G - 1) / (unsigned)G;
      // makespan estimate in tile units: rounds x (longest chunk + per-item prologue/epilogue ~ 5 tiles); ties -> fewer chunks
      const unsigned cost = (rounds * (longest[rep] + 5u)) * 128u + (64u - c);
      if (cost < best) { best = cost; best_c = c; }
    }
  }
  const unsigned bcast = __reduce_min_sync(0xffffffffu, best);
  const unsigned who = __ballot_sync(0xffffffffu, best == bcast);
  int c = (int)__shfl_sync(0xffffffffu, best_c, __ffs(who) - 1);
  if (max_nt > MAX_TILES_PER_SPLIT * READ_MAX_SPLITS) c = (int)c_min;  // huge banks: the slot bound wins over the chain bound
  if (lane == 0) {
    int acc = 0;
    for (int o = 0; o < n_obj; ++o) {
      const int nt = T.nt[o];
      const int ns = (nt > 0 && T.nqt[o] > 0) ? (nt + c - 1) / c : 0;
      T.ns[o] = ns;
      T.ibase[o] = acc;
      acc += ns * 2 * T.nqt[o];
    }
    T.ibase[n_obj] = acc;
  }
  __syncwarp();
}

// device-side resolution of the split -> [tile_begin, tile_begin + n_it) for `count` stored cells
__device__ __forceinline__ void split_range(int count, int n_splits, int split, int &tile_begin, int &n_it) {
  const int n_tiles = (count + KV_TILE - 1) / KV_TILE;
  int per = (n_tiles + n_splits - 1) / n_splits;
  if (per < MIN_TILES_PER_SPLIT) per = MIN_TILES_PER_SPLIT;
  tile_begin = split * per;
  n_it = min(n_tiles, tile_begin + per) - tile_begin;
  if (n_it < 0) n_it = 0;
}
#endif

}  // namespace rmnet
