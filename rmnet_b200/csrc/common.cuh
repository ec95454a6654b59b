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


// Loads of data that the PREDECESSOR grid of the chain wrote (it may still be running when this grid starts): they must
// stay behind pdl_wait().  `__ldg` / `const __restrict__` loads are "invariant" to the compiler, which is free to hoist
// them above the wait's memory clobber once their address is known there -- seen in merge.cu when the address arithmetic
// moved in front of the wait: the release build read the previous launch's partial results, a build with a time stamp
// between the wait and the loads did not.  These are volatile asm with a memory clobber (ld.global.cg: L2, no L1 line).
__device__ __forceinline__ int ld_dep(const int *p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_dep(const float *p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long ld_dep(const long long *p) {
  long long v;
  asm volatile("ld.global.cg.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float2 ld_dep(const float2 *p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int2 ld_dep(const int2 *p) {
  int2 v;
  asm volatile("ld.global.cg.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int4 ld_dep(const int4 *p) {
  int4 v;
  asm volatile("ld.global.cg.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_dep(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// The other way to keep such loads behind the wait, for hot loops that want ordinary (schedulable, `__ldg`) loads: pass the
// base pointer through this AFTER pdl_wait().  The loads' addresses then depend on an asm volatile that is ordered after
// the wait, so nothing derived from the returned pointer can be hoisted above it.
template <typename T>
__device__ __forceinline__ T *pdl_launder(T *p) {
  asm volatile("" : "+l"(p) : : "memory");
  return p;
}

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

// ---- development aid (`make DEV=1` only): first-start / last-end %globaltimer stamps of the kernels of the frame-step
//      chain (tools/chain_timeline.py).  In the release build the macros are empty.
#if defined(RMNET_DEV) && defined(__CUDACC__)
__device__ __forceinline__ unsigned long long dev_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RMNET_DEV_STAMPS(tu)                                                  \
  static __device__ unsigned long long *s_chain_stamps = nullptr;             \
  void dev_set_chain_stamps_##tu(unsigned long long *p) { cudaMemcpyToSymbol(s_chain_stamps, &p, sizeof(p)); }
#define DEV_STAMP_MIN(k) do { if (s_chain_stamps && threadIdx.x == 0) atomicMin(s_chain_stamps + (k), dev_globaltimer()); } while (0)
#define DEV_STAMP_MAX(k) do { if (s_chain_stamps && threadIdx.x == 0) atomicMax(s_chain_stamps + (k), dev_globaltimer()); } while (0)
#else
#define RMNET_DEV_STAMPS(tu)
#define DEV_STAMP_MIN(k) do { } while (0)
#define DEV_STAMP_MAX(k) do { } while (0)
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
  // work plan of the tcgen05 read that follows (sched.cuh), built by one extra CTA of this launch; plan_hdr == nullptr: none
  const int *plan_bank_meta;                     // the bank's per-slot counters
  int *plan_ns, *plan_hdr, *plan_pieces;
  int plan_piece_cap, plan_ctas, plan_precision, plan_cap_cells;
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
// The work plan of the persistent tcgen05 kernel (sched.cuh) lives in the workspace too: it is written by the plan role
// of the pack / query-side launch that precedes the read and consumed by the read kernel (piece lists) and merge.cu (ns).
enum { READ_MAX_SPLITS = 16, KV_TILE = 64, MAX_TILES_PER_SPLIT = 64, SCHED_MAX_OBJ = 64, UMMA_QT = 128, PLAN_HDR_CTAS = 256 };
struct ReadWorkspace {
  float *opart, *ml;
  int *sched;           // [SCHED_MAX_OBJ] partial slots per object
  int *plan_hdr;        // [PLAN_HDR_CTAS][2] (pieces, first piece) of every persistent CTA
  int *plan_pieces;     // [plan_piece_cap][4]
  uint16_t *qhi, *qlo;  // [n_obj][nq_pad][128] packed query keys (QuerySide)
  int n_splits, nq_pad;
  size_t plan_piece_cap;
  size_t total;
};
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
  o = align_up(o + SCHED_MAX_OBJ * sizeof(int), 1024);
  W.plan_hdr = (int *)((char *)ws + o);
  o = align_up(o + (size_t)PLAN_HDR_CTAS * 2 * sizeof(int), 1024);
  {  // piece lists: a dealt plan needs G * rounds <= items + G entries, a water-filling plan G * 16 (sched.cuh)
    const size_t deal = (size_t)READ_MAX_SPLITS * 2 * n_obj * (W.nq_pad / 128) + PLAN_HDR_CTAS, fill = (size_t)PLAN_HDR_CTAS * 16;
    W.plan_piece_cap = deal > fill ? deal : fill;
  }
  W.plan_pieces = (int *)((char *)ws + o);
  o = align_up(o + W.plan_piece_cap * 4 * sizeof(int), 1024);
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
