// bank.cu -- the preallocated, region-compacted memory bank ("pack at memorise").
//
// Replaces models/rmnet.py:191-205 (pad_memory: zeroed [B,K,C,1,h,w] + scatter), :245-248 (nearest /16 +
// k4*att, v4*att), :416-426 (torch.cat growth: the whole bank re-copied every frame) and :348-349 (the
// per-object gather before the read).  One pass over the new frame's K/V (HBM-bound: 4 B read + 4 B written
// per in-region element; masked cells are never stored, only counted):
//   keys   -> position-major rows  khi/klo[slot][cell][128]   (K-major UMMA / TMA operand, 256 B rows)
//   values -> channel-major rows   vhi/vlo[slot][512][cell]   (K-major B operand of the P.V product)
//   16-bit hi/lo split (x ~= hi + lo) so the read kernel needs no operand conversion stage,
//   per-channel sum of the stored values (the read of an out-of-region query is sum(V)/M).
#include "common.cuh"
namespace rmnet {
RMNET_DEV_STAMPS(bank)
}
#include "sched.cuh"

namespace rmnet {
namespace {

constexpr int kCellsPerCta = 64;
constexpr int kPackThreads = 256;
constexpr int kChunk = 128;                          // channels per CTA
constexpr int kKRowStride = RMNET_CK + 4;            // ushorts; keeps 8-byte row alignment, spreads banks
constexpr int kGroups = kPackThreads / kCellsPerCta; // 4 channel groups; lanes run along cells (coalesced gathers)
constexpr int kPerThread = kChunk / kGroups;         // 32 channels per thread

// CTA roles: the memory side of models/rmnet.py:239-248 and the query side of :355-358, :163
enum { ROLE_MEM_KEYS = 0, ROLE_MEM_VALS = 1 /* ..4 */, ROLE_Q_KEYS = 5, ROLE_Q_PASS = 6 /* ..9 */, ROLE_PLAN = 10 /* one CTA: sched.cuh */ };
// The grid is (cell tiles, objects, roles) with the roles SLOWEST and in the order the step needs them: CTAs are dispatched
// in index order and only four fit an SM, so with the objects slowest (rounds 1-2a) the last objects' memory roles started
// a wave late (the last object's memory values ended the pack stage); now the plan CTA is the very first one, then the
// memory roles of ALL objects (the slowest: gather + split + scattered 2-byte stores + value sums), the query keys, and
// the bandwidth-bound q_val passthrough last.
enum { PACK_MEMORIZE = 0 /* memory roles */, PACK_FRAME = 1 /* plan + memory + query roles */, PACK_QUERY = 2 /* plan + query roles */ };
__device__ __forceinline__ int pack_role(int kind, int z) {
  if (kind == PACK_MEMORIZE) return z < 4 ? ROLE_MEM_VALS + z : ROLE_MEM_KEYS;
  if (kind == PACK_QUERY) return z == 0 ? ROLE_PLAN : (z == 1 ? ROLE_Q_KEYS : ROLE_Q_PASS + (z - 2));
  return z == 0 ? ROLE_PLAN : (z <= 4 ? ROLE_MEM_VALS + (z - 1) : (z == 5 ? ROLE_MEM_KEYS : (z == 6 ? ROLE_Q_KEYS : ROLE_Q_PASS + (z - 7))));
}

struct PackSmem {
  __align__(16) uint16_t hi[kCellsPerCta][kKRowStride];
  __align__(16) uint16_t lo[kCellsPerCta][kKRowStride];
  float vsum[kChunk];
};
static_assert(sizeof(PlanSmem) <= sizeof(PackSmem), "the plan role reuses the pack CTA's shared memory");

// 64 compact cells x 128 key channels: gather from the channels-first frame -> 16-bit hi/lo split -> transpose through
// smem -> 256 B position-major rows.  Rows [cnt, rows) are written as zeros.
// INTERLEAVED = false: plain rows (the TMA / UMMA operand layout of the bank).
// INTERLEAVED = true : query planes for the read kernel, whose thread r of a warp owns row r of a 32-row group and loads
//   it 16 B at a time: chunk j of the 32 rows is stored contiguously ([group][16 chunks][32 rows][8 ch]) so that
//   every such warp load is one coalesced 512 B access.  dst_* point at the first row of a 32-aligned group.
template <int FMT, bool INTERLEAVED>
__device__ __forceinline__ bool pack_key_rows(PackSmem &sm, const float *__restrict__ src, long long ch_stride, const int4 rect,
                                              int w, int i0, int cnt, int rows, uint16_t *__restrict__ dst_hi,
                                              uint16_t *__restrict__ dst_lo) {
  const int li = threadIdx.x & (kCellsPerCta - 1);  // cell within the tile
  const int cg = threadIdx.x / kCellsPerCta;        // channel group
  const bool live = li < cnt;
  const float *kp = src + (live ? rect_pos(rect, i0 + li, w) : 0);
  float x[kPerThread];
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) x[k] = live ? __ldg(kp + (long long)(cg + kGroups * k) * ch_stride) : 0.f;
  bool sat = false;
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    uint16_t hi, lo;
    sat |= split16(x[k], FMT, hi, lo);
    sm.hi[li][cg + kGroups * k] = hi;
    sm.lo[li][cg + kGroups * k] = lo;
  }
  sat = FMT == 1 ? (__syncthreads_or(sat) != 0) : (__syncthreads(), false);
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  if (INTERLEAVED) {
    // warp `wrp` writes chunks 2*wrp, 2*wrp + 1 of both 32-row groups of the tile; lane = row within the group
    for (int g32 = 0; g32 * 32 < rows; ++g32) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = wrp * 2 + jj, row = g32 * 32 + lane;
        const uint2 h0 = *reinterpret_cast<const uint2 *>(&sm.hi[row][j * 8]), h1 = *reinterpret_cast<const uint2 *>(&sm.hi[row][j * 8 + 4]);
        const uint2 l0 = *reinterpret_cast<const uint2 *>(&sm.lo[row][j * 8]), l1 = *reinterpret_cast<const uint2 *>(&sm.lo[row][j * 8 + 4]);
        const size_t g = ((size_t)(g32 * 16 + j) * 32 + lane) * 8;  // ushorts
        *reinterpret_cast<uint4 *>(dst_hi + g) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4 *>(dst_lo + g) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
    }
  } else {
    for (int row = wrp; row < rows; row += kPackThreads / 32) {
      const size_t g = (size_t)row * RMNET_CK + lane * 4;
      *reinterpret_cast<uint2 *>(dst_hi + g) = *reinterpret_cast<const uint2 *>(&sm.hi[row][lane * 4]);
      *reinterpret_cast<uint2 *>(dst_lo + g) = *reinterpret_cast<const uint2 *>(&sm.lo[row][lane * 4]);
    }
  }
  return sat;
}

template <int FMT>
__global__ void __launch_bounds__(kPackThreads)
bank_pack_kernel(BankView bank, const float *__restrict__ k4, long long k_obj_stride, long long k_ch_stride,
                 const float *__restrict__ v4, long long v_obj_stride, long long v_ch_stride,
                 const int *__restrict__ rects, QuerySide qs, int kind, int h, int w) {
  __shared__ PackSmem sm;
  DEV_STAMP_MIN(2);
  const int o = blockIdx.y;
  const int role = pack_role(kind, (int)blockIdx.z);
  if (role == ROLE_PLAN && (blockIdx.x != 0 || o != 0 || !qs.plan_hdr)) return;  // the plan is ONE CTA (the first of the grid)
  pdl_wait();     // chained launch: the rectangles come from the region kernel right before us
  pdl_trigger();  // (after the wait: the successor's prologue may then rely on everything before this kernel)
  DEV_STAMP_MIN(3);
  const int N = h * w;

  if (role == ROLE_PLAN) {
    // ---- work plan of the tcgen05 read that follows this launch (one CTA; sched.cuh).  With memory roles in the launch
    //      the temporary frame's cell counts come from `rects`, exactly as the memory roles below derive them.
    plan_build(*reinterpret_cast<PlanSmem *>(&sm), qs.plan_bank_meta, qs.q_rects, kind == PACK_FRAME ? rects : nullptr,
               qs.plan_cap_cells, (int)gridDim.y, h, w, qs.plan_ctas, qs.plan_precision, qs.plan_ns,
               reinterpret_cast<int2 *>(qs.plan_hdr), reinterpret_cast<int4 *>(qs.plan_pieces), qs.plan_piece_cap);
    DEV_STAMP_MAX(11);
    return;
  }
  if (role >= ROLE_Q_KEYS) {
    const int4 qrect = qs.q_rects ? ld_dep(reinterpret_cast<const int4 *>(qs.q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
    if (role == ROLE_Q_KEYS) {
      // ---- query keys (k4e * att16, :357): compact 16-bit planes [o][nq_pad][128] (32-row interleaved, see
      //      pack_key_rows), zero rows up to the next 128
      const int r = rect_cells(qrect);
      const int r_pad = min(qs.nq_pad, (r + 127) / 128 * 128);
      const int i0 = blockIdx.x * kCellsPerCta;
      if (i0 >= r_pad) return;
      const size_t row0 = ((size_t)o * qs.nq_pad + i0) * RMNET_CK;
      // r_pad is a multiple of 128: the tile's 64 rows are two whole 32-row groups
      const bool sat = pack_key_rows<FMT, true>(sm, qs.q_key + (long long)o * qs.q_key_obj_stride, (long long)N, qrect, w, i0,
                                                max(0, min(kCellsPerCta, r - i0)), kCellsPerCta, qs.qhi + row0, qs.qlo + row0);
      if (sat && qs.range_flag && threadIdx.x == 0) atomicOr(qs.range_flag, 1);
      DEV_STAMP_MAX(4);
      return;
    }
    // ---- q_val passthrough (v4e * att16 into channels 512..1023 of mem_val, :358 + :163): 64 cells x 128 channels.
    //      A literal multiply by {0,1}: keeps the sign of zero (and NaN/Inf) exactly like the reference.
    const int c0 = (role - ROLE_Q_PASS) * kChunk;
    const int p_base = blockIdx.x * kCellsPerCta;
    if (p_base >= N) return;
    const float *qv = qs.q_val + (long long)o * qs.q_val_obj_stride;
    float *out = qs.mem_val + ((size_t)o * 2 * RMNET_CV + RMNET_CV) * N;
    if (qs.vec4) {
      const int pa = p_base + (threadIdx.x & 15) * 4;
      if (pa >= N) return;
      float in_q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int pos = pa + e, cy = pos / w, cx = pos - cy * w;
        in_q[e] = (cx >= qrect.x && cx <= qrect.y && cy >= qrect.z && cy <= qrect.w) ? 1.0f : 0.0f;
      }
      const int cb = c0 + (threadIdx.x >> 4) * 8;
      float4 x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = __ldg(reinterpret_cast<const float4 *>(qv + (size_t)(cb + k) * N + pa));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        x[k].x *= in_q[0]; x[k].y *= in_q[1]; x[k].z *= in_q[2]; x[k].w *= in_q[3];
        *reinterpret_cast<float4 *>(out + (size_t)(cb + k) * N + pa) = x[k];
      }
      DEV_STAMP_MAX(4);
    } else {
      const int pos = p_base + (threadIdx.x & (kCellsPerCta - 1));
      if (pos >= N) return;
      const int cy = pos / w, cx = pos - cy * w;
      const float in_q = (cx >= qrect.x && cx <= qrect.y && cy >= qrect.z && cy <= qrect.w) ? 1.0f : 0.0f;
      for (int c = c0 + threadIdx.x / kCellsPerCta; c < c0 + kChunk; c += kGroups)
        out[(size_t)c * N + pos] = __ldg(qv + (size_t)c * N + pos) * in_q;
    }
    return;
  }

  const int4 rect = ld_dep(reinterpret_cast<const int4 *>(rects) + o);  // (from the region kernel, our predecessor)
  const int r = rect_cells(rect);
  int *meta = bank.meta + o * 8;
  const int base = meta[META_CELLS_C];
  const bool overflow = base + r > bank.cap;
  if (blockIdx.x == 0 && role == ROLE_MEM_KEYS && threadIdx.x == 0) {
    meta[META_CELLS_T] = overflow ? 0 : r;
    meta[META_ZEROS_T] = overflow ? h * w : h * w - r;
    meta[META_FRAMES_T] = 1;
    if (overflow) meta[META_OVERFLOW] = 1;
  }
  const int i0 = blockIdx.x * kCellsPerCta;
  if (i0 >= r || overflow) return;
  const int cnt = min(kCellsPerCta, r - i0);

  if (role == ROLE_MEM_KEYS) {
    const size_t row0 = ((size_t)o * bank.cap + base + i0) * RMNET_CK;
    if (pack_key_rows<FMT, false>(sm, k4 + (long long)o * k_obj_stride, k_ch_stride, rect, w, i0, cnt, cnt, bank.khi + row0, bank.klo + row0) &&
        threadIdx.x == 0)
      atomicOr(meta + META_RANGE, 1);
    DEV_STAMP_MAX(4);
    return;
  }

  // ---- values: gather -> split -> channel-major rows (lanes along cells), plus per-channel sums of the chunk
  const int li = threadIdx.x & (kCellsPerCta - 1);  // cell within the tile
  const int cg = threadIdx.x / kCellsPerCta;        // channel group
  const bool live = li < cnt;
  const int pos = live ? rect_pos(rect, i0 + li, w) : 0;
  const int cbase = (role - ROLE_MEM_VALS) * kChunk;
  if (threadIdx.x < kChunk) sm.vsum[threadIdx.x] = 0.f;
  __syncthreads();
  const float *vp = v4 + (long long)o * v_obj_stride + pos;
  const size_t vrow0 = (size_t)o * RMNET_CV * bank.cap + base + i0 + li;
  float x[kPerThread];
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) x[k] = live ? __ldg(vp + (long long)(cbase + cg + kGroups * k) * v_ch_stride) : 0.f;
  bool sat = false;
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    if (live) {
      const int c = cbase + cg + kGroups * k;
      uint16_t hi, lo;
      sat |= split16(x[k], FMT, hi, lo);
      bank.vhi[vrow0 + (size_t)c * bank.cap] = hi;
      bank.vlo[vrow0 + (size_t)c * bank.cap] = lo;
    }
  }
  if (FMT == 1 && sat) atomicOr(meta + META_RANGE, 1);   // (never taken for in-range features)
  // Per-channel sums over the warp's 32 cells by a transpose-reduction: in step s every lane keeps half of its values
  // and adds the partner's copy of the same half (31 shuffles instead of 32 x 5); lane l ends up with the total of its
  // l-th channel.  Dead lanes hold zeros.
  static_assert(kPerThread == 32, "the transpose-reduction below assumes 32 channels per thread");
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int sft = 16; sft >= 1; sft >>= 1) {
    const bool upper = (lane & sft) != 0;
#pragma unroll
    for (int j = 0; j < sft; ++j) {
      const float keep = upper ? x[j + sft] : x[j];
      const float send = upper ? x[j] : x[j + sft];
      x[j] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
    }
  }
  atomicAdd(&sm.vsum[cg + kGroups * lane], x[0]);  // the two warps of a channel group meet here
  __syncthreads();
  // (sm.vsum[c] = 0 + a + b of the group's two warps: float addition is commutative, so the CTA partial is reproducible;
  //  the cross-CTA accumulation is an integer atomic on the 2^-24 fixed-point image of the partial -- see BankView::vsum)
  if (threadIdx.x < kChunk)
    atomicAdd(reinterpret_cast<unsigned long long *>(bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV + cbase + threadIdx.x),
              (unsigned long long)__float2ll_rn(sm.vsum[threadIdx.x] * VSUM_SCALE));
  DEV_STAMP_MAX(4);
}

// `keys = this_keys` (models/rmnet.py:424-426): the temporary frame becomes permanent.
__global__ void bank_commit_kernel(BankView bank, int n_obj) {
  pdl_wait();
  pdl_trigger();
  const int o = blockIdx.x;
  if (o >= n_obj) return;
  long long *vc = bank.vsum + (size_t)o * RMNET_CV, *vt = bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV;
  for (int c = threadIdx.x; c < RMNET_CV; c += blockDim.x) { vc[c] += vt[c]; vt[c] = 0; }
  if (threadIdx.x == 0) {
    int *m = bank.meta + o * 8;
    m[META_CELLS_C] += m[META_CELLS_T];
    m[META_ZEROS_C] += m[META_ZEROS_T];
    m[META_FRAMES_C] += m[META_FRAMES_T];
    m[META_CELLS_T] = 0; m[META_ZEROS_T] = 0; m[META_FRAMES_T] = 0;
  }
}

}  // namespace
}  // namespace rmnet

using namespace rmnet;
// Query side alone (standalone rmnet_bank_memory_read): roles 5..9 of the pack kernel.
int rmnet::launch_query_side(const QuerySide &qs, int n_obj, int h, int w, int elem_format, cudaStream_t st) {
  BankView none = {};
  dim3 grid(cdiv(cdiv(h * w, 128) * 128, kCellsPerCta), n_obj, 6);
  if (elem_format == 0)
    RMNET_CUDA(launch_kernel(bank_pack_kernel<0>, grid, dim3(kPackThreads), 0, st, false, none, (const float *)nullptr, 0LL, 0LL,
                             (const float *)nullptr, 0LL, 0LL, (const int *)nullptr, qs, (int)PACK_QUERY, h, w));
  else
    RMNET_CUDA(launch_kernel(bank_pack_kernel<1>, grid, dim3(kPackThreads), 0, st, false, none, (const float *)nullptr, 0LL, 0LL,
                             (const float *)nullptr, 0LL, 0LL, (const int *)nullptr, qs, (int)PACK_QUERY, h, w));
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

extern "C" {

size_t rmnet_bank_bytes(int n_slots, int cap_cells) {
  if (n_slots <= 0 || cap_cells <= 0) return 0;
  return bank_layout(n_slots, cap_cells).total;
}

int rmnet_bank_reset(void *bank, size_t bank_bytes, int n_slots, int cap_cells, void *stream) {
  RMNET_CHECK_ARG(bank && n_slots > 0 && cap_cells > 0, "bad argument");
  RMNET_CHECK_ARG(cap_cells % 8 == 0, "cap_cells must be a multiple of 8 (16-byte TMA row pitch)");
  RMNET_CHECK_ARG((uintptr_t)bank % 1024 == 0, "bank must be 1024-byte aligned");
  BankLayout L = bank_layout(n_slots, cap_cells);
  if (bank_bytes < L.total) { set_error("bank too small: %zu < %zu", bank_bytes, L.total); return RMNET_E_WORKSPACE; }
  // zero everything: the V planes must stay finite beyond the stored cells (the last KV tile multiplies them by p = 0)
  RMNET_CUDA(cudaMemsetAsync(bank, 0, L.total, (cudaStream_t)stream));
  return RMNET_OK;
}

int rmnet_bank_memorize(void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *k4,
                        long long k_obj_stride, long long k_ch_stride, const float *v4, long long v_obj_stride,
                        long long v_ch_stride, const int *rects, int n_obj, int h, int w, int elem_format, int commit,
                        void *stream) {
  return bank_memorize_impl(bank, bank_bytes, n_slots, cap_cells, k4, k_obj_stride, k_ch_stride, v4, v_obj_stride, v_ch_stride,
                            rects, n_obj, h, w, elem_format, commit, /*chained=*/false, nullptr, stream);
}
}  // extern "C"

// chained = true (rmnet_frame_step): the launches are programmatic dependents of the region kernel, which has already
// zeroed the temporary frame's value sums; otherwise a memset node does that and the launches are ordinary.
int rmnet::bank_memorize_impl(void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *k4,
                              long long k_obj_stride, long long k_ch_stride, const float *v4, long long v_obj_stride,
                              long long v_ch_stride, const int *rects, int n_obj, int h, int w, int elem_format, int commit,
                              bool chained, const QuerySide *query_side, void *stream) {
  RMNET_CHECK_ARG(bank && k4 && v4 && rects, "null pointer argument");
  RMNET_CHECK_ARG(n_obj > 0 && n_obj <= n_slots && h > 0 && w > 0, "bad shape n_obj=%d n_slots=%d h=%d w=%d", n_obj, n_slots, h, w);
  RMNET_CHECK_ARG(elem_format == 0 || elem_format == 1, "elem_format must be 0 (bf16) or 1 (fp16)");
  RMNET_CHECK_ARG((uintptr_t)rects % 16 == 0, "rects must be 16-byte aligned");
  BankLayout L = bank_layout(n_slots, cap_cells);
  if (bank_bytes < L.total) { set_error("bank too small: %zu < %zu", bank_bytes, L.total); return RMNET_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  BankView bv = bank_view(bank, n_slots, cap_cells);
  // vsum of the temporary frame restarts from zero
  if (!chained)
    RMNET_CUDA(cudaMemsetAsync(bv.vsum + (size_t)n_slots * RMNET_CV, 0, (size_t)n_slots * RMNET_CV * sizeof(long long), st));
  // with a query side (rmnet_frame_step) the same launch also packs the query keys and writes the q_val passthrough
  QuerySide qs = {};
  if (query_side) qs = *query_side;
  dim3 grid(cdiv(cdiv(h * w, 128) * 128, kCellsPerCta), n_obj, query_side ? 11 : 5);
  const int kind = query_side ? PACK_FRAME : PACK_MEMORIZE;
  if (elem_format == 0)
    RMNET_CUDA(launch_kernel(bank_pack_kernel<0>, grid, dim3(kPackThreads), 0, st, chained, bv, k4, k_obj_stride, k_ch_stride, v4,
                             v_obj_stride, v_ch_stride, rects, qs, kind, h, w));
  else
    RMNET_CUDA(launch_kernel(bank_pack_kernel<1>, grid, dim3(kPackThreads), 0, st, chained, bv, k4, k_obj_stride, k_ch_stride, v4,
                             v_obj_stride, v_ch_stride, rects, qs, kind, h, w));
  RMNET_LAUNCH_CHECK();
  if (commit) {
    RMNET_CUDA(launch_kernel(bank_commit_kernel, dim3(n_obj), dim3(128), 0, st, chained, bv, n_obj));
    RMNET_LAUNCH_CHECK();
  }
  return RMNET_OK;
}
extern "C" {

int rmnet_bank_stats_host(const void *bank, int n_slots, int cap_cells, int *out_host, void *stream) {
  RMNET_CHECK_ARG(bank && out_host && n_slots > 0 && cap_cells > 0, "bad argument");
  BankLayout L = bank_layout(n_slots, cap_cells);
  RMNET_CUDA(cudaMemcpyAsync(out_host, (const char *)bank + L.off_meta, (size_t)n_slots * 8 * sizeof(int),
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  RMNET_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  for (int s = 0; s < n_slots; ++s)
    if (out_host[s * 8 + META_RANGE]) {
      set_error("fp16 planes saturated: slot %d stored a key / value (or packed a query key) beyond +-65504; use elem_format 0 (bf16 planes)", s);
      return RMNET_E_UNSUPPORTED;
    }
  for (int s = 0; s < n_slots; ++s)
    if (out_host[s * 8 + META_OVERFLOW]) {
      set_error("memory bank overflow: slot %d dropped a frame (cap_cells = %d, %d committed cells)", s, cap_cells,
                out_host[s * 8 + META_CELLS_C]);
      return RMNET_E_WORKSPACE;
    }
  return RMNET_OK;
}
}
