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
namespace {

constexpr int kCellsPerCta = 64;
constexpr int kPackThreads = 256;
constexpr int kChunk = 128;                          // channels per CTA: blockIdx.y = 0 -> keys, 1..4 -> value chunks
constexpr int kKRowStride = RMNET_CK + 4;            // ushorts; keeps 8-byte row alignment, spreads banks
constexpr int kGroups = kPackThreads / kCellsPerCta; // 4 channel groups; lanes run along cells (coalesced gathers)
constexpr int kPerThread = kChunk / kGroups;         // 32 channels per thread

template <int FMT>
__global__ void __launch_bounds__(kPackThreads)
bank_pack_kernel(BankView bank, const float *__restrict__ k4, long long k_obj_stride, long long k_ch_stride,
                 const float *__restrict__ v4, long long v_obj_stride, long long v_ch_stride,
                 const int *__restrict__ rects, int h, int w) {
  __shared__ __align__(16) uint16_t s_hi[kCellsPerCta][kKRowStride];
  __shared__ __align__(16) uint16_t s_lo[kCellsPerCta][kKRowStride];
  __shared__ float s_vsum[kChunk];

  const int o = blockIdx.z;
  const int4 rect = __ldg(reinterpret_cast<const int4 *>(rects) + o);
  const int r = rect_cells(rect);
  int *meta = bank.meta + o * 8;
  const int base = meta[META_CELLS_C];
  const bool overflow = base + r > bank.cap;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    meta[META_CELLS_T] = overflow ? 0 : r;
    meta[META_ZEROS_T] = overflow ? h * w : h * w - r;
    meta[META_FRAMES_T] = 1;
    if (overflow) meta[META_OVERFLOW] = 1;
  }
  const int i0 = blockIdx.x * kCellsPerCta;
  if (i0 >= r || overflow) return;
  const int cnt = min(kCellsPerCta, r - i0);

  const int li = threadIdx.x & (kCellsPerCta - 1);  // cell within the tile
  const int cg = threadIdx.x / kCellsPerCta;        // channel group
  const bool live = li < cnt;
  const int pos = live ? rect_pos(rect, i0 + li, w) : 0;

  if (blockIdx.y == 0) {
    // ---- keys: gather -> split -> transpose through smem -> 256 B position-major rows
    const float *kp = k4 + (long long)o * k_obj_stride + pos;
    float x[kPerThread];
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) x[k] = live ? __ldg(kp + (long long)(cg + kGroups * k) * k_ch_stride) : 0.f;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
      uint16_t hi, lo;
      split16(x[k], FMT, hi, lo);
      s_hi[li][cg + kGroups * k] = hi;
      s_lo[li][cg + kGroups * k] = lo;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    for (int row = wrp; row < cnt; row += kPackThreads / 32) {
      const size_t g = ((size_t)o * bank.cap + base + i0 + row) * RMNET_CK + lane * 4;
      *reinterpret_cast<uint2 *>(bank.khi + g) = *reinterpret_cast<const uint2 *>(&s_hi[row][lane * 4]);
      *reinterpret_cast<uint2 *>(bank.klo + g) = *reinterpret_cast<const uint2 *>(&s_lo[row][lane * 4]);
    }
    return;
  }

  // ---- values: gather -> split -> channel-major rows (lanes along cells), plus per-channel sums of the chunk
  const int cbase = (blockIdx.y - 1) * kChunk;
  if (threadIdx.x < kChunk) s_vsum[threadIdx.x] = 0.f;
  __syncthreads();
  const float *vp = v4 + (long long)o * v_obj_stride + pos;
  const size_t vrow0 = (size_t)o * RMNET_CV * bank.cap + base + i0 + li;
  float x[kPerThread];
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) x[k] = live ? __ldg(vp + (long long)(cbase + cg + kGroups * k) * v_ch_stride) : 0.f;
#pragma unroll
  for (int k = 0; k < kPerThread; ++k) {
    const int c = cbase + cg + kGroups * k;
    if (live) {
      uint16_t hi, lo;
      split16(x[k], FMT, hi, lo);
      bank.vhi[vrow0 + (size_t)c * bank.cap] = hi;
      bank.vlo[vrow0 + (size_t)c * bank.cap] = lo;
    }
    float sum = x[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_vsum[c - cbase], sum);  // two warps share a channel
  }
  __syncthreads();
  if (threadIdx.x < kChunk)
    atomicAdd(bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV + cbase + threadIdx.x, s_vsum[threadIdx.x]);
}

// `keys = this_keys` (models/rmnet.py:424-426): the temporary frame becomes permanent.
__global__ void bank_commit_kernel(BankView bank, int n_obj) {
  const int o = blockIdx.x;
  if (o >= n_obj) return;
  float *vc = bank.vsum + (size_t)o * RMNET_CV, *vt = bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV;
  for (int c = threadIdx.x; c < RMNET_CV; c += blockDim.x) { vc[c] += vt[c]; vt[c] = 0.f; }
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
  RMNET_CHECK_ARG(bank && k4 && v4 && rects, "null pointer argument");
  RMNET_CHECK_ARG(n_obj > 0 && n_obj <= n_slots && h > 0 && w > 0, "bad shape n_obj=%d n_slots=%d h=%d w=%d", n_obj, n_slots, h, w);
  RMNET_CHECK_ARG(elem_format == 0 || elem_format == 1, "elem_format must be 0 (bf16) or 1 (fp16)");
  RMNET_CHECK_ARG((uintptr_t)rects % 16 == 0, "rects must be 16-byte aligned");
  BankLayout L = bank_layout(n_slots, cap_cells);
  if (bank_bytes < L.total) { set_error("bank too small: %zu < %zu", bank_bytes, L.total); return RMNET_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  BankView bv = bank_view(bank, n_slots, cap_cells);
  // vsum of the temporary frame restarts from zero
  RMNET_CUDA(cudaMemsetAsync(bv.vsum + (size_t)n_slots * RMNET_CV, 0, (size_t)n_slots * RMNET_CV * sizeof(float), st));
  dim3 grid(cdiv(h * w, kCellsPerCta), 1 + RMNET_CV / kChunk, n_obj);
  if (elem_format == 0)
    bank_pack_kernel<0><<<grid, kPackThreads, 0, st>>>(bv, k4, k_obj_stride, k_ch_stride, v4, v_obj_stride, v_ch_stride, rects, h, w);
  else
    bank_pack_kernel<1><<<grid, kPackThreads, 0, st>>>(bv, k4, k_obj_stride, k_ch_stride, v4, v_obj_stride, v_ch_stride, rects, h, w);
  RMNET_LAUNCH_CHECK();
  if (commit) {
    bank_commit_kernel<<<n_obj, 128, 0, st>>>(bv, n_obj);
    RMNET_LAUNCH_CHECK();
  }
  return RMNET_OK;
}

int rmnet_bank_stats_host(const void *bank, int n_slots, int cap_cells, int *out_host, void *stream) {
  RMNET_CHECK_ARG(bank && out_host && n_slots > 0 && cap_cells > 0, "bad argument");
  BankLayout L = bank_layout(n_slots, cap_cells);
  RMNET_CUDA(cudaMemcpyAsync(out_host, (const char *)bank + L.off_meta, (size_t)n_slots * 8 * sizeof(int),
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  RMNET_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return RMNET_OK;
}
}
