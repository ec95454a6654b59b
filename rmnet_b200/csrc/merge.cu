// merge.cu -- combine the split-KV partial results, apply the masked-cell correction, scatter into mem_val.
//
// For object o and query cell `pos` of the h x w grid (models/rmnet.py:147-165 with the regional masks of
// :245-248 / :355-358 applied analytically, SURVEY 8a "memory-read regional identity"):
//   in-region query n :  m* = max_s m_s (and 0 if Z > 0);  L = sum_s l_s 2^(m_s - m*) + Z 2^(-m*);
//                        mem[c] = sum_s O_s[c] 2^(m_s - m*) / L
//       Z = number of masked memory cells (score exactly 0, value exactly 0: they only enter the denominator)
//   out-of-region query:  every score is 0  =>  p = 1/M  =>  mem[c] = sum_j V_j[c] / M   (the bank's vsum)
//   mem_val[o, 512 + c, pos] = q_val[c, pos] * att16(o, pos)                              (:358, :163)
// HBM-bound: lanes run along cells (coalesced mem_val writes, coalesced partial reads along compact queries).
#include "common.cuh"

namespace rmnet {
namespace {

constexpr int kMergeThreads = 128;
constexpr int kChPerCta = 8;

__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(BankView bank, const float *__restrict__ q_val, long long q_obj_stride, const int *__restrict__ q_rects,
             int h, int w, int n_obj, int n_splits, const float *__restrict__ opart, const float *__restrict__ ml,
             int nq_pad, float *__restrict__ mem_val) {
  const int N = h * w;
  const int o = blockIdx.z;
  const int pos = blockIdx.x * kMergeThreads + threadIdx.x;
  const int c0 = blockIdx.y * kChPerCta;
  if (pos >= N) return;
  const int4 qrect = q_rects ? __ldg(reinterpret_cast<const int4 *>(q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
  const int cy = pos / w, cx = pos - cy * w;
  const bool in_q = cx >= qrect.x && cx <= qrect.y && cy >= qrect.z && cy <= qrect.w;
  float *out = mem_val + (size_t)o * 2 * RMNET_CV * N + pos;

  // q_val passthrough, channels 512..1023: v4e * att16 (literal multiply keeps the sign of zero like the reference)
  {
    const float att = in_q ? 1.0f : 0.0f;
    const float *qv = q_val + (long long)o * q_obj_stride + pos;
#pragma unroll 4
    for (int c = c0; c < c0 + kChPerCta; ++c) out[(size_t)(RMNET_CV + c) * N] = __ldg(qv + (size_t)c * N) * att;
  }

  const int *meta = bank.meta + o * 8;
  const int Z = meta[META_ZEROS_C] + meta[META_ZEROS_T];
  const int M = Z + meta[META_CELLS_C] + meta[META_CELLS_T];
  const float *vs_c = bank.vsum + (size_t)o * RMNET_CV, *vs_t = bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV;

  if (!in_q) {
    const float inv_m = 1.0f / (float)M;
#pragma unroll 4
    for (int c = c0; c < c0 + kChPerCta; ++c) out[(size_t)c * N] = (vs_c[c] + vs_t[c]) * inv_m;
    return;
  }
  const int n = (cy - qrect.z) * (qrect.y - qrect.x + 1) + (cx - qrect.x);  // compact query index
  const int half = c0 / (RMNET_CV / 2);
  const float2 *mlp = reinterpret_cast<const float2 *>(ml) + ((size_t)o * 2 + half) * nq_pad + n;
  const size_t ml_stride = (size_t)n_obj * 2 * nq_pad;  // between consecutive splits
  // Single pass over the splits, four per round with every load of the round issued before use (the kernel is
  // latency bound).  Online max: the running numerators / denominator are rescaled when a round raises it.
  // Loads of the partial numerators are unconditional -- a split that saw no cells (max = -inf) left them
  // unwritten, so its values are discarded by selection, never multiplied.
  float m_run = Z > 0 ? 0.f : -INFINITY;
  float L = Z > 0 ? (float)Z : 0.f;  // Z * 2^(0 - m_run) with m_run = 0
  float num[kChPerCta];
#pragma unroll
  for (int k = 0; k < kChPerCta; ++k) num[k] = 0.f;
  const size_t op_stride = (size_t)n_obj * RMNET_CV * nq_pad;
  const float *op0 = opart + ((size_t)o * RMNET_CV + c0) * nq_pad + n;
  for (int s0 = 0; s0 < n_splits; s0 += 4) {
    float2 st[4];
    float v[4][kChPerCta];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = min(s0 + u, n_splits - 1);
      st[u] = __ldg(mlp + (size_t)s * ml_stride);
      if (s0 + u >= n_splits) st[u] = make_float2(-INFINITY, 0.f);
#pragma unroll
      for (int k = 0; k < kChPerCta; ++k) v[u][k] = __ldg(op0 + (size_t)s * op_stride + (size_t)k * nq_pad);
    }
    float m_new = m_run;
#pragma unroll
    for (int u = 0; u < 4; ++u) m_new = fmaxf(m_new, st[u].x);
    if (m_new == -INFINITY) continue;  // nothing seen so far
    const float f = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
    L *= f;
#pragma unroll
    for (int k = 0; k < kChPerCta; ++k) num[k] *= f;
    m_run = m_new;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (st[u].x == -INFINITY) continue;
      const float wgt = exp2f(st[u].x - m_run);
      L = fmaf(st[u].y, wgt, L);
#pragma unroll
      for (int k = 0; k < kChPerCta; ++k) num[k] = fmaf(v[u][k], wgt, num[k]);
    }
  }
  const float inv_l = 1.0f / L;
#pragma unroll
  for (int k = 0; k < kChPerCta; ++k) out[(size_t)(c0 + k) * N] = num[k] * inv_l;
}

}  // namespace

int launch_merge(const BankView &bank, const float *q_val, long long q_obj_stride, const int *q_rects, int n_obj, int h,
                 int w, int n_splits, const ReadWorkspace &W, float *mem_val, cudaStream_t st) {
  dim3 grid(cdiv(h * w, kMergeThreads), RMNET_CV / kChPerCta, n_obj);
  merge_kernel<<<grid, kMergeThreads, 0, st>>>(bank, q_val, q_obj_stride, q_rects, h, w, n_obj, n_splits, W.opart, W.ml,
                                               W.nq_pad, mem_val);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // namespace rmnet
