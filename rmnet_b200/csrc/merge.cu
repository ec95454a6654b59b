// merge.cu -- combine the split-KV partial results, apply the masked-cell correction, scatter into mem_val.
//
// For object o and query cell `pos` of the h x w grid (models/rmnet.py:147-165 with the regional masks of
// :245-248 / :355-358 applied analytically, SURVEY 8a "memory-read regional identity"):
//   in-region query n :  m* = max_s m_s (and 0 if Z > 0);  L = sum_s l_s 2^(m_s - m*) + Z 2^(-m*);
//                        mem[c] = sum_s O_s[c] 2^(m_s - m*) / L
//       Z = number of masked memory cells (score exactly 0, value exactly 0: they only enter the denominator)
//   out-of-region query:  every score is 0  =>  p = 1/M  =>  mem[c] = sum_j V_j[c] / M   (the bank's vsum)
//   mem_val[o, 512 + c, pos] = q_val[c, pos] * att16(o, pos)                              (:358, :163)
// Memory / latency bound: lanes run along cells (coalesced mem_val writes, coalesced partial reads along compact
// queries); one thread owns 32 channels of one cell so the per-cell statistics are computed once per 32 outputs,
// the split weights are parked in shared memory and the partial loads are issued two splits x eight channels at a time.
#include "common.cuh"

namespace rmnet {
namespace {

constexpr int kMergeThreads = 128;
constexpr int kChPerCta = 32;
constexpr int kGroup = 8;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(BankView bank, const float *__restrict__ q_val, long long q_obj_stride, const int *__restrict__ q_rects,
             int h, int w, int n_obj, int n_splits, const int *__restrict__ sched_ns, const float *__restrict__ opart,
             const float *__restrict__ ml, int nq_pad, float *__restrict__ mem_val) {
  __shared__ float s_w[READ_MAX_SPLITS][kMergeThreads];  // split weights of this thread's cell
  __shared__ float s_uniform[kChPerCta];                 // sum(V)/M of the CTA's channels (out-of-region read)
  const int N = h * w;
  const int o = blockIdx.z;
  const int pos = blockIdx.x * kMergeThreads + threadIdx.x;
  const int c0 = blockIdx.y * kChPerCta;
  const int tid = threadIdx.x;

  const int *meta = bank.meta + o * 8;
  const int Z = meta[META_ZEROS_C] + meta[META_ZEROS_T];
  const int M = Z + meta[META_CELLS_C] + meta[META_CELLS_T];
  if (tid < kChPerCta) {
    const float *vs_c = bank.vsum + (size_t)o * RMNET_CV, *vs_t = bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV;
    s_uniform[tid] = (vs_c[c0 + tid] + vs_t[c0 + tid]) * (1.0f / (float)M);
  }
  __syncthreads();
  if (pos >= N) return;

  const int4 qrect = q_rects ? __ldg(reinterpret_cast<const int4 *>(q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
  const int cy = pos / w, cx = pos - cy * w;
  const bool in_q = cx >= qrect.x && cx <= qrect.y && cy >= qrect.z && cy <= qrect.w;
  float *out = mem_val + ((size_t)o * 2 * RMNET_CV + c0) * N + pos;

  // q_val passthrough, channels 512..1023: v4e * att16 (literal multiply keeps the sign of zero like the reference)
  {
    const float att = in_q ? 1.0f : 0.0f;
    const float *qv = q_val + (long long)o * q_obj_stride + (size_t)c0 * N + pos;
    float *oq = out + (size_t)RMNET_CV * N;
#pragma unroll
    for (int g = 0; g < kChPerCta; g += kGroup) {
      float x[kGroup];
#pragma unroll
      for (int k = 0; k < kGroup; ++k) x[k] = __ldg(qv + (size_t)(g + k) * N);
#pragma unroll
      for (int k = 0; k < kGroup; ++k) oq[(size_t)(g + k) * N] = x[k] * att;
    }
  }

  if (!in_q) {
#pragma unroll 8
    for (int k = 0; k < kChPerCta; ++k) out[(size_t)k * N] = s_uniform[k];
    return;
  }

  const int n = (cy - qrect.z) * (qrect.y - qrect.x + 1) + (cx - qrect.x);  // compact query index
  const int half = c0 / (RMNET_CV / 2);
  if (sched_ns) n_splits = __ldg(sched_ns + o);  // KV chunks the persistent tcgen05 kernel used for this object
  const float2 *mlp = reinterpret_cast<const float2 *>(ml) + ((size_t)o * 2 + half) * nq_pad + n;
  const size_t ml_stride = (size_t)n_obj * 2 * nq_pad;  // between consecutive splits
  // statistics of every split: reference max, then weights (parked in smem) and the denominator
  float m_star = Z > 0 ? 0.f : -INFINITY;
  for (int s = 0; s < n_splits; ++s) m_star = fmaxf(m_star, __ldg(mlp + (size_t)s * ml_stride).x);
  float L = Z > 0 ? (float)Z * ex2f(-m_star) : 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float2 st = __ldg(mlp + (size_t)s * ml_stride);
    const float wgt = (st.x == -INFINITY) ? 0.f : ex2f(st.x - m_star);  // a split that saw no cells has undefined numerators
    s_w[s][tid] = wgt;
    L = fmaf(st.y, wgt, L);
  }
  const float inv_l = 1.0f / L;
  const size_t op_stride = (size_t)n_obj * RMNET_CV * nq_pad;
  const float *op0 = opart + ((size_t)o * RMNET_CV + c0) * nq_pad + n;
#pragma unroll 1
  for (int g = 0; g < kChPerCta; g += kGroup) {
    float num[kGroup];
#pragma unroll
    for (int k = 0; k < kGroup; ++k) num[k] = 0.f;
    const float *opg = op0 + (size_t)g * nq_pad;
    for (int s = 0; s < n_splits; s += 2) {
      const float w0 = s_w[s][tid], w1 = (s + 1 < n_splits) ? s_w[s + 1][tid] : 0.f;
      const float *p0 = opg + (size_t)s * op_stride, *p1 = p0 + op_stride;
      float v0[kGroup], v1[kGroup];
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        v0[k] = (w0 != 0.f) ? __ldg(p0 + (size_t)k * nq_pad) : 0.f;
        v1[k] = (w1 != 0.f) ? __ldg(p1 + (size_t)k * nq_pad) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < kGroup; ++k) num[k] = fmaf(v1[k], w1, fmaf(v0[k], w0, num[k]));
    }
#pragma unroll
    for (int k = 0; k < kGroup; ++k) out[(size_t)(g + k) * N] = num[k] * inv_l;
  }
}

}  // namespace

int launch_merge(const BankView &bank, const float *q_val, long long q_obj_stride, const int *q_rects, int n_obj, int h,
                 int w, int n_splits, bool device_sched, const ReadWorkspace &W, float *mem_val, cudaStream_t st) {
  dim3 grid(cdiv(h * w, kMergeThreads), RMNET_CV / kChPerCta, n_obj);
  merge_kernel<<<grid, kMergeThreads, 0, st>>>(bank, q_val, q_obj_stride, q_rects, h, w, n_obj, n_splits, device_sched ? W.sched : nullptr, W.opart,
                                               W.ml, W.nq_pad, mem_val);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // namespace rmnet
