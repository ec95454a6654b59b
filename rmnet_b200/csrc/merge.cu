// merge.cu -- combine the split-KV partial results, apply the masked-cell correction, scatter into mem_val.
//
// For object o and query cell `pos` of the h x w grid (models/rmnet.py:147-165 with the regional masks of
// :245-248 / :355-358 applied analytically, SURVEY 8a "memory-read regional identity"):
//   in-region query n :  m* = max_s m_s (and 0 if Z > 0);  L = sum_s l_s 2^(m_s - m*) + Z 2^(-m*);
//                        mem[c] = sum_s O_s[c] 2^(m_s - m*) / L
//       Z = number of masked memory cells (score exactly 0, value exactly 0: they only enter the denominator)
//   out-of-region query:  every score is 0  =>  p = 1/M  =>  mem[c] = sum_j V_j[c] / M   (the bank's vsum)
//   mem_val[o, 512 + c, pos] = q_val[c, pos] * att16(o, pos)                              (:358, :163)
// Memory / latency bound.  A thread owns VEC (4) consecutive cells x 8 channels: 128-bit loads / stores along the
// cells for the q_val passthrough and the uniform rows, scalar gathers of the partial numerators only for in-region cells.
#include "common.cuh"

namespace rmnet {
namespace {

constexpr int kMergeThreads = 128;
constexpr int kChPerCta = 32;
constexpr int kChPerThread = 8;  // 4 warps x 8 channels

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// CTA = 128 cells x 32 channels of one object.
//   phase A (fill): thread = VEC consecutive cells x 8 channels, 128-bit accesses: q_val passthrough (all cells) and
//                   the uniform rows (out-of-region cells only);
//   phase B (gather): thread = ONE in-region cell x 32 channels: statistics of all splits in one batch of loads,
//                   then the partial numerators two splits x eight channels at a time (independent loads in flight).
template <int VEC>
__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(BankView bank, const float *__restrict__ q_val, long long q_obj_stride, const int *__restrict__ q_rects,
             int h, int w, int n_obj, int n_splits, const int *__restrict__ sched_ns, const float *__restrict__ opart,
             const float *__restrict__ ml, int nq_pad, float *__restrict__ mem_val) {
  __shared__ float s_uniform[kChPerCta];  // sum(V)/M of the CTA's channels (out-of-region read)
  const int N = h * w;
  const int o = blockIdx.z;
  const int c0 = blockIdx.y * kChPerCta;
  const int tid = threadIdx.x;

  const int *meta = bank.meta + o * 8;
  const int Z = meta[META_ZEROS_C] + meta[META_ZEROS_T];
  const int M = Z + meta[META_CELLS_C] + meta[META_CELLS_T];
  if (tid < kChPerCta) {
    const float *vs_c = bank.vsum + (size_t)o * RMNET_CV, *vs_t = bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV;
    s_uniform[tid] = (vs_c[c0 + tid] + vs_t[c0 + tid]) * (1.0f / (float)M);
  }
  const int4 qrect = q_rects ? __ldg(reinterpret_cast<const int4 *>(q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
  const int rw = qrect.y - qrect.x + 1;
  const unsigned uN = (unsigned)N;
  float *out_o = mem_val + (unsigned)o * 2u * RMNET_CV * uN;
  __syncthreads();

  // ------------------------------ phase A: fill ------------------------------
  {
    const int lane = tid & 31, cl = tid >> 5;
    const int p0 = blockIdx.x * 128 + lane * VEC;  // first cell of this thread (128 cells per CTA; VEC = 1 covers them in 4 passes)
#pragma unroll 1
    for (int pass = 0; pass < (VEC == 4 ? 1 : 4); ++pass) {
      const int pa = p0 + pass * 32;
      if (pa >= N) break;
      const int ck = c0 + cl * kChPerThread;
      bool in_q[VEC];
      bool any_in = false;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int pos = pa + e;
        const int cy = pos / w, cx = pos - cy * w;
        in_q[e] = cx >= qrect.x && cx <= qrect.y && cy >= qrect.z && cy <= qrect.w;
        any_in |= in_q[e];
      }
      float *out = out_o + (unsigned)ck * uN + pa;
      const float *qv = q_val + (long long)o * q_obj_stride + (unsigned)ck * uN + pa;
      float *oq = out + (unsigned)RMNET_CV * uN;
      if (VEC == 4) {
        // q_val passthrough, channels 512..1023: v4e * att16 (literal multiply keeps the sign of zero like the reference)
        float4 x[kChPerThread];
#pragma unroll
        for (int k = 0; k < kChPerThread; ++k) x[k] = __ldg(reinterpret_cast<const float4 *>(qv + (unsigned)k * uN));
#pragma unroll
        for (int k = 0; k < kChPerThread; ++k) {
          x[k].x *= in_q[0] ? 1.0f : 0.0f; x[k].y *= in_q[1 % VEC] ? 1.0f : 0.0f;
          x[k].z *= in_q[2 % VEC] ? 1.0f : 0.0f; x[k].w *= in_q[3 % VEC] ? 1.0f : 0.0f;
          *reinterpret_cast<float4 *>(oq + (unsigned)k * uN) = x[k];
        }
        if (!any_in) {
#pragma unroll
          for (int k = 0; k < kChPerThread; ++k) {
            const float u = s_uniform[cl * kChPerThread + k];
            *reinterpret_cast<float4 *>(out + (unsigned)k * uN) = make_float4(u, u, u, u);
          }
        } else {
#pragma unroll
          for (int k = 0; k < kChPerThread; ++k) {
            const float u = s_uniform[cl * kChPerThread + k];
#pragma unroll
            for (int e = 0; e < VEC; ++e) if (!in_q[e]) out[(unsigned)k * uN + e] = u;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < kChPerThread; ++k) {
          oq[(unsigned)k * uN] = __ldg(qv + (unsigned)k * uN) * (in_q[0] ? 1.0f : 0.0f);
          if (!in_q[0]) out[(unsigned)k * uN] = s_uniform[cl * kChPerThread + k];
        }
      }
    }
  }

  // ------------------------------ phase B: gather the in-region cells ------------------------------
  const int pos = blockIdx.x * 128 + tid;
  if (pos >= N) return;
  const int cy = pos / w, cx = pos - cy * w;
  if (!(cx >= qrect.x && cx <= qrect.y && cy >= qrect.z && cy <= qrect.w)) return;
  const int n = (cy - qrect.z) * rw + (cx - qrect.x);  // compact query index
  const int half = c0 / (RMNET_CV / 2);
  if (sched_ns) n_splits = __ldg(sched_ns + o);  // KV chunks the persistent tcgen05 kernel used for this object
  const unsigned ml_stride = (unsigned)n_obj * 2u * (unsigned)nq_pad;        // float2 units between splits
  const unsigned op_stride = (unsigned)n_obj * RMNET_CV * (unsigned)nq_pad;  // floats between splits
  const float2 *mlp = reinterpret_cast<const float2 *>(ml) + ((unsigned)o * 2u + half) * (unsigned)nq_pad + n;
  float wgt[READ_MAX_SPLITS];
  float2 st[READ_MAX_SPLITS];
#pragma unroll
  for (int s = 0; s < READ_MAX_SPLITS; ++s) st[s] = (s < n_splits) ? __ldg(mlp + (unsigned)s * ml_stride) : make_float2(-INFINITY, 0.f);
  float m_star = Z > 0 ? 0.f : -INFINITY;
#pragma unroll
  for (int s = 0; s < READ_MAX_SPLITS; ++s) m_star = fmaxf(m_star, st[s].x);
  float L = Z > 0 ? (float)Z * ex2f(-m_star) : 0.f;
#pragma unroll
  for (int s = 0; s < READ_MAX_SPLITS; ++s) {
    wgt[s] = (st[s].x == -INFINITY) ? 0.f : ex2f(st[s].x - m_star);  // a split that saw no cells left its numerators unwritten
    L = fmaf(st[s].y, wgt[s], L);
  }
  const float inv_l = 1.0f / L;
  const float *opb = opart + ((unsigned)o * RMNET_CV + (unsigned)c0) * (unsigned)nq_pad + n;
  float *outp = out_o + (unsigned)c0 * uN + pos;
  // 16 channels x 4 splits = 64 independent loads in flight per round trip
#pragma unroll 1
  for (int g = 0; g < kChPerCta; g += 16) {
    float num[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) num[k] = 0.f;
#pragma unroll
    for (int s0 = 0; s0 < READ_MAX_SPLITS; s0 += 4) {
      if (s0 >= n_splits) break;
      float v[4][16];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float *q = opb + (unsigned)(s0 + u) * op_stride + (unsigned)g * (unsigned)nq_pad;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[u][k] = (wgt[s0 + u] != 0.f) ? __ldg(q + (unsigned)k * (unsigned)nq_pad) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 16; ++k) num[k] = fmaf(v[u][k], wgt[s0 + u], num[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) outp[(unsigned)(g + k) * uN] = num[k] * inv_l;
  }
}

}  // namespace

int launch_merge(const BankView &bank, const float *q_val, long long q_obj_stride, const int *q_rects, int n_obj, int h,
                 int w, int n_splits, bool device_sched, const ReadWorkspace &W, float *mem_val, cudaStream_t st) {
  const int N = h * w;
  const bool vec = N % 4 == 0 && ((uintptr_t)q_val % 16 == 0) && ((uintptr_t)mem_val % 16 == 0) && (q_obj_stride % 4 == 0);
  const int *ns = device_sched ? W.sched : nullptr;
  if (vec) {
    dim3 grid(cdiv(N, 128), RMNET_CV / kChPerCta, n_obj);
    merge_kernel<4><<<grid, kMergeThreads, 0, st>>>(bank, q_val, q_obj_stride, q_rects, h, w, n_obj, n_splits, ns, W.opart, W.ml,
                                                    W.nq_pad, mem_val);
  } else {
    dim3 grid(cdiv(N, 128), RMNET_CV / kChPerCta, n_obj);
    merge_kernel<1><<<grid, kMergeThreads, 0, st>>>(bank, q_val, q_obj_stride, q_rects, h, w, n_obj, n_splits, ns, W.opart, W.ml,
                                                    W.nq_pad, mem_val);
  }
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // namespace rmnet
