// attention_probs.cu -- the second output of MemoryReader.forward, on request only.
//
//   p = F.softmax(torch.bmm(mi, qi) / math.sqrt(128), dim=1)      [n, T*h*w, h*w]      (models/rmnet.py:155-157)
//
// The fused read never materialises p (210 MB per object at 480p, T = 20; the only caller discards it, :361).  For
// callers that do want the reference's `viz` tensor this kernel recomputes the scores in fp32 FFMA from the RAW fp32
// keys (not the 16-bit planes) and writes the normalised probabilities: thread = one query, two passes over the
// memory cells (online max / sum, then exp(s - max) / sum).  Write-bound (4*M*N bytes per object); not on the per-frame path.
#include "common.cuh"

namespace rmnet {
namespace {

constexpr int kThreads = 128;
constexpr int kCells = 32;  // memory cells per shared-memory chunk

__global__ void __launch_bounds__(kThreads)
attention_probs_kernel(const float *__restrict__ m_key, const float *__restrict__ q_key, int M, int N, float *__restrict__ p) {
  __shared__ __align__(16) float s_k[kCells][RMNET_CK];
  const int o = blockIdx.y;
  const int q = blockIdx.x * kThreads + threadIdx.x;
  const bool live = q < N;
  const float *mk = m_key + (size_t)o * RMNET_CK * M;
  float qv[RMNET_CK];
#pragma unroll
  for (int c = 0; c < RMNET_CK; ++c) qv[c] = live ? __ldg(q_key + ((size_t)o * RMNET_CK + c) * N + q) : 0.f;
  const float denom = sqrtf((float)RMNET_CK);  // `p / math.sqrt(self.keydim)`-style scaling (:156): a float32 division
  float mx = -INFINITY, sum = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    for (int m0 = 0; m0 < M; m0 += kCells) {
      __syncthreads();
      // thread t stages channel t of the chunk's cells (its 128 B segment stays in L1 across the loop)
      for (int i = 0; i < kCells; ++i) s_k[i][threadIdx.x] = (m0 + i < M) ? __ldg(mk + (size_t)threadIdx.x * M + m0 + i) : 0.f;
      __syncthreads();
      const int cnt = min(kCells, M - m0);
      for (int i = 0; i < cnt; ++i) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < RMNET_CK; c += 4) {
          const float4 kv = *reinterpret_cast<const float4 *>(&s_k[i][c]);  // broadcast read
          s0 = fmaf(qv[c], kv.x, s0); s1 = fmaf(qv[c + 1], kv.y, s1);
          s2 = fmaf(qv[c + 2], kv.z, s2); s3 = fmaf(qv[c + 3], kv.w, s3);
        }
        const float s = __fdiv_rn((s0 + s1) + (s2 + s3), denom);
        if (pass == 0) {
          const float nm = fmaxf(mx, s);
          sum = sum * expf(mx - nm) + expf(s - nm);   // expf(-inf) = 0 on the first cell
          mx = nm;
        } else if (live) {
          p[((size_t)o * M + m0 + i) * N + q] = __fdiv_rn(expf(s - mx), sum);
        }
      }
    }
  }
}

}  // namespace

int launch_attention_probs(const float *m_key, const float *q_key, int n, int M, int N, float *p, cudaStream_t st) {
  dim3 grid(cdiv(N, kThreads), n);
  attention_probs_kernel<<<grid, kThreads, 0, st>>>(m_key, q_key, M, N, p);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // namespace rmnet
