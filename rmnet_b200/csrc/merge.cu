// merge.cu -- combine the split-KV partial results, apply the masked-cell correction, scatter into mem_val.
//
// For object o and query cell `pos` of the h x w grid (models/rmnet.py:147-165 with the regional masks of
// :245-248 / :355-358 applied analytically, SURVEY 8a "memory-read regional identity"):
//   in-region query n :  m* = max_s m_s (and 0 if Z > 0);  L = sum_s l_s 2^(m_s - m*) + Z 2^(-m*);
//                        mem[c] = sum_s O_s[c] 2^(m_s - m*) / L
//       Z = number of masked memory cells (score exactly 0, value exactly 0: they only enter the denominator)
//   out-of-region query:  every score is 0  =>  p = 1/M  =>  mem[c] = sum_j V_j[c] / M   (the bank's vsum)
//   (mem_val[o, 512 + c, pos] = q_val[c, pos] * att16(o, pos), :358 + :163, is written by the pack kernel's query role)
// Memory / latency bound.  A thread owns VEC (4) consecutive cells x 8 channels: 128-bit stores along the cells for
// the uniform rows, scalar gathers of the partial numerators only for in-region cells.
#include "common.cuh"

namespace rmnet {
RMNET_DEV_STAMPS(merge)
namespace {

constexpr int kMergeThreads = 128;
constexpr int kChPerCta = 32;
constexpr int kChPerThread = 8;  // 4 warps x 8 channels
constexpr int kQueriesPerCta = 32;  // gather role

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Two CTA roles in one launch, both 128 threads x 32 channels of one object:
//   gather (blockIdx.x < n_gather_tiles): 32 in-region queries (compact indices) x 32 channels; thread = one query x 8
//       channels: statistics of all splits in one batch of loads, then the partial numerators of eight splits x eight
//       channels in one batch (64 independent loads in flight), scattered to the query's cell.
//   fill   (blockIdx.x >= n_gather_tiles): thread = VEC consecutive cells x 8 channels: the uniform rows of the cells
//       outside the query region (128-bit stores).  Needs only the bank, so in a chained launch it runs before the wait.
template <int VEC>
__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(BankView bank, const int *__restrict__ q_rects, int h, int w, int n_obj, int n_splits,
             const int *__restrict__ sched_ns, const float *__restrict__ opart, const float *__restrict__ ml, int nq_pad,
             float *__restrict__ mem_val, int n_gather_tiles) {
  DEV_STAMP_MIN(7);
  const int N = h * w;
  const int o = blockIdx.z;
  const int c0 = blockIdx.y * kChPerCta;
  const int tid = threadIdx.x;
  const int4 qrect = q_rects ? __ldg(reinterpret_cast<const int4 *>(q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
  const unsigned uN = (unsigned)N;
  float *out_o = mem_val + (unsigned)o * 2u * RMNET_CV * uN;
  const int *meta = bank.meta + o * 8;

  if ((int)blockIdx.x >= n_gather_tiles) {
    // ------------------------------ fill role ------------------------------
    const int p_tile = ((int)blockIdx.x - n_gather_tiles) * 128;
    // tiles entirely inside the query region have nothing to fill (rows of the rectangle are usually narrower than a
    // tile, so this only triggers for dense reads)
    if (rect_cells(qrect) == N) return;
    __shared__ float s_uniform[kChPerCta];  // sum(V)/M of the CTA's channels (out-of-region read)
    const int Z = meta[META_ZEROS_C] + meta[META_ZEROS_T];
    const int M = Z + meta[META_CELLS_C] + meta[META_CELLS_T];
    if (tid < kChPerCta) {
      const long long *vs_c = bank.vsum + (size_t)o * RMNET_CV, *vs_t = bank.vsum + ((size_t)bank.n_slots + o) * RMNET_CV;
      s_uniform[tid] = (__ll2float_rn(vs_c[c0 + tid] + vs_t[c0 + tid]) * VSUM_INV_SCALE) * (1.0f / (float)M);
    }
    __syncthreads();
    const int lane = tid & 31, cl = tid >> 5;
    const int p0 = p_tile + lane * VEC;  // first cell of this thread (VEC = 1 covers the 128 cells in 4 passes)
#pragma unroll 1
    for (int pass = 0; pass < (VEC == 4 ? 1 : 4); ++pass) {
      const int pa = p0 + pass * 32;
      if (pa >= N) break;
      const int ck = c0 + cl * kChPerThread;
      bool in_q[VEC];
      bool any_in = false, all_in = true;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int pos = pa + e;
        const int cy = pos / w, cx = pos - cy * w;
        in_q[e] = cx >= qrect.x && cx <= qrect.y && cy >= qrect.z && cy <= qrect.w;
        any_in |= in_q[e];
        all_in &= in_q[e];
      }
      if (all_in) continue;
      float *out = out_o + (unsigned)ck * uN + pa;
      if (VEC == 4 && !any_in) {
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
    }
    DEV_STAMP_MAX(10);
    return;
  }

  // ------------------------------ gather role ------------------------------
  // CTA = 32 compact queries x 32 channels: warp k owns channels [8k, 8k+8) of the same 32 queries.
  // Before the dependency wait: everything that comes from further up the chain (rectangles: region kernel; counters
  // and the plan's split counts: pack kernel -- both complete before the read kernel, our predecessor, could trigger us).
  const int count = rect_cells(qrect);
  if ((int)blockIdx.x * kQueriesPerCta >= count) return;  // CTA-uniform
  const int n = (int)blockIdx.x * kQueriesPerCta + (tid & 31);  // compact query index
  const int cw = c0 + (tid >> 5) * 8;  // first channel of this warp
  const int Z = meta[META_ZEROS_C] + meta[META_ZEROS_T];
  const int half = c0 / (RMNET_CV / 2);
  if (sched_ns) n_splits = __ldg(sched_ns + o);  // KV chunks the persistent tcgen05 kernel used for this object
  const unsigned ml_stride = (unsigned)n_obj * 2u * (unsigned)nq_pad;        // float2 units between splits
  const unsigned op_stride = (unsigned)n_obj * RMNET_CV * (unsigned)nq_pad;  // floats between splits
  const bool live = n < count;
  const int nn = live ? n : count - 1;  // dead lanes of the last tile shadow a live query (no divergent exit before the wait)
  const int pos = rect_pos(qrect, nn, w);
  const float2 *mlp = reinterpret_cast<const float2 *>(ml) + ((unsigned)o * 2u + half) * (unsigned)nq_pad + nn;
  const float *opb = opart + ((unsigned)o * RMNET_CV + (unsigned)cw) * (unsigned)nq_pad + nn;
  float *outp = out_o + (unsigned)cw * uN + pos;
  // Chained launch: the partial results below come from the read kernel.
  pdl_wait();
  DEV_STAMP_MIN(8);
  // ONE round trip: the statistics of all splits and the partial numerators of the first eight splits x eight channels
  // are requested together (72-80 independent loads per thread); a split that saw no cells left its numerators unwritten,
  // so what was loaded for it is replaced by zero, not multiplied by it.
  float2 st[READ_MAX_SPLITS];
#pragma unroll
  for (int s = 0; s < READ_MAX_SPLITS; ++s) st[s] = (s < n_splits) ? ld_dep(mlp + (unsigned)s * ml_stride) : make_float2(-INFINITY, 0.f);
  float v[8][8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float *q = opb + (unsigned)u * op_stride;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[u][k] = (u < n_splits) ? ld_dep(q + (unsigned)k * (unsigned)nq_pad) : 0.f;
  }
  float wgt[READ_MAX_SPLITS];
  {
    float m_star = Z > 0 ? 0.f : -INFINITY;
#pragma unroll
    for (int s = 0; s < READ_MAX_SPLITS; ++s) m_star = fmaxf(m_star, st[s].x);
    float L = Z > 0 ? (float)Z * ex2f(-m_star) : 0.f;
#pragma unroll
    for (int s = 0; s < READ_MAX_SPLITS; ++s) {
      wgt[s] = (st[s].x == -INFINITY) ? 0.f : ex2f(st[s].x - m_star);
      L = fmaf(st[s].y, wgt[s], L);
    }
    const float inv_l = 1.0f / L;
#pragma unroll
    for (int s = 0; s < READ_MAX_SPLITS; ++s) wgt[s] *= inv_l;  // fold the normalisation into the split weights
  }
  float num[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) num[k] = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int k = 0; k < 8; ++k) num[k] = fmaf(wgt[u] != 0.f ? v[u][k] : 0.f, wgt[u], num[k]);
  if (n_splits > 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float *q = opb + (unsigned)(8 + u) * op_stride;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[u][k] = (wgt[8 + u] != 0.f) ? ld_dep(q + (unsigned)k * (unsigned)nq_pad) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) num[k] = fmaf(v[u][k], wgt[8 + u], num[k]);
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < 8; ++k) outp[(unsigned)k * uN] = num[k];
  }
  DEV_STAMP_MAX(9);
}

}  // namespace

// fill_uniform = false: the uniform rows have been written by the read kernel (memory_read_umma.cu), gather role only.
int launch_merge(const BankView &bank, const int *q_rects, int n_obj, int h, int w, int n_splits, bool device_sched,
                 const ReadWorkspace &W, float *mem_val, bool fill_uniform, bool pdl, cudaStream_t st) {
  const int N = h * w;
  const bool vec = N % 4 == 0 && ((uintptr_t)mem_val % 16 == 0);
  const int *ns = device_sched ? W.sched : nullptr;
  const int n_gather_tiles = W.nq_pad / kQueriesPerCta;
  if (vec) {
    dim3 grid(n_gather_tiles + (fill_uniform ? cdiv(N, 128) : 0), RMNET_CV / kChPerCta, n_obj);
    RMNET_CUDA(launch_kernel(merge_kernel<4>, grid, dim3(kMergeThreads), 0, st, pdl, bank, q_rects, h, w, n_obj, n_splits, ns,
                             W.opart, W.ml, W.nq_pad, mem_val, n_gather_tiles));
  } else {
    dim3 grid(n_gather_tiles + (fill_uniform ? cdiv(N, 128) : 0), RMNET_CV / kChPerCta, n_obj);
    RMNET_CUDA(launch_kernel(merge_kernel<1>, grid, dim3(kMergeThreads), 0, st, pdl, bank, q_rects, h, w, n_obj, n_splits, ns,
                             W.opart, W.ml, W.nq_pad, mem_val, n_gather_tiles));
  }
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // namespace rmnet
