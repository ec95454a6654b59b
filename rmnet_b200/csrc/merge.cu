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

// A persistent grid walking a device-built work list (the query rectangles are only known on the device; a grid sized for
// the whole frame per object was mostly CTAs that exit at once, dispatched in two to three waves):
//   fill items   (before the dependency wait: they need only the bank): 128 cells x 32 channels of one object, thread =
//       VEC consecutive cells x 8 channels: the uniform rows of the cells outside the query region (128-bit stores);
//   gather items (after the wait): 32 in-region queries (compact indices) x 32 channels; thread = one query x 8 channels.
//       Two passes over the splits with run-time trip counts -- the reference maximum, then weights, normaliser and the
//       weighted numerators four splits (36 independent loads) at a time -- in 64 registers, so that the whole list is
//       resident at once (8 CTAs per SM); the former single batch of 80 loads per thread needed 144 registers and ran
//       in two to three rounds.
struct MergeList {
  int4 rect[SCHED_MAX_OBJ];
  int fill_pre[SCHED_MAX_OBJ + 1], gather_pre[SCHED_MAX_OBJ + 1];
};

template <int VEC>
__global__ void __launch_bounds__(kMergeThreads, 8)
merge_kernel(BankView bank, const int *__restrict__ q_rects, int h, int w, int n_obj, int o_base, int n_obj_total, int n_splits_arg,
             const int *__restrict__ sched_ns, const float *__restrict__ opart, const float *__restrict__ ml, int nq_pad,
             float *__restrict__ mem_val) {
  DEV_STAMP_MIN(7);
  __shared__ MergeList L;
  __shared__ float s_uniform[kChPerCta];  // sum(V)/M of a fill item's channels (out-of-region read)
  const int N = h * w;
  const int tid = threadIdx.x;
  const unsigned uN = (unsigned)N;
  const int fill_tiles = (N + 127) / 128, ch_groups = RMNET_CV / kChPerCta;
  // ---- work list.  Everything read here comes from further up the chain (rectangles: region kernel; counters and the
  //      plan's split counts: pack kernel -- both complete before the read kernel, our predecessor, could trigger us).
  if (tid < n_obj) {
    const int4 r = q_rects ? __ldg(reinterpret_cast<const int4 *>(q_rects) + o_base + tid) : make_int4(0, w - 1, 0, h - 1);
    const int count = rect_cells(r);
    L.rect[tid] = r;
    L.fill_pre[tid] = count == N ? 0 : fill_tiles * ch_groups;  // (a dense read has nothing to fill)
    L.gather_pre[tid] = ((count + kQueriesPerCta - 1) / kQueriesPerCta) * ch_groups;
  }
  __syncthreads();
  if (tid < 2) {
    int *pre = tid ? L.gather_pre : L.fill_pre;
    int acc = 0;
    for (int o = 0; o < n_obj; ++o) { const int c = pre[o]; pre[o] = acc; acc += c; }
    pre[n_obj] = acc;
  }
  __syncthreads();

  // ------------------------------ fill items ------------------------------
  for (int item = blockIdx.x; item < L.fill_pre[n_obj]; item += gridDim.x) {
    int ol = 0;
    while (L.fill_pre[ol + 1] <= item) ++ol;
    const int q = item - L.fill_pre[ol];
    const int o = o_base + ol;
    const int c0 = (q % ch_groups) * kChPerCta, p_tile = (q / ch_groups) * 128;
    const int4 qrect = L.rect[ol];
    const int *meta = bank.meta + o * 8;
    float *out_o = mem_val + (unsigned)o * 2u * RMNET_CV * uN;
    const int Z = meta[META_ZEROS_C] + meta[META_ZEROS_T];
    const int M = Z + meta[META_CELLS_C] + meta[META_CELLS_T];
    __syncthreads();  // (s_uniform of the previous item)
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
  }
  DEV_STAMP_MAX(10);
  if ((int)blockIdx.x >= L.gather_pre[n_obj]) return;  // (CTA-uniform) nothing to gather for this CTA

  // ------------------------------ gather items ------------------------------
  // Chained launch: the partial results come from the read kernel.  They are read with ld_dep (common.cuh).
  pdl_wait();
  DEV_STAMP_MIN(8);
  for (int item = blockIdx.x; item < L.gather_pre[n_obj]; item += gridDim.x) {
    int ol = 0;
    while (L.gather_pre[ol + 1] <= item) ++ol;
    const int q = item - L.gather_pre[ol];
    const int o = o_base + ol;
    const int c0 = (q % ch_groups) * kChPerCta, tile = q / ch_groups;
    const int4 qrect = L.rect[ol];
    const int count = rect_cells(qrect);
    const int n = tile * kQueriesPerCta + (tid & 31);  // compact query index
    if (n >= count) continue;
    const int cw = c0 + (tid >> 5) * 8;  // first channel of this warp
    const int *meta = bank.meta + o * 8;
    const int Z = meta[META_ZEROS_C] + meta[META_ZEROS_T];
    const int half = c0 / (RMNET_CV / 2);
    const int n_splits = sched_ns ? __ldg(sched_ns + o) : n_splits_arg;  // KV chunks the persistent tcgen05 kernel used for this object
    const unsigned ml_stride = (unsigned)n_obj_total * 2u * (unsigned)nq_pad;        // float2 units between splits
    const unsigned op_stride = (unsigned)n_obj_total * RMNET_CV * (unsigned)nq_pad;  // floats between splits
    const int pos = rect_pos(qrect, n, w);
    const float2 *mlp = reinterpret_cast<const float2 *>(ml) + ((unsigned)o * 2u + half) * (unsigned)nq_pad + n;
    const float *opb = opart + ((unsigned)o * RMNET_CV + (unsigned)cw) * (unsigned)nq_pad + n;
    float *outp = mem_val + (unsigned)o * 2u * RMNET_CV * uN + (unsigned)cw * uN + pos;
    // pass 1: the reference maximum over the splits (and 0 when masked cells exist: their score is exactly 0)
    float m_star = Z > 0 ? 0.f : -INFINITY;
#pragma unroll 4
    for (int s = 0; s < n_splits; ++s) m_star = fmaxf(m_star, ld_dep(mlp + (unsigned)s * ml_stride).x);
    // pass 2: split weights, normaliser, weighted numerators.  A split that saw no cells (m = -inf) left its numerators
    // unwritten: what was loaded for it is replaced by zero, not multiplied by it.
    float Lsum = Z > 0 ? (float)Z * ex2f(-m_star) : 0.f;
    float num[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) num[k] = 0.f;
#pragma unroll 4
    for (int s = 0; s < n_splits; ++s) {
      const float2 st = ld_dep(mlp + (unsigned)s * ml_stride);
      const float *qp = opb + (unsigned)s * op_stride;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ld_dep(qp + (unsigned)k * (unsigned)nq_pad);
      const float wgt = (st.x == -INFINITY) ? 0.f : ex2f(st.x - m_star);
      Lsum = fmaf(st.y, wgt, Lsum);
#pragma unroll
      for (int k = 0; k < 8; ++k) num[k] = fmaf(wgt != 0.f ? v[k] : 0.f, wgt, num[k]);
    }
    const float inv_l = 1.0f / Lsum;
#pragma unroll
    for (int k = 0; k < 8; ++k) outp[(unsigned)k * uN] = num[k] * inv_l;
  }
  DEV_STAMP_MAX(9);
}

}  // namespace

int launch_merge(const BankView &bank, const int *q_rects, int n_obj, int h, int w, int n_splits, bool device_sched,
                 const ReadWorkspace &W, float *mem_val, bool pdl, cudaStream_t st) {
  const int N = h * w;
  const bool vec = N % 4 == 0 && ((uintptr_t)mem_val % 16 == 0);
  const int *ns = device_sched ? W.sched : nullptr;
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0, v = 0;
    n_sms = (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
  }
  // one launch per batch of at most SCHED_MAX_OBJ objects (the work list's tables); persistent grid: 8 CTAs per SM
  for (int o0 = 0; o0 < n_obj; o0 += SCHED_MAX_OBJ) {
    const int nb = n_obj - o0 < SCHED_MAX_OBJ ? n_obj - o0 : SCHED_MAX_OBJ;
    const long long max_items = (long long)nb * (RMNET_CV / kChPerCta) * (cdiv(N, kQueriesPerCta) > cdiv(N, 128) ? cdiv(N, kQueriesPerCta) : cdiv(N, 128));
    const long long cap = 8LL * n_sms;
    dim3 grid((unsigned)(max_items < cap ? max_items : cap));
    if (vec)
      RMNET_CUDA(launch_kernel(merge_kernel<4>, grid, dim3(kMergeThreads), 0, st, pdl, bank, q_rects, h, w, nb, o0, n_obj, n_splits, ns,
                               W.opart, W.ml, W.nq_pad, mem_val));
    else
      RMNET_CUDA(launch_kernel(merge_kernel<1>, grid, dim3(kMergeThreads), 0, st, pdl, bank, q_rects, h, w, nb, o0, n_obj, n_splits, ns,
                               W.opart, W.ml, W.nq_pad, mem_val));
    RMNET_LAUNCH_CHECK();
  }
  return RMNET_OK;
}

}  // namespace rmnet
