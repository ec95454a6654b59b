// memory_read_simt.cu -- CUDA-core (fp32 FFMA) flash-style regional memory read over the packed bank.
//
// Same contract as the tcgen05 kernel (memory_read_umma.cu): one CTA per (query tile, object, Cv half, KV split)
// writes unnormalised partial numerators + (max, sum) row statistics; merge.cu combines the splits, applies the
// analytic correction for the masked (never stored) memory cells and scatters into mem_val.
// It exists as the in-repo cross-check of the tensor-core kernel and for shapes the UMMA path rejects; operands are
// reconstructed as hi+lo (16 mantissa bits for bf16 planes) and multiplied in full fp32.
//
// Math (models/rmnet.py:155-160): t_j = (k_j . q) * log2(e)/sqrt(128);  p_j = 2^(t_j - m);  O = sum_j p_j v_j.
#include "common.cuh"

namespace rmnet {
namespace {

constexpr int QT = 64;    // queries per CTA
constexpr int MT = 64;    // memory cells per KV tile
constexpr int CVH = 256;  // value channels per CTA (one half of Cv)
constexpr int kThreads = 256;
constexpr int PADQ = QT + 1, PADM = MT + 1;

struct SimtSmem {
  float q[RMNET_CK][PADQ];  // [c][query]
  float k[RMNET_CK][PADM];  // [c][cell]
  float v[CVH][PADM];       // [cv][cell]   (reused as the [cv][query] staging tile of the epilogue)
  float p[QT][PADM];        // [query][cell]
};

__global__ void __launch_bounds__(kThreads, 1)
memory_read_simt_kernel(BankView bank, const float *__restrict__ q_key, long long q_obj_stride,
                        const int *__restrict__ q_rects, int h, int w, int fmt, int use_lo, int n_splits,
                        float *__restrict__ opart, float *__restrict__ ml, int nq_pad, int n_obj) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SimtSmem &S = *reinterpret_cast<SimtSmem *>(smem_raw);

  const int N = h * w;
  const int o = blockIdx.y;
  const int half = blockIdx.z & 1, split = blockIdx.z >> 1;
  const int4 qrect = q_rects ? __ldg(reinterpret_cast<const int4 *>(q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
  const int nq = rect_cells(qrect);
  const int q0 = blockIdx.x * QT;
  if (q0 >= nq) return;

  const int *meta = bank.meta + o * 8;
  const int count = meta[META_CELLS_C] + meta[META_CELLS_T];
  int tile_begin, n_it;
  split_range(count, n_splits, split, tile_begin, n_it);
  const int tile_end = tile_begin + n_it;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float scale = 1.4426950408889634f / sqrtf((float)RMNET_CK);

  // ---- Q tile: gather the in-region query cells, [c][query]
  {
    const int qi = threadIdx.x & (QT - 1), cg = threadIdx.x / QT;
    const bool live = q0 + qi < nq;
    const int pos = live ? rect_pos(qrect, q0 + qi, w) : 0;
    const float *qp = q_key + (long long)o * q_obj_stride + pos;
    for (int c = cg; c < RMNET_CK; c += kThreads / QT) S.q[c][qi] = live ? __ldg(qp + (long long)c * N) : 0.f;
  }

  float m_run[4], l_run[4], acc[4][16];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[a][k] = 0.f;
  }

  const uint16_t *khi = bank.khi + (size_t)o * bank.cap * RMNET_CK, *klo = bank.klo + (size_t)o * bank.cap * RMNET_CK;
  const uint16_t *vhi = bank.vhi + ((size_t)o * RMNET_CV + half * CVH) * bank.cap;
  const uint16_t *vlo = bank.vlo + ((size_t)o * RMNET_CV + half * CVH) * bank.cap;

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int m0 = tile * MT;
    __syncthreads();  // previous tile's readers are done with k / v / p (also orders the Q fill on the first pass)
    // ---- K tile: rows [cell][128] -> [c][cell]
    for (int e = threadIdx.x; e < MT * RMNET_CK; e += kThreads) {
      const int cell = e / RMNET_CK, c = e % RMNET_CK;
      float x = 0.f;
      if (m0 + cell < count) {
        const size_t g = (size_t)(m0 + cell) * RMNET_CK + c;
        x = use_lo ? join16(khi[g], klo[g], fmt) : cvt16(khi[g], fmt);
      }
      S.k[c][cell] = x;
    }
    // ---- V tile: rows [cv][cell] -> [cv][cell]
    for (int e = threadIdx.x; e < CVH * MT; e += kThreads) {
      const int cv = e / MT, cell = e % MT;
      float x = 0.f;
      if (m0 + cell < count) {
        const size_t g = (size_t)cv * bank.cap + m0 + cell;
        x = use_lo ? join16(vhi[g], vlo[g], fmt) : cvt16(vhi[g], fmt);
      }
      S.v[cv][cell] = x;
    }
    __syncthreads();

    // ---- S = Q^T K : thread owns queries ty+16a, cells tx+16b
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = 0.f;
#pragma unroll 4
    for (int c = 0; c < RMNET_CK; ++c) {
      float qa[4], kb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qa[a] = S.q[c][ty + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) kb[b] = S.k[c][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s[a][b] = fmaf(qa[a], kb[b], s[a][b]);
    }
    // ---- online softmax (rows are spread over the 16 tx lanes of a half-warp)
    float alpha[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = -INFINITY;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        s[a][b] = (m0 + tx + 16 * b < count) ? s[a][b] * scale : -INFINITY;
        mx = fmaxf(mx, s[a][b]);
      }
#pragma unroll
      for (int d = 8; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      const float m_new = fmaxf(m_run[a], mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      alpha[a] = exp2f(m_run[a] - m_use);  // exp2f(-inf) = 0 on the first tile
      float rs = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float pv = exp2f(s[a][b] - m_use);
        S.p[ty + 16 * a][tx + 16 * b] = pv;
        rs += pv;
      }
#pragma unroll
      for (int d = 8; d > 0; d >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, d);
      l_run[a] = l_run[a] * alpha[a] + rs;
      m_run[a] = m_new;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[a][k] *= alpha[a];
    }
    __syncthreads();
    // ---- O += P V^T : thread owns queries ty+16a, channels tx+16k
#pragma unroll 2
    for (int mm = 0; mm < MT; ++mm) {
      float pa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) pa[a] = S.p[ty + 16 * a][mm];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float vv = S.v[tx + 16 * k][mm];
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][k] = fmaf(pa[a], vv, acc[a][k]);
      }
    }
  }

  // ---- epilogue: stage [cv][query] through smem so the global writes run along queries (coalesced)
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 16; ++k) S.v[tx + 16 * k][ty + 16 * a] = acc[a][k];
  if (tx == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int n = q0 + ty + 16 * a;
      float2 *dst = reinterpret_cast<float2 *>(ml) + (((size_t)split * n_obj + o) * 2 + half) * nq_pad + n;
      *dst = make_float2(m_run[a], l_run[a]);
    }
  }
  __syncthreads();
  float *ob = opart + (((size_t)split * n_obj + o) * RMNET_CV + half * CVH) * nq_pad + q0;
  for (int e = threadIdx.x; e < CVH * QT; e += kThreads) {
    const int cv = e / QT, qi = e % QT;
    ob[(size_t)cv * nq_pad + qi] = S.v[cv][qi];
  }
}

}  // namespace

int launch_memory_read_simt(const BankView &bank, const float *q_key, long long q_obj_stride, const int *q_rects,
                            int n_obj, int h, int w, int fmt, int precision, int n_splits, const ReadWorkspace &W,
                            cudaStream_t st) {
  static_assert(sizeof(SimtSmem) <= 227 * 1024, "smem budget");
  RMNET_CUDA(cudaFuncSetAttribute(memory_read_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(SimtSmem)));
  dim3 grid(cdiv(h * w, QT), n_obj, 2 * n_splits);
  memory_read_simt_kernel<<<grid, kThreads, sizeof(SimtSmem), st>>>(
      bank, q_key, q_obj_stride, q_rects, h, w, fmt, precision != RMNET_PREC_SINGLE ? 1 : 0  /* the FFMA cross-check kernel has no mixed mode: it runs MIXED as strict */, n_splits, W.opart, W.ml,
      W.nq_pad, n_obj);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // namespace rmnet
