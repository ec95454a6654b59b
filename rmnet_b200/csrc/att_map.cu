// att_map.cu -- regional attention-map generator, flow warp, and the fused warp+bbox kernel.
//
// Replaces (see include/rmnet_b200.h for the ABI):
//   extensions/reg_att_map_generator/reg_att_map_generator.cu:15-123   (1 block x 512 threads, 5 global
//       atomics per foreground pixel)  ->  multi-CTA vectorised scan, REDUX warp reductions, one atomic
//       set per CTA, last-CTA finalise, self-cleaning workspace.
//   models/rmnet.py:252-278 RMNet.warp (~15 ATen kernels + 2 grid_samples + a CPU-built grid)
//       ->  one elementwise kernel; and fused with the generator (get_att_map, :280-287) so the warped
//       mask never touches HBM.
// All kernels are HBM-bound scans: coalesced (128-bit where alignment allows) loads, grids sized to
// several CTAs per SM on 148 SMs.
#include "common.cuh"

namespace rmnet {
RMNET_DEV_STAMPS(att_map)
namespace {

constexpr int kThreads = 256;
constexpr int kWsIntsPerChannel = 8;  // cnt, inv_xmin, xmax, inv_ymin, ymax, ticket, -, -
// The fused kernel's ~600 CTAs publish into kWsCopies copies of the accumulator array (CTA c into copy c % kWsCopies):
// atomics on one 32-byte sector serialise in L2, and with one copy the 5 words of a channel took ~3 000 of them -- 4 to
// 6 us between the last CTA leaving the pixel loop and the last ticket.  The finalising CTA folds the copies.
constexpr int kWsCopies = 16;

// ---- per-thread bbox accumulator -------------------------------------------------------------
struct BoxAcc {
  int cnt, xmin, xmax, ymin, ymax;
  __device__ __forceinline__ void init() { cnt = 0; xmin = 32767; xmax = 0; ymin = 32767; ymax = 0; }
  __device__ __forceinline__ void hit(int x, int y) {
    ++cnt;
    xmin = min(xmin, x); xmax = max(xmax, x);
    ymin = min(ymin, y); ymax = max(ymax, y);
  }
  __device__ __forceinline__ void warp_reduce() {
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    xmin = __reduce_min_sync(0xffffffffu, xmin);
    xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin);
    ymax = __reduce_max_sync(0xffffffffu, ymax);
  }
};

// Accumulate a CTA-level result into the global (zero-identity) workspace of one channel.
// mins are stored as max(32767 - v) so that an all-zero workspace is the identity for every field.
__device__ __forceinline__ void ws_accumulate(int *ws, const BoxAcc &a) {
  if (a.cnt > 0) {
    atomicAdd(ws + 0, a.cnt);
    atomicMax(ws + 1, 32767 - a.xmin);
    atomicMax(ws + 2, a.xmax);
    atomicMax(ws + 3, 32767 - a.ymin);
    atomicMax(ws + 4, a.ymax);
  }
}

// How the raw scan result of one channel becomes the outputs.  The scan runs on the UNPADDED mask; `off_*` shifts
// it into the frame the reference computes the box in (the zero-padded frame on the memorise path,
// models/rmnet.py:212+:244; the raw frame on the segment path, :431), Hf x Wf is that frame's size.
// rects (nullable): the /16 cell rectangle of the box, i.e. pad (:307) + F.interpolate(1/16) (:245, :356) in closed form.
struct BoxFinalize {
  int Hf, Wf, off_x, off_y;  // frame of the bbox
  int n_pts_threshold, loose, force_full;
  int k_scan;                // channels [1, k_scan) are scanned; channels >= k_scan are reported as absent objects
  int *rects;                // [B,K,4] or nullptr
  int rect_pad_l, rect_pad_t, cell_h, cell_w;
};

__device__ __forceinline__ void write_rect(const BoxFinalize &f, long long ch, const int4 bb, bool empty) {
  if (!f.rects) return;
  int4 r;
  r.x = max(0, (bb.x + f.rect_pad_l + 15) >> 4);
  r.y = min(f.cell_w - 1, (bb.y + f.rect_pad_l) >> 4);
  r.z = max(0, (bb.z + f.rect_pad_t + 15) >> 4);
  r.w = min(f.cell_h - 1, (bb.w + f.rect_pad_t) >> 4);
  if (empty || r.x > r.y || r.z > r.w) r = make_int4(0, -1, 0, -1);
  reinterpret_cast<int4 *>(f.rects)[ch] = r;
}

// reg_att_map_generator.cu:55-77 -- loosen / clamp, or full frame when too few points.
// Reads AND clears the workspace accumulators (self-cleaning).
// From the folded accumulators of one channel (mins stored inverted) to the outputs.
__device__ __forceinline__ void finalize_values(int cnt, int a1, int a2, int a3, int a4, int *bboxes, long long ch,
                                                const BoxFinalize &f) {
  int xmin = 32767 - a1 + f.off_x;
  int xmax = a2 + f.off_x;
  int ymin = 32767 - a3 + f.off_y;
  int ymax = a4 + f.off_y;
  int4 r;
  if (cnt < f.n_pts_threshold || f.force_full) {
    r = make_int4(0, f.Wf - 1, 0, f.Hf - 1);
  } else {
    r.x = xmin <= f.loose ? 0 : xmin - f.loose;
    r.y = xmax + f.loose >= f.Wf ? f.Wf - 1 : xmax + f.loose;
    r.z = ymin <= f.loose ? 0 : ymin - f.loose;
    r.w = ymax + f.loose >= f.Hf ? f.Hf - 1 : ymax + f.loose;
  }
  reinterpret_cast<int4 *>(bboxes)[ch] = r;  // bboxes are [.,4] i32, 16 B aligned per entry
  write_rect(f, ch, r, false);
}
// Reads AND clears the workspace accumulators of one channel (self-cleaning), single-copy form.
__device__ __forceinline__ void finalize_channel(int *ws, int *bboxes, long long ch, const BoxFinalize &f) {
  const int cnt = atomicExch(ws + 0, 0), a1 = atomicExch(ws + 1, 0), a2 = atomicExch(ws + 2, 0), a3 = atomicExch(ws + 3, 0),
            a4 = atomicExch(ws + 4, 0);
  finalize_values(cnt, a1, a2, a3, a4, bboxes, ch, f);
}
__device__ __forceinline__ void finalize_channel0(int *bboxes, long long ch, const BoxFinalize &f) {
  reinterpret_cast<int4 *>(bboxes)[ch] = make_int4(0, 0, 0, 0);  // channel 0 keeps the zero fill (.cu:104)
  write_rect(f, ch, make_int4(0, 0, 0, 0), true);               // its att_map is all zero
}

// ---- plain generator: grid (chunks, K-1, B) ----------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
bbox_scan_kernel(const float *__restrict__ mask, int K, int H, int W, float thr, BoxFinalize fin,
                 int elems_per_cta, int *__restrict__ bboxes, int *__restrict__ ws) {
  const int b = blockIdx.z, i = blockIdx.y + 1;
  const long long n_pixels = (long long)H * W;
  const float *plane = mask + ((long long)b * K + i) * n_pixels;
  int *ws_ch = ws + ((long long)b * 2 * (K + 1) + i) * kWsIntsPerChannel;
  const long long begin = (long long)blockIdx.x * elems_per_cta;
  const long long end = min(begin + (long long)elems_per_cta, n_pixels);

  BoxAcc acc;
  acc.init();
  if (VEC == 4) {
    // 128-bit loads, 4 independent requests in flight per thread
    const float4 *p4 = reinterpret_cast<const float4 *>(plane);
    for (long long j0 = begin + (long long)threadIdx.x * 4; j0 < end; j0 += (long long)kThreads * 4 * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        long long j = j0 + (long long)u * kThreads * 4;
        v[u] = (j < end) ? __ldg(p4 + (j >> 2)) : make_float4(-1.f, -1.f, -1.f, -1.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        long long j = j0 + (long long)u * kThreads * 4;
        if (j >= end) continue;
        // fast path: most 16-byte groups hold no foreground pixel at all (4 compares, no index arithmetic)
        if (!((v[u].x >= thr) | (v[u].y >= thr) | (v[u].z >= thr) | (v[u].w >= thr))) continue;
        int y = (int)(j / W), x = (int)(j - (long long)y * W);
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (e[k] >= thr) acc.hit(x, y);  // NaN compares false, as in the reference (:42)
          if (++x == W) { x = 0; ++y; }
        }
      }
    }
  } else {
    for (long long j = begin + threadIdx.x; j < end; j += kThreads) {
      float v = __ldg(plane + j);
      if (v >= thr) { int y = (int)(j / W); acc.hit((int)(j - (long long)y * W), y); }
    }
  }

  __shared__ int s_red[kThreads / 32][5];
  __shared__ bool s_last;
  acc.warp_reduce();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_red[warp][0] = acc.cnt; s_red[warp][1] = acc.xmin; s_red[warp][2] = acc.xmax; s_red[warp][3] = acc.ymin; s_red[warp][4] = acc.ymax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    BoxAcc t;
    t.init();
    for (int wdx = 0; wdx < kThreads / 32; ++wdx) {
      t.cnt += s_red[wdx][0];
      t.xmin = min(t.xmin, s_red[wdx][1]); t.xmax = max(t.xmax, s_red[wdx][2]);
      t.ymin = min(t.ymin, s_red[wdx][3]); t.ymax = max(t.ymax, s_red[wdx][4]);
    }
    ws_accumulate(ws_ch, t);
    __threadfence();
    int ticket = atomicAdd(ws_ch + 5, 1);
    s_last = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    finalize_channel(ws_ch, bboxes, (long long)b * K + i, fin);
    atomicExch(ws_ch + 5, 0);
    if (i == 1) {
      finalize_channel0(bboxes, (long long)b * K, fin);
      for (int u = fin.k_scan; u < K; ++u)  // unscanned channels: their (zero) accumulators give the absent-object box
        finalize_channel(ws + ((long long)b * 2 * (K + 1) + u) * kWsIntsPerChannel, bboxes, (long long)b * K + u, fin);
    }
  }
}

// ---- att_map fill from final bboxes: att[b,i,y,x] = 1 inside the box (reg_att_map_generator.cu:81-92) ----
template <int VEC>
__global__ void __launch_bounds__(kThreads)
att_fill_kernel(const int *__restrict__ bboxes, int K, int H, int W, long long total, float *__restrict__ att) {
  const long long n_pixels = (long long)H * W;
  for (long long j = ((long long)blockIdx.x * kThreads + threadIdx.x) * VEC; j < total;
       j += (long long)gridDim.x * kThreads * VEC) {
    long long ch = j / n_pixels;  // = b*K + i
    long long r = j - ch * n_pixels;
    int y = (int)(r / W), x = (int)(r - (long long)y * W);
    int i = (int)(ch % K);
    const int4 bb = __ldg(reinterpret_cast<const int4 *>(bboxes) + ch);
    float e[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      e[k] = (i != 0 && x >= bb.x && x <= bb.y && y >= bb.z && y <= bb.w) ? 1.0f : 0.0f;
      if (++x == W) { x = 0; ++y; }
    }
    if (VEC == 4) *reinterpret_cast<float4 *>(att + j) = make_float4(e[0], e[1], e[2], e[3]);
    else att[j] = e[0];
  }
}

// ---- flow warp sample point (models/rmnet.py:263-273 on torch's CUDA backend) -------------------
struct Tap {
  int x0, y0;           // north-west corner (may be out of bounds)
  float nw, ne, sw, se; // bilinear weights
  float valid;          // validity mask in {0,1}  (:272-275)
};
template <int ORDER>  // 0 = cuDNN tap order (ne, nw, sw, se), 1 = ATen tap order (nw, ne, sw, se)
__device__ __forceinline__ Tap make_tap(int x, int y, float fx, float fy, int H, int W, float inv_w, float inv_h) {
  Tap t;
  // vgrid = grid + flow (:263); 2.0*v / max(W-1,1) - 1.0 (:265-266): ATen's CUDA div-by-scalar multiplies by the
  // host-computed reciprocal (BinaryDivTrueKernel.cu); each op is a separately rounded fp32 kernel.
  float gx = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, __fadd_rn((float)x, fx)), inv_w), 1.0f);
  float gy = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, __fadd_rn((float)y, fy)), inv_h), 1.0f);
  // grid_sampler_unnormalize(align_corners=True): ((coord + 1) / 2) * (size - 1)   (GridSampler.cuh:23-31)
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
  float flx = floorf(ix), fly = floorf(iy);
  // clamp before the float->int cast (only matters for non-finite / absurd flows; keeps taps out of bounds)
  t.x0 = (flx > -2.0f && flx < (float)W + 1.0f) ? (int)flx : -2;
  t.y0 = (fly > -2.0f && fly < (float)H + 1.0f) ? (int)fly : -2;
  const float wx0 = __fsub_rn(__fadd_rn(flx, 1.0f), ix), wy0 = __fsub_rn(__fadd_rn(fly, 1.0f), iy);
  // east / south weights: ATen uses x - floor(x); cuDNN uses 1 - (west / north weight), which differs by an ulp when
  // that weight is inexact (floor == 0 or -1).  Established empirically against both samplers on a B200.
  const float wx1 = ORDER == 1 ? __fsub_rn(ix, flx) : __fsub_rn(1.0f, wx0);
  const float wy1 = ORDER == 1 ? __fsub_rn(iy, fly) : __fsub_rn(1.0f, wy0);
  t.nw = __fmul_rn(wx0, wy0); t.ne = __fmul_rn(wx1, wy0);
  t.sw = __fmul_rn(wx0, wy1); t.se = __fmul_rn(wx1, wy1);
  const bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  float ones = 0.0f;  // grid_sample of the ones tensor: acc = fma(1, w, acc) in the sampler's tap order
  if (ORDER == 1) {
    if (xin0 && yin0) ones = __fmaf_rn(1.0f, t.nw, ones);
    if (xin1 && yin0) ones = __fmaf_rn(1.0f, t.ne, ones);
  } else {
    if (xin1 && yin0) ones = __fmaf_rn(1.0f, t.ne, ones);
    if (xin0 && yin0) ones = __fmaf_rn(1.0f, t.nw, ones);
  }
  if (xin0 && yin1) ones = __fmaf_rn(1.0f, t.sw, ones);
  if (xin1 && yin1) ones = __fmaf_rn(1.0f, t.se, ones);
  float m = ones;
  if (m < 0.9999f) m = 0.0f;  // :274
  if (m > 0.0f) m = 1.0f;     // :275
  t.valid = m;
  return t;
}
template <int ORDER>
__device__ __forceinline__ float sample_tap(const float *__restrict__ plane, const Tap &t, int H, int W) {
  const bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  const float *r0 = plane + (long long)t.y0 * W + t.x0;
  float acc = 0.0f;  // ATen GridSampler.cu: out_acc += v * w under nvcc -fmad=true; cuDNN: same chain, ne first
  if (ORDER == 1) {
    if (xin0 && yin0) acc = __fmaf_rn(__ldg(r0), t.nw, acc);
    if (xin1 && yin0) acc = __fmaf_rn(__ldg(r0 + 1), t.ne, acc);
  } else {
    if (xin1 && yin0) acc = __fmaf_rn(__ldg(r0 + 1), t.ne, acc);
    if (xin0 && yin0) acc = __fmaf_rn(__ldg(r0), t.nw, acc);
  }
  if (xin0 && yin1) acc = __fmaf_rn(__ldg(r0 + W), t.sw, acc);
  if (xin1 && yin1) acc = __fmaf_rn(__ldg(r0 + W + 1), t.se, acc);
  return __fmul_rn(acc, t.valid);  // img1 * mask (:277)
}

// ---- literal RMNet.warp: one thread per pixel, loop over channels --------------------------------
template <int ORDER>
__global__ void __launch_bounds__(kThreads)
warp_kernel(const float *__restrict__ img0, const float *__restrict__ flow, int C, int H, int W, float inv_w,
            float inv_h, float *__restrict__ img1, float *__restrict__ valid) {
  const int b = blockIdx.y;
  const long long n_pixels = (long long)H * W;
  const long long j = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (j >= n_pixels) return;
  const int y = (int)(j / W), x = (int)(j - (long long)y * W);
  const float *fl = flow + (long long)b * 2 * n_pixels;
  const Tap t = make_tap<ORDER>(x, y, __ldg(fl + j), __ldg(fl + n_pixels + j), H, W, inv_w, inv_h);
  for (int c = 0; c < C; ++c) {
    const long long off = ((long long)b * C + c) * n_pixels;
    img1[off + j] = sample_tap<ORDER>(img0 + off, t, H, W);
    if (valid) valid[off + j] = t.valid;
  }
}

// ---- fused warp + threshold + bbox: grid (pixel tiles, B), one pixel per thread ------------------------
// One pass over prev_mask produces the box of the flow-warped mask (segment side, models/rmnet.py:431) and, when
// DIRECT, also the box of the mask itself (memorise side, :244) -- both read the same est_masks[t-1].
// Channels are processed five at a time with all their loads (4 taps + 1 direct) issued before use: the gathers
// follow the flow and are the latency bottleneck.  A warp covers 32 consecutive pixels; when they lie in one row
// the per-channel box of the warp comes straight from the ballot (popc / ffs / clz), else from REDUX.
constexpr int kChunkCh = 5;

__device__ __forceinline__ void warp_box_to_smem(int *s_acc, unsigned ballot, bool one_row, int x_lane0, int y_lane0,
                                                 bool hit, int x, int y, int lane) {
  if (ballot == 0u) return;
  int cnt, xmin, xmax, ymin, ymax;
  if (one_row) {
    cnt = __popc(ballot);
    xmin = x_lane0 + (__ffs(ballot) - 1);
    xmax = x_lane0 + (31 - __clz(ballot));
    ymin = ymax = y_lane0;
  } else {
    cnt = __popc(ballot);
    xmin = __reduce_min_sync(0xffffffffu, hit ? x : 32767);
    xmax = __reduce_max_sync(0xffffffffu, hit ? x : 0);
    ymin = __reduce_min_sync(0xffffffffu, hit ? y : 32767);
    ymax = __reduce_max_sync(0xffffffffu, hit ? y : 0);
  }
  if (lane == 0) {
    atomicAdd(s_acc + 0, cnt);
    atomicMax(s_acc + 1, 32767 - xmin);
    atomicMax(s_acc + 2, xmax);
    atomicMax(s_acc + 3, 32767 - ymin);
    atomicMax(s_acc + 4, ymax);
  }
}

template <int ORDER, bool DIRECT>
__global__ void __launch_bounds__(kThreads, 4)  // 4 CTAs per SM = the one-wave grid of launch_frame_boxes (the warp-only variant took 80 registers uncapped)
frame_boxes_kernel(const float *__restrict__ prev_mask, const float *__restrict__ flow, int K, int H, int W,
                   float inv_w, float inv_h, float thr, BoxFinalize fin_warp, BoxFinalize fin_direct,
                   int *__restrict__ bboxes_warp, int *__restrict__ bboxes_direct, int *__restrict__ ws,
                   float *__restrict__ clear, int n_clear) {
  extern __shared__ int s_acc[];  // [2][K][5] CTA accumulators (zero identity, mins inverted): warped, direct
  __shared__ bool s_last;
  pdl_trigger();  // head of the frame-step chain: the successor (bank pack) may start launching; it waits for this grid
  DEV_STAMP_MIN(0);
  const int b = blockIdx.y;
  // side duty for rmnet_frame_step: zero the bank's temporary-frame value sums (replaces a memset node in the chain)
  if (clear && blockIdx.x == 0 && b == 0) for (int k = threadIdx.x; k < n_clear; k += kThreads) clear[k] = 0.f;
  const long long n_pixels = (long long)H * W;
  const float *fl = flow + (long long)b * 2 * n_pixels;
  for (int k = threadIdx.x; k < 2 * K * 5; k += kThreads) s_acc[k] = 0;
  __syncthreads();

  // Persistent CTAs (one wave, launch_frame_boxes sizes the grid): each walks pixel tiles of kThreads consecutive pixels;
  // the flow of the NEXT tile is loaded before the gathers of the current one are consumed.
  const long long n_tiles = (n_pixels + kThreads - 1) / kThreads;
  const int lane = threadIdx.x & 31;
  const int Ks = fin_warp.k_scan;  // channels >= Ks are known to be empty (absent objects): not read at all
  long long tile = blockIdx.x;
  float fx_n = 0.f, fy_n = 0.f;
  if (tile < n_tiles) {
    const long long jn = min(tile * kThreads + threadIdx.x, n_pixels - 1);
    fx_n = __ldg(fl + jn);
    fy_n = __ldg(fl + n_pixels + jn);
  }
  for (; tile < n_tiles; tile += gridDim.x) {
    // H, W <= 32767 (check_common), so pixel indices fit 32 bits: a 64-bit division here was 5 % of the kernel's instructions
    const unsigned j = (unsigned)tile * kThreads + threadIdx.x;
    const bool live = j < (unsigned)n_pixels;
    const unsigned jj = live ? j : (unsigned)n_pixels - 1u;
    const float fx = fx_n, fy = fy_n;
    if (tile + gridDim.x < n_tiles) {
      const long long jn = min((tile + gridDim.x) * kThreads + threadIdx.x, n_pixels - 1);
      fx_n = __ldg(fl + jn);
      fy_n = __ldg(fl + n_pixels + jn);
    }
    const int y = (int)(jj / (unsigned)W), x = (int)(jj - (unsigned)y * (unsigned)W);
    const int x_lane0 = __shfl_sync(0xffffffffu, x, 0), y_lane0 = __shfl_sync(0xffffffffu, y, 0);
    const bool one_row = __shfl_sync(0xffffffffu, y, 31) == y_lane0 && __all_sync(0xffffffffu, live);
    const Tap t = make_tap<ORDER>(x, y, fx, fy, H, W, inv_w, inv_h);
    const bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    const bool p_nw = xin0 && yin0, p_ne = xin1 && yin0, p_sw = xin0 && yin1, p_se = xin1 && yin1;
    // 32-bit in-plane offsets; an out-of-bounds tap reads the thread's own pixel instead and is then replaced by 0
    // (selected, not multiplied: a NaN there must not leak), so the loop body has no address predication.
    const int o_self = (int)jj;
    const int o_nw = t.y0 * W + t.x0;
    const int a_nw = p_nw ? o_nw : o_self, a_ne = p_ne ? o_nw + 1 : o_self;
    const int a_sw = p_sw ? o_nw + W : o_self, a_se = p_se ? o_nw + W + 1 : o_self;
    const float *plane = prev_mask + ((long long)b * K + 1) * n_pixels;  // channel 1; advanced by n_pixels per channel

    for (int i0 = 1; i0 < Ks; i0 += kChunkCh) {
      const int nch = min(kChunkCh, Ks - i0);  // warp-uniform
      float v_nw[kChunkCh], v_ne[kChunkCh], v_sw[kChunkCh], v_se[kChunkCh], v_d[kChunkCh];
      {
        const float *pl = plane;
#pragma unroll
        for (int u = 0; u < kChunkCh; ++u) {
          if (u < nch) {
            v_nw[u] = __ldg(pl + a_nw);
            v_ne[u] = __ldg(pl + a_ne);
            v_sw[u] = __ldg(pl + a_sw);
            v_se[u] = __ldg(pl + a_se);
            if (DIRECT) v_d[u] = __ldg(pl + o_self);
            pl += n_pixels;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kChunkCh; ++u) {
        if (u >= nch) break;  // warp-uniform
        const int i = i0 + u;
        // same FMA chain as sample_tap (an out-of-bounds tap contributes fma(0, w, acc) = acc exactly)
        const float x_nw = p_nw ? v_nw[u] : 0.f, x_ne = p_ne ? v_ne[u] : 0.f;
        const float x_sw = p_sw ? v_sw[u] : 0.f, x_se = p_se ? v_se[u] : 0.f;
        float acc;
        if (ORDER == 1) acc = __fmaf_rn(x_ne, t.ne, __fmul_rn(x_nw, t.nw));
        else acc = __fmaf_rn(x_nw, t.nw, __fmul_rn(x_ne, t.ne));
        acc = __fmaf_rn(x_se, t.se, __fmaf_rn(x_sw, t.sw, acc));
        const bool hit_w = live && (__fmul_rn(acc, t.valid) >= thr);
        warp_box_to_smem(s_acc + i * 5, __ballot_sync(0xffffffffu, hit_w), one_row, x_lane0, y_lane0, hit_w, x, y, lane);
        if (DIRECT) {
          const bool hit_d = live && (v_d[u] >= thr);
          warp_box_to_smem(s_acc + (K + i) * 5, __ballot_sync(0xffffffffu, hit_d), one_row, x_lane0, y_lane0, hit_d, x, y, lane);
        }
      }
      plane += (long long)kChunkCh * n_pixels;
    }
  }
  __syncthreads();
  DEV_STAMP_MIN(37); DEV_STAMP_MAX(36);
  // CTA -> global workspace: set 0 (warped) at ws_b, set 1 (direct) right behind it.  Only the threads that publish
  // accumulators pay for the fence (a device-wide fence by all 256 threads was 13 % of this kernel's stall samples).
  int *ws_b = ws + (long long)b * 2 * (K + 1) * kWsIntsPerChannel;
  const long long copy_stride = (long long)gridDim.y * 2 * (K + 1) * kWsIntsPerChannel;  // ints between two copies of the array
  bool published = false;
  for (int e = threadIdx.x; e < (DIRECT ? 2 : 1) * K; e += kThreads) {
    const int set = e / K, i = e - set * K;
    if (i >= 1 && s_acc[e * 5] > 0) {
      int *w = ws_b + (blockIdx.x % kWsCopies) * copy_stride + (set * (K + 1) + i) * kWsIntsPerChannel;
      atomicAdd(w + 0, s_acc[e * 5 + 0]);
      atomicMax(w + 1, s_acc[e * 5 + 1]);
      atomicMax(w + 2, s_acc[e * 5 + 2]);
      atomicMax(w + 3, s_acc[e * 5 + 3]);
      atomicMax(w + 4, s_acc[e * 5 + 4]);
      published = true;
    }
  }
  if (published) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    int ticket = atomicAdd(ws_b + K * kWsIntsPerChannel, 1);
    s_last = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  DEV_STAMP_MAX(38);
  if (s_last) {
    __threadfence();
    // fold the kWsCopies copies: 16 consecutive lanes read-and-clear one channel's words of the 16 copies (one round
    // trip for the whole array), a half-warp shuffle reduction folds them, the copy-0 lane finalises the channel
    static_assert(kWsCopies == 16 && kThreads % kWsCopies == 0, "half-warp fold");
    const int n_e = (DIRECT ? 2 : 1) * K;
    constexpr int kPerPass = kThreads / kWsCopies, kBatch = 4;  // 16 channels per pass, the loads of 4 passes in one round trip
    const int c = threadIdx.x % kWsCopies;
    for (int e0 = 0; e0 < n_e; e0 += kPerPass * kBatch) {
      int v[kBatch][5];
#pragma unroll
      for (int p = 0; p < kBatch; ++p) {
        const int e = e0 + p * kPerPass + threadIdx.x / kWsCopies;
        const int set = e / K, i = e - set * K;
#pragma unroll
        for (int u = 0; u < 5; ++u) v[p][u] = 0;
        if (e < n_e && i >= 1) {
          // every publisher fenced before it took its ticket and ours was the last: plain (L2) loads see the final
          // values and plain stores may clear them -- 5 atomic exchanges per word made this fold a 3 us tail of its own
          int *w = ws_b + c * copy_stride + (set * (K + 1) + i) * kWsIntsPerChannel;
          const int4 lo = ld_dep(reinterpret_cast<const int4 *>(w));
          v[p][4] = ld_dep(w + 4);
          v[p][0] = lo.x; v[p][1] = lo.y; v[p][2] = lo.z; v[p][3] = lo.w;
        }
      }
#pragma unroll
      for (int p = 0; p < kBatch; ++p) {
        const int e = e0 + p * kPerPass + threadIdx.x / kWsCopies;
        const bool act = e < n_e;
        const int set = act ? e / K : 0, i = act ? e - set * K : 0;
        if (act && i >= 1 && v[p][0] != 0) {  // (a copy that saw no hit is still all zero)
          int *w = ws_b + c * copy_stride + (set * (K + 1) + i) * kWsIntsPerChannel;
          *reinterpret_cast<int4 *>(w) = make_int4(0, 0, 0, 0);
          w[4] = 0;
        }
#pragma unroll
        for (int d = kWsCopies / 2; d >= 1; d >>= 1) {
          v[p][0] += __shfl_xor_sync(0xffffffffu, v[p][0], d);
#pragma unroll
          for (int u = 1; u < 5; ++u) v[p][u] = max(v[p][u], __shfl_xor_sync(0xffffffffu, v[p][u], d));
        }
        if (act && c == 0) {
          int *bb = set ? bboxes_direct : bboxes_warp;
          const BoxFinalize &f = set ? fin_direct : fin_warp;
          if (i == 0) finalize_channel0(bb, (long long)b * K, f);
          else finalize_values(v[p][0], v[p][1], v[p][2], v[p][3], v[p][4], bb, (long long)b * K + i, f);
        }
      }
    }
    if (threadIdx.x == 0) atomicExch(ws_b + K * kWsIntsPerChannel, 0);
  }
  DEV_STAMP_MAX(1);
}

// ---- closed-form /16 cell rectangles (models/rmnet.py:245, :307+:356) -----------------------------
__global__ void cell_rects_kernel(const int *__restrict__ bboxes, int count, int pad_l, int pad_t, int h, int w,
                                  int skip_every, int *__restrict__ rects) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  const int4 bb = reinterpret_cast<const int4 *>(bboxes)[idx];
  int4 r;
  // cell (cy,cx) is live iff the padded full-res att at (16cy,16cx) is 1  <=>  x0+pad_l <= 16cx <= x1+pad_l
  r.x = max(0, (bb.x + pad_l + 15) >> 4);
  r.y = min(w - 1, (bb.y + pad_l) >> 4);
  r.z = max(0, (bb.z + pad_t + 15) >> 4);
  r.w = min(h - 1, (bb.w + pad_t) >> 4);
  if (skip_every > 0 && idx % skip_every == 0) r = make_int4(0, -1, 0, -1);  // channel 0: att_map is all zero
  if (r.x > r.y || r.z > r.w) r = make_int4(0, -1, 0, -1);
  reinterpret_cast<int4 *>(rects)[idx] = r;
}

int launch_fill(const int *bboxes, int B, int K, int H, int W, float *att_full, cudaStream_t st) {
  const long long total = (long long)B * K * H * W;
  const bool vec = ((long long)H * W) % 4 == 0 && W >= 4 && ((uintptr_t)att_full % 16 == 0);
  const int per = vec ? 4 : 1;
  long long blocks = (total / per + kThreads - 1) / kThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (vec) att_fill_kernel<4><<<(unsigned)blocks, kThreads, 0, st>>>(bboxes, K, H, W, total, att_full);
  else att_fill_kernel<1><<<(unsigned)blocks, kThreads, 0, st>>>(bboxes, K, H, W, total, att_full);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

int check_common(const void *mask, int B, int K, int H, int W, const int *bboxes, const void *ws, size_t ws_bytes) {
  RMNET_CHECK_ARG(mask && bboxes && ws, "null pointer argument");
  RMNET_CHECK_ARG(B > 0 && K > 1 && H > 0 && W > 0, "bad shape B=%d K=%d H=%d W=%d (need K >= 2)", B, K, H, W);
  RMNET_CHECK_ARG(H <= 32767 && W <= 32767, "H, W must be <= 32767 (reference min-init, reg_att_map_generator.cu:32)");
  RMNET_CHECK_ARG(B <= 65535, "B too large");
  RMNET_CHECK_ARG((uintptr_t)bboxes % 16 == 0, "bboxes must be 16-byte aligned");
  RMNET_CHECK_ARG((uintptr_t)ws % 16 == 0, "workspace must be 16-byte aligned");
  if (ws_bytes < rmnet_reg_att_map_workspace_bytes(B, K)) {
    set_error("workspace too small: %zu < %zu", ws_bytes, rmnet_reg_att_map_workspace_bytes(B, K));
    return RMNET_E_WORKSPACE;
  }
  return RMNET_OK;
}

}  // namespace
}  // namespace rmnet

using namespace rmnet;

extern "C" {

size_t rmnet_reg_att_map_workspace_bytes(int B, int K) {
  return (size_t)kWsCopies * B * 2 * (K + 1) * kWsIntsPerChannel * sizeof(int);  // two accumulator sets (warped, direct) x copies
}

static void make_finalize(BoxFinalize &fin, bool padded, int H, int W, int pad_l, int pad_r, int pad_t, int pad_b, float thr,
                          int n_pts, int loose, int k_scan, int *rects) {
  fin.k_scan = k_scan;
  fin.Hf = padded ? H + pad_t + pad_b : H;
  fin.Wf = padded ? W + pad_l + pad_r : W;
  fin.off_x = padded ? pad_l : 0;
  fin.off_y = padded ? pad_t : 0;
  fin.n_pts_threshold = n_pts;
  fin.loose = loose;
  // zero padding passes a threshold <= 0: every padded pixel is a point, the loosened box is the whole frame
  fin.force_full = (padded && thr <= 0.0f && (pad_l | pad_r | pad_t | pad_b) && (long long)fin.Hf * fin.Wf >= n_pts) ? 1 : 0;
  fin.rects = rects;
  fin.rect_pad_l = padded ? 0 : pad_l;
  fin.rect_pad_t = padded ? 0 : pad_t;
  fin.cell_h = (H + pad_t + pad_b) / 16;
  fin.cell_w = (W + pad_l + pad_r) / 16;
}

static int launch_scan(const float *mask, int B, int K, int H, int W, float thr, const BoxFinalize &fin, int *bboxes,
                       void *workspace, cudaStream_t st) {
  const long long n_pixels = (long long)H * W;
  const bool vec = n_pixels % 4 == 0 && W >= 4 && ((uintptr_t)mask % 16 == 0);
  // ~8 CTAs per SM over all channels: each CTA streams a contiguous chunk (multiple of the 4096-float tile)
  long long want_ctas = 148LL * 8;
  const int n_scan = fin.k_scan - 1;
  long long per_channel = (want_ctas + (long long)n_scan * B - 1) / ((long long)n_scan * B);
  if (per_channel < 1) per_channel = 1;
  long long elems = (n_pixels + per_channel - 1) / per_channel;
  elems = (elems + 4095) / 4096 * 4096;
  const int chunks = (int)((n_pixels + elems - 1) / elems);
  dim3 grid(chunks, n_scan, B);
  if (vec) bbox_scan_kernel<4><<<grid, kThreads, 0, st>>>(mask, K, H, W, thr, fin, (int)elems, bboxes, (int *)workspace);
  else bbox_scan_kernel<1><<<grid, kThreads, 0, st>>>(mask, K, H, W, thr, fin, (int)elems, bboxes, (int *)workspace);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

static int launch_frame_boxes(const float *prev_mask, const float *flow, int B, int K, int H, int W, int sampler, float thr,
                              const BoxFinalize &fin_warp, int *bboxes_warp, const BoxFinalize *fin_direct, int *bboxes_direct,
                              void *workspace, cudaStream_t st, float *clear = nullptr, int n_clear = 0) {
  const float inv_w = 1.0f / (float)(W - 1 > 1 ? W - 1 : 1);  // models/rmnet.py:265 max(W-1,1); host fp32 reciprocal like ATen
  const float inv_h = 1.0f / (float)(H - 1 > 1 ? H - 1 : 1);
  // persistent CTAs: one wave of (4 CTAs x 148 SMs) / B per batch item, or fewer when the frame is small
  const long long n_tiles = ((long long)H * W + kThreads - 1) / kThreads;
  long long per_b = (148LL * 4 + B - 1) / B;
  if (per_b > n_tiles) per_b = n_tiles;
  dim3 grid((unsigned)per_b, B);
  const size_t smem = 2 * K * 5 * sizeof(int);
  const BoxFinalize fd = fin_direct ? *fin_direct : fin_warp;
#define RMNET_LAUNCH_FB(O, D)                                                                                              \
  frame_boxes_kernel<O, D><<<grid, kThreads, smem, st>>>(prev_mask, flow, K, H, W, inv_w, inv_h, thr, fin_warp, fd, bboxes_warp, \
                                                         bboxes_direct, (int *)workspace, clear, n_clear)
  if (sampler == RMNET_SAMPLER_CUDNN) { if (fin_direct) RMNET_LAUNCH_FB(0, true); else RMNET_LAUNCH_FB(0, false); }
  else { if (fin_direct) RMNET_LAUNCH_FB(1, true); else RMNET_LAUNCH_FB(1, false); }
#undef RMNET_LAUNCH_FB
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}
static int launch_warp_scan(const float *prev_mask, const float *flow, int B, int K, int H, int W, int sampler, float thr,
                            const BoxFinalize &fin, int *bboxes, void *workspace, cudaStream_t st) {
  return launch_frame_boxes(prev_mask, flow, B, K, H, W, sampler, thr, fin, bboxes, nullptr, nullptr, workspace, st);
}

int rmnet_reg_att_map_forward(const float *mask, int B, int K, int H, int W, float prob_threshold,
                              int n_pts_threshold, int n_bbox_loose_pixels, int *bboxes, float *att_full,
                              void *workspace, size_t workspace_bytes, void *stream) {
  int rc = check_common(mask, B, K, H, W, bboxes, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  BoxFinalize fin = {H, W, 0, 0, n_pts_threshold, n_bbox_loose_pixels, 0, K, nullptr, 0, 0, 0, 0};
  if ((rc = launch_scan(mask, B, K, H, W, prob_threshold, fin, bboxes, workspace, st))) return rc;
  if (att_full) return launch_fill(bboxes, B, K, H, W, att_full, st);
  return RMNET_OK;
}

int rmnet_regional_boxes_forward(const float *mask, const float *flow, int B, int K, int H, int W, int sampler,
                                 float prob_threshold, int n_pts_threshold, int n_bbox_loose_pixels, int pad_l, int pad_r,
                                 int pad_t, int pad_b, int bbox_in_padded_frame, int k_scan, int *bboxes, int *cell_rects,
                                 void *workspace, size_t workspace_bytes, void *stream) {
  int rc = check_common(mask, B, K, H, W, bboxes, workspace, workspace_bytes);
  if (rc) return rc;
  RMNET_CHECK_ARG(pad_l >= 0 && pad_r >= 0 && pad_t >= 0 && pad_b >= 0, "negative padding");
  RMNET_CHECK_ARG((H + pad_t + pad_b) % 16 == 0 && (W + pad_l + pad_r) % 16 == 0, "padded frame must be a multiple of 16");
  RMNET_CHECK_ARG(H + pad_t + pad_b <= 32767 && W + pad_l + pad_r <= 32767, "padded frame too large");
  RMNET_CHECK_ARG(cell_rects == nullptr || (uintptr_t)cell_rects % 16 == 0, "cell_rects must be 16-byte aligned");
  RMNET_CHECK_ARG(flow == nullptr || sampler == RMNET_SAMPLER_CUDNN || sampler == RMNET_SAMPLER_ATEN, "bad sampler %d", sampler);
  RMNET_CHECK_ARG(K <= 1024, "K too large");
  cudaStream_t st = (cudaStream_t)stream;
  BoxFinalize fin;
  if (k_scan <= 0 || k_scan > K) k_scan = K;
  RMNET_CHECK_ARG(k_scan >= 2, "k_scan must be >= 2");
  make_finalize(fin, bbox_in_padded_frame != 0, H, W, pad_l, pad_r, pad_t, pad_b, prob_threshold, n_pts_threshold,
                n_bbox_loose_pixels, k_scan, cell_rects);
  if (flow) return launch_warp_scan(mask, flow, B, K, H, W, sampler, prob_threshold, fin, bboxes, workspace, st);
  return launch_scan(mask, B, K, H, W, prob_threshold, fin, bboxes, workspace, st);
}

int rmnet_warp_forward(const float *img0, const float *flow, int B, int C, int H, int W, int sampler, float *img1,
                       float *valid, void *stream) {
  RMNET_CHECK_ARG(img0 && flow && img1, "null pointer argument");
  RMNET_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "bad shape");
  RMNET_CHECK_ARG(sampler == RMNET_SAMPLER_CUDNN || sampler == RMNET_SAMPLER_ATEN, "bad sampler %d", sampler);
  const float inv_w = 1.0f / (float)(W - 1 > 1 ? W - 1 : 1);  // models/rmnet.py:265 max(W-1,1); host fp32 reciprocal like ATen
  const float inv_h = 1.0f / (float)(H - 1 > 1 ? H - 1 : 1);
  dim3 grid((unsigned)(((long long)H * W + kThreads - 1) / kThreads), B);
  if (sampler == RMNET_SAMPLER_CUDNN)
    warp_kernel<0><<<grid, kThreads, 0, (cudaStream_t)stream>>>(img0, flow, C, H, W, inv_w, inv_h, img1, valid);
  else
    warp_kernel<1><<<grid, kThreads, 0, (cudaStream_t)stream>>>(img0, flow, C, H, W, inv_w, inv_h, img1, valid);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

int rmnet_frame_regions_forward(const float *prev_mask, const float *flow, int B, int K, int H, int W, int sampler,
                                float prob_threshold, int n_pts_threshold, int n_bbox_loose_pixels, int pad_l, int pad_r,
                                int pad_t, int pad_b, int k_scan, int *mem_bboxes, int *mem_rects, int *cur_bboxes, int *cur_rects,
                                void *workspace, size_t workspace_bytes, void *stream) {
  return rmnet::frame_regions_chain_head(prev_mask, flow, B, K, H, W, sampler, prob_threshold, n_pts_threshold, n_bbox_loose_pixels,
                                         pad_l, pad_r, pad_t, pad_b, k_scan, mem_bboxes, mem_rects, cur_bboxes, cur_rects, workspace,
                                         workspace_bytes, nullptr, 0, stream);
}
}  // extern "C"

// Internal entry (also the head of rmnet_frame_step's PDL chain): `clear[0..n_clear)` is zeroed by the same launch.
int rmnet::frame_regions_chain_head(const float *prev_mask, const float *flow, int B, int K, int H, int W, int sampler,
                                    float prob_threshold, int n_pts_threshold, int n_bbox_loose_pixels, int pad_l, int pad_r,
                                    int pad_t, int pad_b, int k_scan, int *mem_bboxes, int *mem_rects, int *cur_bboxes,
                                    int *cur_rects, void *workspace, size_t workspace_bytes, float *clear, int n_clear,
                                    void *stream) {
  int rc = check_common(prev_mask, B, K, H, W, mem_bboxes, workspace, workspace_bytes);
  if (rc) return rc;
  RMNET_CHECK_ARG(flow && cur_bboxes && mem_rects && cur_rects, "null pointer argument");
  RMNET_CHECK_ARG(pad_l >= 0 && pad_r >= 0 && pad_t >= 0 && pad_b >= 0, "negative padding");
  RMNET_CHECK_ARG((H + pad_t + pad_b) % 16 == 0 && (W + pad_l + pad_r) % 16 == 0, "padded frame must be a multiple of 16");
  RMNET_CHECK_ARG(H + pad_t + pad_b <= 32767 && W + pad_l + pad_r <= 32767, "padded frame too large");
  RMNET_CHECK_ARG(((uintptr_t)cur_bboxes | (uintptr_t)mem_rects | (uintptr_t)cur_rects) % 16 == 0, "outputs must be 16-byte aligned");
  RMNET_CHECK_ARG(sampler == RMNET_SAMPLER_CUDNN || sampler == RMNET_SAMPLER_ATEN, "bad sampler %d", sampler);
  RMNET_CHECK_ARG(K <= 512, "K too large");
  BoxFinalize fw, fd;
  if (k_scan <= 0 || k_scan > K) k_scan = K;
  RMNET_CHECK_ARG(k_scan >= 2, "k_scan must be >= 2");
  make_finalize(fw, false, H, W, pad_l, pad_r, pad_t, pad_b, prob_threshold, n_pts_threshold, n_bbox_loose_pixels, k_scan, cur_rects);
  make_finalize(fd, true, H, W, pad_l, pad_r, pad_t, pad_b, prob_threshold, n_pts_threshold, n_bbox_loose_pixels, k_scan, mem_rects);
  return launch_frame_boxes(prev_mask, flow, B, K, H, W, sampler, prob_threshold, fw, cur_bboxes, &fd, mem_bboxes, workspace,
                            (cudaStream_t)stream, clear, n_clear);
}
extern "C" {

int rmnet_warp_att_map_forward(const float *prev_mask, const float *flow, int B, int K, int H, int W, int sampler,
                               float prob_threshold, int n_pts_threshold, int n_bbox_loose_pixels, int *bboxes,
                               float *att_full, void *workspace, size_t workspace_bytes, void *stream) {
  int rc = check_common(prev_mask, B, K, H, W, bboxes, workspace, workspace_bytes);
  if (rc) return rc;
  RMNET_CHECK_ARG(flow != nullptr, "null flow");
  RMNET_CHECK_ARG(K <= 1024, "K too large");
  RMNET_CHECK_ARG(sampler == RMNET_SAMPLER_CUDNN || sampler == RMNET_SAMPLER_ATEN, "bad sampler %d", sampler);
  cudaStream_t st = (cudaStream_t)stream;
  BoxFinalize fin = {H, W, 0, 0, n_pts_threshold, n_bbox_loose_pixels, 0, K, nullptr, 0, 0, 0, 0};
  if ((rc = launch_warp_scan(prev_mask, flow, B, K, H, W, sampler, prob_threshold, fin, bboxes, workspace, st))) return rc;
  if (att_full) return launch_fill(bboxes, B, K, H, W, att_full, st);
  return RMNET_OK;
}

int rmnet_cell_rects_from_bboxes(const int *bboxes, int count, int pad_l, int pad_t, int h, int w,
                                 int skip_channel0_every, int *rects, void *stream) {
  RMNET_CHECK_ARG(bboxes && rects && count > 0 && h > 0 && w > 0, "bad argument");
  RMNET_CHECK_ARG((uintptr_t)bboxes % 16 == 0 && (uintptr_t)rects % 16 == 0, "bboxes/rects must be 16-byte aligned");
  cell_rects_kernel<<<cdiv(count, 128), 128, 0, (cudaStream_t)stream>>>(bboxes, count, pad_l, pad_t, h, w,
                                                                         skip_channel0_every, rects);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // extern "C"
