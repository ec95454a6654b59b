/*
 * oracle/rmnet_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU, sequential restatement of the integer / byte-exact parts of
 * hzxie/RMNet's regional memory-read hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (rmnet_b200/) never does and fails loudly when its CUDA library is missing.
 *
 * Every function cites the reference file:line (relative to the hzxie/RMNet tree) that
 * it follows.  Build: `make -C oracle` (gcc -O2 -ffp-contract=off: no FMA contraction,
 * like the reference's own -O2 -g build of the NumPy extension).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * Regional attention map generator.
 * Follows extensions/reg_att_map_generator/reg_att_map_generator.cu:31-92 under
 * *sequential* semantics (the kernel's unsynchronised init at :31-34 is a formal race).
 *   mask   [B,K,H,W] f32;  bboxes [B,K,4] i32 = (x_min, x_max, y_min, y_max) inclusive;
 *   att    [B,K,H,W] f32 in {0,1} (nullable).  Channel 0 is never touched (:32,:37,:56,:81),
 *   so it keeps the zero fill of the host wrapper (.cu:104-109).
 * ---------------------------------------------------------------------------------- */
ORACLE_API int oracle_reg_att_map(const float *mask, int B, int K, int H, int W,
                                  float prob_threshold, int n_pts_threshold,
                                  int n_bbox_loose_pixels, int *bboxes, float *att) {
  const long n_pixels = (long)H * W;
  for (int b = 0; b < B; ++b) {
    const float *m_b = mask + (long)b * K * n_pixels;
    int *bb_b = bboxes + (long)b * K * 4;
    float *att_b = att ? att + (long)b * K * n_pixels : NULL;
    /* torch::zeros for every output (.cu:104-109) */
    memset(bb_b, 0, sizeof(int) * 4 * K);
    if (att_b) memset(att_b, 0, sizeof(float) * K * n_pixels);
    for (int i = 1; i < K; ++i) {
      int x_min = 32767, x_max = 0, y_min = 32767, y_max = 0, n_points = 0; /* :31-34 */
      const float *m = m_b + (long)i * n_pixels;
      for (long j = 0; j < n_pixels; ++j) { /* :38-49 */
        int x = (int)(j % W), y = (int)(j / W);
        if (m[j] >= prob_threshold) { /* NaN compares false */
          ++n_points;
          if (x < x_min) x_min = x;
          if (x > x_max) x_max = x;
          if (y < y_min) y_min = y;
          if (y > y_max) y_max = y;
        }
      }
      if (n_points < n_pts_threshold) { /* :57-61 */
        x_min = 0; x_max = W - 1; y_min = 0; y_max = H - 1;
      } else { /* :63-74 */
        x_min = x_min <= n_bbox_loose_pixels ? 0 : x_min - n_bbox_loose_pixels;
        x_max = x_max + n_bbox_loose_pixels >= W ? W - 1 : x_max + n_bbox_loose_pixels;
        y_min = y_min <= n_bbox_loose_pixels ? 0 : y_min - n_bbox_loose_pixels;
        y_max = y_max + n_bbox_loose_pixels >= H ? H - 1 : y_max + n_bbox_loose_pixels;
      }
      bb_b[i * 4 + 0] = x_min; bb_b[i * 4 + 1] = x_max;
      bb_b[i * 4 + 2] = y_min; bb_b[i * 4 + 3] = y_max;
      if (att_b) { /* :81-92 */
        float *a = att_b + (long)i * n_pixels;
        for (long j = 0; j < n_pixels; ++j) {
          int x = (int)(j % W), y = (int)(j / W);
          if (x >= x_min && x <= x_max && y >= y_min && y <= y_max) a[j] = 1.0f;
        }
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * utils/helpers.py:105-124 pad_divide_by(d): pad = (lw, uw, lh, uh), zero fill (F.pad).
 * ---------------------------------------------------------------------------------- */
ORACLE_API void oracle_pad_amounts(int h, int w, int d, int pad[4]) {
  int new_h = (h % d > 0) ? h + d - h % d : h;
  int new_w = (w % d > 0) ? w + d - w % d : w;
  int lh = (new_h - h) / 2, uh = (new_h - h) - (new_h - h) / 2;
  int lw = (new_w - w) / 2, uw = (new_w - w) - (new_w - w) / 2;
  pad[0] = lw; pad[1] = uw; pad[2] = lh; pad[3] = uh;
}

/* zero-pad a [C,H,W] map to [C,H+lh+uh,W+lw+uw] (utils/helpers.py:121-122) */
ORACLE_API void oracle_pad2d(const float *in, int C, int H, int W, const int pad[4], float *out) {
  int Hp = H + pad[2] + pad[3], Wp = W + pad[0] + pad[1];
  memset(out, 0, sizeof(float) * (size_t)C * Hp * Wp);
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < H; ++y)
      memcpy(out + ((size_t)c * Hp + y + pad[2]) * Wp + pad[0], in + ((size_t)c * H + y) * W,
             sizeof(float) * W);
}

/* ------------------------------------------------------------------------------------
 * F.interpolate(att_map, scale_factor=1/16) (mode='nearest') as used at
 * models/rmnet.py:245 and :356: out[c,y,x] = in[c, min(floor(y*16), Hp-1), min(floor(x*16), Wp-1)]
 * with output size floor(Hp/16) x floor(Wp/16) (ATen upsample_nearest, scale = 1/0.0625 = 16).
 * ---------------------------------------------------------------------------------- */
ORACLE_API void oracle_downsample16_nearest(const float *in, int C, int Hp, int Wp, float *out) {
  int h = Hp / 16, w = Wp / 16;
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int sy = y * 16 < Hp - 1 ? y * 16 : Hp - 1;
        int sx = x * 16 < Wp - 1 ? x * 16 : Wp - 1;
        out[((size_t)c * h + y) * w + x] = in[((size_t)c * Hp + sy) * Wp + sx];
      }
}

/* ------------------------------------------------------------------------------------
 * RMNet.warp, models/rmnet.py:252-278: backward warp of img0 by flow with
 * F.grid_sample(bilinear, zeros padding, align_corners=True) on the image and on a ones
 * tensor; validity = (sampled ones >= 0.9999); img1 = sampled * validity.
 *   img0 [C,H,W], flow [2,H,W] (ch0 = x, ch1 = y, pixels) -> img1 [C,H,W], valid [H,W]
 *
 * `arith` selects which torch backend's floating-point evaluation order is mirrored
 * (torch 2.11; both evaluate the same real-valued formula):
 *   0 = CPU path : true division  2*v / (W-1)            (ATen/native/cpu BinaryOpsKernel)
 *   1 = CUDA path: 2*v * (1/(W-1)) reciprocal multiply   (cuda/BinaryDivTrueKernel.cu, CPU-scalar divisor)
 * Every backend accumulates the four taps as an FMA chain; `tap_order` selects the order:
 *   0 = nw, ne, sw, se : acc = fma(v_se, se, fma(v_sw, sw, fma(v_ne, ne, v_nw * nw)))
 *       ATen's own kernels (cuda/GridSampler.cu under nvcc -fmad=true; cpu/GridSamplerKernel.cpp as built in
 *       the torch 2.11 wheel) -- established empirically: 0 mismatches vs F.grid_sample on CPU and vs the CUDA
 *       kernel with torch.backends.cudnn.enabled=False.
 *   1 = ne, nw, sw, se : acc = fma(v_se, se, fma(v_sw, sw, fma(v_nw, nw, v_ne * ne)))
 *       cudnnSpatialTfSamplerForward (cuDNN 9.x), which torch dispatches to on CUDA for bilinear / zeros /
 *       align_corners=True when cuDNN is enabled (ATen/native/GridSampler.cpp cond_cudnn_grid_sampler) -- the
 *       reference's default GPU path.  cuDNN also forms the east / south weights as 1 - (west / north weight)
 *       instead of x - floor(x).  Established empirically on a B200 (tools/warp_probe*.py): 0 mismatches over all
 *       in-bounds samples incl. the floor == 0 row / column, while every other of 200+ candidate orders / weight
 *       forms mismatches by 1 ulp somewhere.  (Samples with an out-of-bounds tap can still differ by 1 ulp; they are
 *       zeroed by the validity mask unless they are within 1e-4 px of the border.)
 * Unnormalisation ((g+1)/2)*(size-1): ATen/native/cuda/GridSampler.cuh:23-31 (= GridSampler.h on CPU).
 * ---------------------------------------------------------------------------------- */
static inline float oracle_norm_coord(float v, int size, int arith) {
  int d = size - 1 > 1 ? size - 1 : 1;
  float t = 2.0f * v;
  if (arith == 0) t = t / (float)d;
  else { float inv = 1.0f / (float)d; t = t * inv; }
  return t - 1.0f;
}

ORACLE_API int oracle_warp(const float *img0, const float *flow, int C, int H, int W, int arith, int tap_order,
                           float *img1, float *valid) {
  const long n_pixels = (long)H * W;
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      long j = (long)y * W + x;
      float vx = (float)x + flow[j];              /* :263 grid + flow */
      float vy = (float)y + flow[n_pixels + j];
      float gx = oracle_norm_coord(vx, W, arith); /* :265-266 */
      float gy = oracle_norm_coord(vy, H, arith);
      float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1);
      float iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
      float fx = floorf(ix), fy = floorf(iy);
      float wx0 = (fx + 1.0f) - ix, wy0 = (fy + 1.0f) - iy; /* weight of the west / north tap */
      float wx1 = ix - fx, wy1 = iy - fy;                   /* ATen: weight of the east / south tap */
      if (tap_order == 1) { /* cuDNN derives it as 1 - w0, which differs when w0 is inexact (floor == 0 or -1) */
        wx1 = 1.0f - wx0;
        wy1 = 1.0f - wy0;
      }
      float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
      /* float->int like ATen: static_cast<int>(floor(.)); guard non-finite / huge */
      int x0, y0;
      if (!(fx > -2.0f && fx < (float)W + 1.0f)) x0 = -2; else x0 = (int)fx;
      if (!(fy > -2.0f && fy < (float)H + 1.0f)) y0 = -2; else y0 = (int)fy;
      int x1 = x0 + 1, y1 = y0 + 1;
      int in_nw = (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H);
      int in_ne = (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H);
      int in_sw = (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H);
      int in_se = (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H);
      /* sampled ones tensor (:272-273) */
      float ones = 0.0f;
      if (tap_order == 0) {
        if (in_nw) ones = fmaf(1.0f, nw, ones);
        if (in_ne) ones = fmaf(1.0f, ne, ones);
      } else {
        if (in_ne) ones = fmaf(1.0f, ne, ones);
        if (in_nw) ones = fmaf(1.0f, nw, ones);
      }
      if (in_sw) ones = fmaf(1.0f, sw, ones);
      if (in_se) ones = fmaf(1.0f, se, ones);
      float m = ones;
      if (m < 0.9999f) m = 0.0f; /* :274 */
      if (m > 0.0f) m = 1.0f;    /* :275 (NaN stays NaN, as in torch) */
      if (valid) valid[j] = m;
      for (int c = 0; c < C; ++c) {
        const float *p = img0 + (long)c * n_pixels;
        float v_nw = in_nw ? p[(long)y0 * W + x0] : 0.0f;
        float v_ne = in_ne ? p[(long)y0 * W + x1] : 0.0f;
        float v_sw = in_sw ? p[(long)y1 * W + x0] : 0.0f;
        float v_se = in_se ? p[(long)y1 * W + x1] : 0.0f;
        float s = 0.0f;
        if (tap_order == 0) {
          if (in_nw) s = fmaf(v_nw, nw, s);
          if (in_ne) s = fmaf(v_ne, ne, s);
        } else {
          if (in_ne) s = fmaf(v_ne, ne, s);
          if (in_nw) s = fmaf(v_nw, nw, s);
        }
        if (in_sw) s = fmaf(v_sw, sw, s);
        if (in_se) s = fmaf(v_se, se, s);
        img1[(long)c * n_pixels + j] = s * m; /* :277 */
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * flow_affine_transformation.update_optical_flow,
 * extensions/flow_affine_transformation/flow_affine_transformation.cpp:63-83.
 * All float32, evaluation order (a*j + b*i) + c, no FMA (-ffp-contract=off), size_t
 * indices converted to float, std::round = half away from zero (roundf), the y1 update
 * uses the ALREADY-UPDATED x1 (:72-73), clamps at :75-78.
 * ---------------------------------------------------------------------------------- */
ORACLE_API int oracle_update_optical_flow(const float *of, const float *m1, const float *m2, int H,
                                          int W, float *out) {
  size_t height = (size_t)H, width = (size_t)W;
  for (size_t i = 0; i < height; ++i) {
    for (size_t j = 0; j < width; ++j) {
      size_t idx = (i * width + j) * 2;
      float x2 = roundf(m2[0] * j + m2[1] * i + m2[2]);
      float y2 = roundf(m2[3] * j + m2[4] * i + m2[5]);
      float x1 = j + of[idx];
      float y1 = i + of[idx + 1];
      x1 = roundf(m1[0] * x1 + m1[1] * y1 + m1[2]);
      y1 = roundf(m1[3] * x1 + m1[4] * y1 + m1[5]);
      x1 = x1 < 0 ? 0 : (x1 >= width ? width - 1 : x1);
      y1 = y1 < 0 ? 0 : (y1 >= height ? height - 1 : y1);
      x2 = x2 < 0 ? 0 : (x2 >= width ? width - 1 : x2);
      y2 = y2 < 0 ? 0 : (y2 >= height ? height - 1 : y2);
      out[idx] = x1 - x2;
      out[idx + 1] = y1 - y2;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * MemoryReader.forward, models/rmnet.py:147-165, one object, scalar loops with double
 * accumulation (the "floor" reference for small cases; the numpy restatement in
 * oracle/memory_read.py is the BLAS-backed one used for timing and larger sizes).
 *   m_key [Ck,M] m_val [Cv,M] q_key [Ck,N] q_val [Cv,N] -> mem_val [2*Cv? no: Cv+Cv, N], p [M,N] nullable
 * ---------------------------------------------------------------------------------- */
ORACLE_API int oracle_memory_read_f64(const float *m_key, const float *m_val, const float *q_key,
                                      const float *q_val, int Ck, int Cv, int M, int N,
                                      float *mem_val, float *p_out) {
  double *s = (double *)malloc(sizeof(double) * (size_t)M);
  double *acc = (double *)malloc(sizeof(double) * (size_t)Cv);
  if (!s || !acc) { free(s); free(acc); return -1; }
  const double inv_sqrt = 1.0 / sqrt((double)Ck); /* :156 */
  for (int q = 0; q < N; ++q) {
    double mx = -INFINITY;
    for (int j = 0; j < M; ++j) { /* :155 */
      double d = 0.0;
      for (int c = 0; c < Ck; ++c) d += (double)m_key[(size_t)c * M + j] * (double)q_key[(size_t)c * N + q];
      d *= inv_sqrt;
      s[j] = d;
      if (d > mx) mx = d;
    }
    double sum = 0.0;
    for (int j = 0; j < M; ++j) { s[j] = exp(s[j] - mx); sum += s[j]; } /* :157 softmax over dim=1 */
    for (int c = 0; c < Cv; ++c) acc[c] = 0.0;
    for (int j = 0; j < M; ++j) {
      double pj = s[j] / sum;
      if (p_out) p_out[(size_t)j * N + q] = (float)pj;
      for (int c = 0; c < Cv; ++c) acc[c] += (double)m_val[(size_t)c * M + j] * pj; /* :160 */
    }
    for (int c = 0; c < Cv; ++c) {
      mem_val[(size_t)c * N + q] = (float)acc[c];
      mem_val[(size_t)(Cv + c) * N + q] = q_val[(size_t)c * N + q]; /* :163 cat */
    }
  }
  free(s); free(acc);
  return 0;
}
