// mask_epilogue.cu -- the per-frame tail of RMNet.segment / RMNet.forward after the decoder, in ONE pass.
//
// Replaces (models/rmnet.py):
//   :368-370  ps = F.softmax(logits, dim=1)[:, 1]                       decoder logits [n,2,Hp,Wp]
//   :289-302  soft_aggregation: em[0] = prod(1 - ps), em[1..n] = ps, em[n+1..] = 0; clamp(1e-7, 1-1e-7); log(em / (1 - em))
//   :376-380  un-pad (pad_divide_by amounts)
//   :436-448  new-object / non-existing-object overrides of whole logit channels
//   :450      est_masks[:, t] = F.softmax(logit, dim=1)
// -- about ten elementwise ATen passes over [K,Hp,Wp] in the reference.  HBM-bound: 8*n*Hp*Wp bytes read,
// 4*K*H*W (+ 4*K*H*W when the logit map is requested) written; one pixel per thread, coalesced along x.
// The arithmetic mirrors the ATen kernels op for op (separately rounded fp32 steps, libdevice expf / logf, the product
// over objects in torch.prod's four-accumulator order) so that the logit map agrees with the reference's to the last bit
// wherever the two libm's agree; the parity tests bound the difference by 1e-3 (north_star) and report the exact share.
#include "common.cuh"

namespace rmnet {
namespace {

constexpr int kThreads = 256;
struct ChannelModes { unsigned char m[64]; };

template <int KMAX>
__global__ void __launch_bounds__(kThreads)
mask_epilogue_kernel(const float *__restrict__ dec_logits, int n_obj, int K, int H, int W, int Hp, int Wp, int pad_l, int pad_t,
                     ChannelModes modes, const int *__restrict__ new_mask, float *__restrict__ logit_out,
                     float *__restrict__ est_mask) {
  const long long n_pixels = (long long)H * W;
  const long long j = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (j >= n_pixels) return;
  const int y = (int)(j / W), x = (int)(j - (long long)y * W);
  const long long jp = (long long)(y + pad_t) * Wp + (x + pad_l);  // the same pixel in the padded decoder output
  const long long plane_p = (long long)Hp * Wp;
  const float lo = (float)1e-7, hi = (float)(1.0 - 1e-7);  // torch.clamp casts its python-float bounds to float32

  float lg[KMAX];
  // ---- ps of every object (two-class softmax, ATen order: max, exp(x - max), sum c = 0,1, divide), background product
  float acc[4] = {1.f, 1.f, 1.f, 1.f};  // torch.prod over dim 0: four strided accumulators, combined 0..3
#pragma unroll
  for (int o = 0; o < KMAX - 1; ++o) {
    if (o < n_obj) {
      const float l0 = __ldg(dec_logits + (long long)(2 * o) * plane_p + jp);
      const float l1 = __ldg(dec_logits + (long long)(2 * o + 1) * plane_p + jp);
      const float m = fmaxf(l0, l1);
      const float e0 = expf(__fsub_rn(l0, m)), e1 = expf(__fsub_rn(l1, m));
      const float ps = __fdiv_rn(e1, __fadd_rn(e0, e1));
      lg[o + 1] = ps;
      acc[o & 3] = __fmul_rn(acc[o & 3], __fsub_rn(1.0f, ps));
    }
  }
  lg[0] = __fmul_rn(__fmul_rn(__fmul_rn(acc[0], acc[1]), acc[2]), acc[3]);
  // ---- clamp, logit, per-channel overrides.  Channels above n_obj hold em = 0 -> the clamp floor -> one constant logit
  //      (no division / logarithm per pixel); the kernel is bound by these multi-instruction fp32 functions, not by HBM.
  const float floor_logit = logf(__fdiv_rn(lo, __fsub_rn(1.0f, lo)));
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < KMAX; ++c) {
    if (c < K) {
      const int mode = modes.m[c];
      float v;
      if (mode == RMNET_CH_ABSENT) v = -16.1181f;                                                        // :448
      else if (mode == RMNET_CH_NEW)
        v = __fsub_rn(__fmul_rn((float)__ldg(new_mask + (long long)c * n_pixels + j), 32.0605f), 16.1181f);  // :442
      else if (c > n_obj) v = floor_logit;
      else {
        const float em = fminf(fmaxf(lg[c], lo), hi);
        v = logf(__fdiv_rn(em, __fsub_rn(1.0f, em)));
      }
      lg[c] = v;
      if (logit_out) logit_out[(long long)c * n_pixels + j] = v;
      mx = fmaxf(mx, v);
    }
  }
  // ---- channel softmax (:450): exp(x - max) summed c = 0..K-1, then exp(x - max) / sum.  The constant channels share
  //      one exponential and one quotient (same bits as recomputing them), the sum keeps the reference's order.
  const float e_floor = expf(__fsub_rn(floor_logit, mx));
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < KMAX; ++c) {
    if (c < K) {
      const bool is_floor = c > n_obj && modes.m[c] == RMNET_CH_KEEP;
      lg[c] = is_floor ? e_floor : expf(__fsub_rn(lg[c], mx));
      sum = __fadd_rn(sum, lg[c]);
    }
  }
  const float q_floor = __fdiv_rn(e_floor, sum);
#pragma unroll
  for (int c = 0; c < KMAX; ++c) {
    if (c < K) {
      const bool is_floor = c > n_obj && modes.m[c] == RMNET_CH_KEEP;
      est_mask[(long long)c * n_pixels + j] = is_floor ? q_floor : __fdiv_rn(lg[c], sum);
    }
  }
}

}  // namespace
}  // namespace rmnet

using namespace rmnet;
extern "C" {

int rmnet_mask_epilogue_forward(const float *dec_logits, int n_obj, int K, int H, int W, int pad_l, int pad_r, int pad_t,
                                int pad_b, const int *channel_mode_host, const int *new_mask, float *logit_out,
                                float *est_mask, void *stream) {
  RMNET_CHECK_ARG(dec_logits && est_mask, "null pointer argument");
  RMNET_CHECK_ARG(n_obj > 0 && K >= 2 && n_obj < K && K <= 64, "bad shape n_obj=%d K=%d (need 0 < n_obj < K <= 64)", n_obj, K);
  RMNET_CHECK_ARG(H > 0 && W > 0 && pad_l >= 0 && pad_r >= 0 && pad_t >= 0 && pad_b >= 0, "bad frame size / padding");
  ChannelModes modes;
  memset(&modes, 0, sizeof(modes));
  bool any_new = false;
  if (channel_mode_host) {
    for (int c = 0; c < K; ++c) {
      const int m = channel_mode_host[c];
      RMNET_CHECK_ARG(m == RMNET_CH_KEEP || m == RMNET_CH_ABSENT || m == RMNET_CH_NEW, "bad channel mode %d for channel %d", m, c);
      modes.m[c] = (unsigned char)m;
      any_new |= m == RMNET_CH_NEW;
    }
  }
  RMNET_CHECK_ARG(!any_new || new_mask, "a channel is marked RMNET_CH_NEW but new_mask is NULL");
  const int Hp = H + pad_t + pad_b, Wp = W + pad_l + pad_r;
  const long long n_pixels = (long long)H * W;
  dim3 grid((unsigned)((n_pixels + kThreads - 1) / kThreads));
  cudaStream_t st = (cudaStream_t)stream;
#define RMNET_LAUNCH_EPI(KM)                                                                                                   \
  mask_epilogue_kernel<KM><<<grid, kThreads, 0, st>>>(dec_logits, n_obj, K, H, W, Hp, Wp, pad_l, pad_t, modes, new_mask, logit_out, \
                                                      est_mask)
  if (K <= 12) RMNET_LAUNCH_EPI(12);
  else if (K <= 32) RMNET_LAUNCH_EPI(32);
  else RMNET_LAUNCH_EPI(64);
#undef RMNET_LAUNCH_EPI
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}
}
