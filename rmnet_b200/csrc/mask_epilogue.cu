// mask_epilogue.cu -- the per-frame tail of RMNet.segment / RMNet.forward after the decoder, in ONE pass.
//
// Replaces (models/rmnet.py):
//   :368-370  ps = F.softmax(logits, dim=1)[:, 1]                       decoder logits [n,2,Hp,Wp]
//   :289-302  soft_aggregation: em[0] = prod(1 - ps), em[1..n] = ps, em[n+1..] = 0; clamp(1e-7, 1-1e-7); log(em / (1 - em))
//   :376-380  un-pad (pad_divide_by amounts)
//   :436-448  new-object / non-existing-object overrides of whole logit channels
//   :450      est_masks[:, t] = F.softmax(logit, dim=1)
// -- about ten elementwise ATen passes over [K,Hp,Wp] in the reference.  Algorithmic traffic: 8*n*Hp*Wp bytes read,
// 4*K*H*W (+ 4*K*H*W when the logit map is requested) written; one pixel per thread, coalesced along x.  In practice
// the kernel is issue-bound on the IEEE division / logf / expf sequences that bit-exactness needs.
// The arithmetic mirrors the ATen kernels op for op (separately rounded fp32 steps, libdevice expf / logf, the product
// over objects in torch.prod's four-accumulator order) so that the logit map agrees with the reference's to the last bit
// wherever the two libm's agree; the parity tests bound the difference by 1e-3 (north_star) and report the exact share.
#include "common.cuh"

namespace rmnet {
namespace {

constexpr int kThreads = 256;
// per-channel override mode, 2 bits per channel (K <= 64): decoded with shifts, no dynamically indexed parameter array
struct ChannelModes {
  unsigned long long bits[2];
  __host__ __device__ int get(int c) const { return (int)(((c < 32 ? bits[0] : bits[1]) >> (2 * (c & 31))) & 3ull); }
  __host__ void set(int c, int m) { bits[c >> 5] |= (unsigned long long)m << (2 * (c & 31)); }
};

// NOBJ = compile-time bound on n_obj: the channels 0..n_obj (background + real objects) are the only ones with per-pixel
// transcendental work and live in registers (unrolled); the channels above n_obj are the clamp floor or a whole-channel
// override -- constants per pixel -- and are walked by a plain runtime loop.  Summation orders are the reference's.
// HI_MODES = false: no override among the channels above n_obj (always the case in the reference's loop, whose overrides
// touch j <= n_max_objects only): those channels are the clamp floor and need no mode decoding at all.
template <int NOBJ, bool HI_MODES>
__global__ void __launch_bounds__(kThreads)
mask_epilogue_kernel(const float *__restrict__ dec_logits, int n_obj, int K, int H, int W, int Hp, int Wp, int pad_l, int pad_t,
                     ChannelModes modes, const int *__restrict__ new_mask, float *__restrict__ logit_out,
                     float *__restrict__ est_mask) {
  const long long n_pixels = (long long)H * W;
  const long long j = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (j >= n_pixels) return;
  const unsigned uj = (unsigned)j;  // H * W < 2^31 (checked by the host): 32-bit division
  const int y = (int)(uj / (unsigned)W), x = (int)(uj - (unsigned)y * (unsigned)W);
  const long long jp = (long long)(y + pad_t) * Wp + (x + pad_l);  // the same pixel in the padded decoder output
  const long long plane_p = (long long)Hp * Wp;
  const float lo = (float)1e-7, hi = (float)(1.0 - 1e-7);  // torch.clamp casts its python-float bounds to float32

  float lg[NOBJ + 1];
  // ---- ps of every object (two-class softmax, ATen order: max, exp(x - max), sum c = 0,1, divide), background product
  float acc[4] = {1.f, 1.f, 1.f, 1.f};  // torch.prod over dim 0: four strided accumulators, combined 0..3
#pragma unroll
  for (int o = 0; o < NOBJ; ++o) {
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

  // value of a channel that needs no per-pixel arithmetic: the clamp floor (em = 0, :293 + :300-301) or an override
  const float floor_logit = logf(__fdiv_rn(lo, __fsub_rn(1.0f, lo)));
  auto const_channel = [&](int c, int mode) -> float {
    if (mode == RMNET_CH_ABSENT) return -16.1181f;                                                              // :448
    if (mode == RMNET_CH_NEW)
      return __fsub_rn(__fmul_rn((float)__ldg(new_mask + (long long)c * n_pixels + j), 32.0605f), 16.1181f);   // :442
    return floor_logit;
  };
  // ---- pass 1: logits (clamp, log(em / (1 - em)), overrides), their maximum
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c <= NOBJ; ++c) {
    if (c <= n_obj) {
      const int mode = modes.get(c);
      float v;
      if (mode != RMNET_CH_KEEP) v = const_channel(c, mode);
      else {
        const float em = fminf(fmaxf(lg[c], lo), hi);
        v = logf(__fdiv_rn(em, __fsub_rn(1.0f, em)));
      }
      lg[c] = v;
      if (logit_out) logit_out[(long long)c * n_pixels + j] = v;
      mx = fmaxf(mx, v);
    }
  }
  for (int c = n_obj + 1; c < K; ++c) {
    const float v = HI_MODES ? const_channel(c, modes.get(c)) : floor_logit;
    if (logit_out) logit_out[(long long)c * n_pixels + j] = v;
    mx = fmaxf(mx, v);
  }
  // ---- pass 2 (:450): exp(x - max) summed c = 0..K-1; the floor channels share one exponential (same bits)
  const float e_floor = expf(__fsub_rn(floor_logit, mx));
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c <= NOBJ; ++c) {
    if (c <= n_obj) {
      lg[c] = expf(__fsub_rn(lg[c], mx));
      sum = __fadd_rn(sum, lg[c]);
    }
  }
  for (int c = n_obj + 1; c < K; ++c) {
    const int mode = HI_MODES ? modes.get(c) : RMNET_CH_KEEP;
    sum = __fadd_rn(sum, mode == RMNET_CH_KEEP ? e_floor : expf(__fsub_rn(const_channel(c, mode), mx)));
  }
  // ---- pass 3: exp(x - max) / sum; the floor channels share one quotient
#pragma unroll
  for (int c = 0; c <= NOBJ; ++c)
    if (c <= n_obj) est_mask[(long long)c * n_pixels + j] = __fdiv_rn(lg[c], sum);
  const float q_floor = __fdiv_rn(e_floor, sum);
  for (int c = n_obj + 1; c < K; ++c) {
    const int mode = HI_MODES ? modes.get(c) : RMNET_CH_KEEP;
    est_mask[(long long)c * n_pixels + j] =
        mode == RMNET_CH_KEEP ? q_floor : __fdiv_rn(expf(__fsub_rn(const_channel(c, mode), mx)), sum);
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
  RMNET_CHECK_ARG((long long)H * W < (1LL << 31), "frame too large");
  ChannelModes modes;
  memset(&modes, 0, sizeof(modes));
  bool any_new = false;
  if (channel_mode_host) {
    for (int c = 0; c < K; ++c) {
      const int m = channel_mode_host[c];
      RMNET_CHECK_ARG(m == RMNET_CH_KEEP || m == RMNET_CH_ABSENT || m == RMNET_CH_NEW, "bad channel mode %d for channel %d", m, c);
      modes.set(c, m);
      any_new |= m == RMNET_CH_NEW;
    }
  }
  RMNET_CHECK_ARG(!any_new || new_mask, "a channel is marked RMNET_CH_NEW but new_mask is NULL");
  const int Hp = H + pad_t + pad_b, Wp = W + pad_l + pad_r;
  const long long n_pixels = (long long)H * W;
  dim3 grid((unsigned)((n_pixels + kThreads - 1) / kThreads));
  cudaStream_t st = (cudaStream_t)stream;
  bool hi_modes = false;
  for (int c = n_obj + 1; c < K; ++c) hi_modes |= modes.get(c) != RMNET_CH_KEEP;
#define RMNET_LAUNCH_EPI(NO)                                                                                                   \
  do {                                                                                                                         \
    if (hi_modes)                                                                                                              \
      mask_epilogue_kernel<NO, true><<<grid, kThreads, 0, st>>>(dec_logits, n_obj, K, H, W, Hp, Wp, pad_l, pad_t, modes, new_mask, \
                                                                logit_out, est_mask);                                         \
    else                                                                                                                       \
      mask_epilogue_kernel<NO, false><<<grid, kThreads, 0, st>>>(dec_logits, n_obj, K, H, W, Hp, Wp, pad_l, pad_t, modes, new_mask, \
                                                                 logit_out, est_mask);                                        \
  } while (0)
  if (n_obj <= 5) RMNET_LAUNCH_EPI(5);         // DAVIS-like clips
  else if (n_obj <= 11) RMNET_LAUNCH_EPI(11);  // YouTube-VOS-like clips (N_MAX_OBJECTS = 10)
  else if (n_obj <= 31) RMNET_LAUNCH_EPI(31);
  else RMNET_LAUNCH_EPI(63);
#undef RMNET_LAUNCH_EPI
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}
}
