// flow_affine_cpu.cpp -- update_optical_flow in plain host C++ (no CUDA call): the variant the NumPy-facing drop-in
// uses where the reference itself runs this op -- forked DataLoader worker processes (utils/data_transforms.py:293-302),
// which cannot create a CUDA context.  Same arithmetic as flow_affine.cu / flow_affine_transformation.cpp:63-83:
// float32, (a*j + b*i) + c left to right, every product and sum separately rounded (this file is compiled with
// -ffp-contract=off), roundf = half away from zero, the y1 update reads the already-updated x1 (:72-73).
#include <math.h>
#include <stddef.h>

#include "../../include/rmnet_b200.h"

namespace rmnet {
void set_error(const char *fmt, ...);
}

static inline float clampf(float v, float dim) { return v < 0 ? 0 : (v >= dim ? dim - 1 : v); }  // :75-78

extern "C" int rmnet_update_optical_flow_cpu(const float *of_host, const float *m1_host, const float *m2_host, int H, int W,
                                             float *out_host) {
  if (!of_host || !m1_host || !m2_host || !out_host) {
    rmnet::set_error("rmnet_update_optical_flow_cpu: null pointer argument");
    return RMNET_E_INVALID;
  }
  if (H <= 0 || W <= 0 || H >= (1 << 24) || W >= (1 << 24)) {
    rmnet::set_error("rmnet_update_optical_flow_cpu: bad shape H=%d W=%d", H, W);
    return RMNET_E_INVALID;
  }
  const float *a = m1_host, *b = m2_host;
  const float fw = (float)W, fh = (float)H;
  for (int i = 0; i < H; ++i) {
    const float fi = (float)i;
    const float *src = of_host + (size_t)i * W * 2;
    float *dst = out_host + (size_t)i * W * 2;
    for (int j = 0; j < W; ++j) {
      const float fj = (float)j;
      float x2 = roundf(b[0] * fj + b[1] * fi + b[2]);  // :68
      float y2 = roundf(b[3] * fj + b[4] * fi + b[5]);  // :69
      float x1 = fj + src[2 * j], y1 = fi + src[2 * j + 1];  // :71
      x1 = roundf(a[0] * x1 + a[1] * y1 + a[2]);  // :72
      y1 = roundf(a[3] * x1 + a[4] * y1 + a[5]);  // :73 (updated x1)
      x1 = clampf(x1, fw);
      y1 = clampf(y1, fh);
      x2 = clampf(x2, fw);
      y2 = clampf(y2, fh);
      dst[2 * j] = x1 - x2;  // :80-81
      dst[2 * j + 1] = y1 - y2;
    }
  }
  return RMNET_OK;
}
