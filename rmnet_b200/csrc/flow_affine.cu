// flow_affine.cu -- update_optical_flow on the GPU, bit-exact with the reference's scalar C++ loop
// (extensions/flow_affine_transformation/flow_affine_transformation.cpp:63-83).
// HBM-bound elementwise kernel: 16 B/pixel (8 read + 8 written), float2 accesses, one pixel per thread.
// Bit-exactness: the reference is built -O2 without -march (no FMA contraction), so every multiply/add is a
// separately rounded fp32 op -> __fmul_rn / __fadd_rn here; std::round == roundf (half away from zero).
#include "common.cuh"

namespace rmnet {
namespace {
struct Affine { float m[6]; };

__global__ void __launch_bounds__(256)
flow_affine_kernel(const float2 *__restrict__ of, Affine m1, Affine m2, int H, int W, float2 *__restrict__ out) {
  const long long n = (long long)H * W;
  const float fw = (float)W, fh = (float)H;  // `x1 >= width` compares float against size_t converted to float
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / W), j = (int)(idx - (long long)i * W);
    const float fj = (float)j, fi = (float)i;  // size_t -> float (:68)
    // :68-69  (a*j + b*i) + c, left to right
    float x2 = roundf(__fadd_rn(__fadd_rn(__fmul_rn(m2.m[0], fj), __fmul_rn(m2.m[1], fi)), m2.m[2]));
    float y2 = roundf(__fadd_rn(__fadd_rn(__fmul_rn(m2.m[3], fj), __fmul_rn(m2.m[4], fi)), m2.m[5]));
    const float2 f = __ldg(of + idx);
    float x1 = __fadd_rn(fj, f.x);  // :71
    float y1 = __fadd_rn(fi, f.y);
    x1 = roundf(__fadd_rn(__fadd_rn(__fmul_rn(m1.m[0], x1), __fmul_rn(m1.m[1], y1)), m1.m[2]));  // :72
    y1 = roundf(__fadd_rn(__fadd_rn(__fmul_rn(m1.m[3], x1), __fmul_rn(m1.m[4], y1)), m1.m[5]));  // :73 uses the UPDATED x1
    x1 = x1 < 0 ? 0 : (x1 >= fw ? fw - 1 : x1);  // :75-78 (width - 1 is exact in fp32 for W < 2^24)
    y1 = y1 < 0 ? 0 : (y1 >= fh ? fh - 1 : y1);
    x2 = x2 < 0 ? 0 : (x2 >= fw ? fw - 1 : x2);
    y2 = y2 < 0 ? 0 : (y2 >= fh ? fh - 1 : y2);
    out[idx] = make_float2(__fsub_rn(x1, x2), __fsub_rn(y1, y2));  // :80-81
  }
}
}  // namespace
}  // namespace rmnet

using namespace rmnet;
extern "C" {

int rmnet_update_optical_flow(const float *of, const float *m1_host, const float *m2_host, int H, int W, float *out,
                              void *stream) {
  RMNET_CHECK_ARG(of && m1_host && m2_host && out, "null pointer argument");
  RMNET_CHECK_ARG(H > 0 && W > 0 && H < (1 << 24) && W < (1 << 24), "bad shape H=%d W=%d", H, W);
  RMNET_CHECK_ARG((uintptr_t)of % 8 == 0 && (uintptr_t)out % 8 == 0, "of/out must be 8-byte aligned");
  Affine a1, a2;
  memcpy(a1.m, m1_host, sizeof(a1.m));
  memcpy(a2.m, m2_host, sizeof(a2.m));
  long long n = (long long)H * W;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  flow_affine_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float2 *)of, a1, a2, H, W, (float2 *)out);
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

int rmnet_update_optical_flow_host(const float *of_host, const float *m1_host, const float *m2_host, int H, int W,
                                   float *out_host, void *dev_scratch, size_t dev_scratch_bytes, void *stream) {
  RMNET_CHECK_ARG(of_host && out_host && dev_scratch, "null pointer argument");
  RMNET_CHECK_ARG(H > 0 && W > 0, "bad shape");
  const size_t bytes = (size_t)H * W * 2 * sizeof(float);
  if (dev_scratch_bytes < 2 * bytes) {
    set_error("dev_scratch too small: %zu < %zu", dev_scratch_bytes, 2 * bytes);
    return RMNET_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float *d_in = (float *)dev_scratch, *d_out = (float *)((char *)dev_scratch + bytes);
  RMNET_CUDA(cudaMemcpyAsync(d_in, of_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = rmnet_update_optical_flow(d_in, m1_host, m2_host, H, W, d_out, stream);
  if (rc) return rc;
  RMNET_CUDA(cudaMemcpyAsync(out_host, d_out, bytes, cudaMemcpyDeviceToHost, st));
  RMNET_CUDA(cudaStreamSynchronize(st));
  return RMNET_OK;
}
}
