// mma_chain_probe.cu -- microbenchmark (development aid): cycles per tcgen05.mma (kind::f16, M=128, K=16, A from TMEM, B from
// smem) for the issue patterns of the read kernel: dependent chains into one accumulator vs interleaved accumulators.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mma_chain_probe tools/probes/mma_chain_probe.cu && gpurun_out/mma_chain_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }  // fp16 x fp16 -> f32
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// TMEM columns: O [0,256), S0 [256,320), S1 [320,384), S2 [384,448), Q [448,512)
#define TMEM_LD16(addr, r)                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])            \
               : "r"(addr))
#define TMEM_ST16(addr, r)                                                                                         \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])                   \
               : "memory")
// threads 128..255 (4 warps = the 4 TMEM lane quarters) imitate the softmax warpgroup: per "tile" they load 64 columns of
// S2 and store 64 columns back, `bg` times, while thread 0 issues MMAs (patterns 8..9 = patterns 3 and 5 with that traffic)
__global__ void __launch_bounds__(256, 1) probe(long long *out, int reps) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(raw + (base - smem_u32(raw)))[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const uint32_t O = tm, S0 = tm + 256, S1 = tm + 320, S2 = tm + 384, Q = tm + 448;
  const uint64_t bK = desc_sw128(base), bV = desc_sw128(base + 32768), aQ = desc_sw128(base + 131072);
  const uint32_t i64 = idesc(64), i128 = idesc(128), i256 = idesc(256);
  int phase = 0;
  __shared__ volatile int go, stop;
  if (threadIdx.x == 0) { go = 0; stop = 0; }
  __syncthreads();
  if (threadIdx.x >= 128) {
    const uint32_t lane_base = (uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16;
    while (!go) {}
    uint32_t r[16];
    long long n = 0;
    while (!stop) {
      for (int c = 0; c < 64; c += 16) { TMEM_LD16(S2 + lane_base + c, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
      for (int c = 0; c < 64; c += 16) { TMEM_ST16(S2 + lane_base + c, r); }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ++n;
    }
    if ((threadIdx.x & 31) == 0 && threadIdx.x == 128) out[148 * 10 + blockIdx.x] = n;
  }
  if (threadIdx.x == 0) {
    for (int pat = 0; pat < 10; ++pat) {
      if (pat == 8) go = 1;
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        switch (pat) {
          case 0:  // 24 dependent N=64 MMAs into one accumulator (today's score product)
            for (int k = 0; k < 24; ++k) umma_ts(S0, Q + (k & 7) * 8, bK, i64, k > 0);
            break;
          case 1:  // 24 N=64 MMAs alternating between two accumulators
            for (int k = 0; k < 24; ++k) umma_ts((k & 1) ? S1 : S0, Q + (k & 7) * 8, bK, i64, k > 1);
            break;
          case 2:  // 12 dependent N=256 MMAs (today's P.V product)
            for (int k = 0; k < 12; ++k) umma_ts(O, S1 + (k & 3) * 8, bV, i256, 1);
            break;
          case 8:
          case 3:  // today's tile: 12 x N=256 then 24 x N=64
            for (int k = 0; k < 12; ++k) umma_ts(O, S1 + (k & 3) * 8, bV, i256, 1);
            for (int k = 0; k < 24; ++k) umma_ts(S0, Q + (k & 7) * 8, bK, i64, k > 0);
            break;
          case 4:  // interleaved: N=64 (score chain) / N=128 (half of a P.V MMA), 24 + 24
            for (int k = 0; k < 24; ++k) {
              umma_ts(S0, Q + (k & 7) * 8, bK, i64, k > 0);
              umma_ts(O + (k & 1) * 128, S1 + ((k >> 1) & 3) * 8, bV, i128, 1);
            }
            break;
          case 9:
          case 5:  // interleaved: two N=64 then one N=256
            for (int k = 0; k < 12; ++k) {
              umma_ts(S0, Q + ((2 * k) & 7) * 8, bK, i64, k > 0);
              umma_ts(S0, Q + ((2 * k + 1) & 7) * 8, bK, i64, 1);
              umma_ts(O, S1 + (k & 3) * 8, bV, i256, 1);
            }
            break;
          case 6:  // as 4, but every third score MMA takes A from shared memory (Q lo plane in smem)
            for (int k = 0; k < 24; ++k) {
              if (k % 3 == 2) umma_ss(S0, aQ, bK, i64, 1); else umma_ts(S0, Q + (k & 7) * 8, bK, i64, k > 0);
              umma_ts(O + (k & 1) * 128, S1 + ((k >> 1) & 3) * 8, bV, i128, 1);
            }
            break;
          case 7:  // 24 x N=128 dependent (two key tiles per score MMA)
            for (int k = 0; k < 24; ++k) umma_ts(S0, Q + (k & 7) * 8, bK, i128, k > 0);
            break;
        }
      }
      commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      long long t1 = clock64();
      out[blockIdx.x * 10 + pat] = t1 - t0;
    }
    stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  const int reps = 200, grid = 148;
  long long *d, h[148 * 11];
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int it = 0; it < 2; ++it) {
    probe<<<grid, 256, 200 * 1024>>>(d, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char *names[10] = {"24 x N64 dependent (score product today)", "24 x N64, two accumulators", "12 x N256 dependent (P.V today)",
                          "today's tile: 12 x N256 then 24 x N64", "interleaved 24 x (N64 chain, N128)", "interleaved 12 x (N64, N64, N256)",
                          "interleaved as above, every 3rd score MMA with A from smem", "24 x N128 dependent",
                          "today's tile + 4 warps of TMEM ld/st traffic", "interleaved (N64, N64, N256) + 4 warps of TMEM ld/st traffic"};
  const double ideal[10] = {24 * 32, 24 * 32, 12 * 128, 12 * 128 + 24 * 32, 24 * 32 + 24 * 64, 24 * 32 + 12 * 128, 24 * 32 + 24 * 64, 24 * 64, 2304, 2304};
  for (int p = 0; p < 10; ++p) {
    double s = 0;
    for (int b = 0; b < grid; ++b) s += (double)h[b * 10 + p];
    s /= grid * reps;
    printf("%-62s %8.1f cycles per group (floor %6.0f, x%.2f)\n", names[p], s, ideal[p], s / ideal[p]);
  }
  double n = 0, cyc = 0;
  for (int b = 0; b < grid; ++b) { n += (double)h[148 * 10 + b]; cyc += (double)h[b * 10 + 8] + (double)h[b * 10 + 9]; }
  printf("background warps: %.1f ld+st rounds of 64 columns per CTA in %.0f cycles = one round per %.0f cycles\n", n / grid, cyc / grid, cyc / n);
  return 0;
}
