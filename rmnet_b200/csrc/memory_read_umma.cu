// memory_read_umma.cu -- the fused regional memory read on Blackwell tensor cores (sm_100a).
//
// A piece = (128 in-region queries) x (one object) x (one half of the 512 value channels) x (one KV chunk).  The
// persistent grid (one CTA per SM) walks the work plan that the launch before this one built on the device from the actual
// region sizes (sched.cuh: which chunk of which object on which SM, in which order); warp 3 stages the CTA's piece list.
// A piece streams its chunk of the object's region-compacted bank in tiles of 64 memory cells and keeps everything on chip:
//
//   warp 0 : TMA producer for key tiles   (cp.async.bulk.tensor, SWIZZLE_128B, 3-stage mbarrier ring)
//   warp 2 : TMA producer for value tiles (2-stage ring)
//   warp 1 : MMA issuer -- one elected thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM):
//              S[128 q x 64 m]   = Q . K^T      A = Q  from TMEM, B = K tile (smem, K-major)
//              O[128 q x 256 cv] += P . V^T     A = P  from TMEM, B = V tile (smem, K-major)
//            in strict mode every product is the 3-term hi/lo split  Ah.Bh + Ah.Bl + Al.Bh  
//   warps 4-7 : softmax warpgroup, one thread per query row (TMEM lane): tcgen05.ld S -> scale -> lazy online
//            max -> exp2 -> hi/lo split of P -> tcgen05.st P over the S columns; rescales O in TMEM only when
//            the running max grows by more than 2^8; epilogue tcgen05.ld O -> coalesced partial stores.
//
// TMEM (512 columns x 128 lanes): O [0,256) | S/P buffer 0 [256,320) | S/P buffer 1 [320,384) | Q hi [384,448) | Q lo [448,512)
// Masked (never stored) memory cells and out-of-region queries are handled analytically by merge.cu.
//
// Reference math: models/rmnet.py:147-165 (MemoryReader.forward) with the regional masks of :245-248 / :355-358.
#include <cuda.h>

#include "common.cuh"
namespace rmnet {
RMNET_DEV_STAMPS(umma)
}
#include "sched.cuh"

namespace rmnet {
namespace {

constexpr int QT = 128;   // queries per CTA  (UMMA M)
constexpr int MT = 64;    // memory cells per KV tile
constexpr int CVH = 256;  // value channels per CTA (UMMA N of the P.V product)
constexpr int KST = 3;    // key-tile ring depth
constexpr int VST = 2;    // value-tile ring depth
constexpr int kThreads = 256;
constexpr float kTau = 8.0f;  // lazy-rescale threshold (log2 units): P stays <= 2^8

constexpr uint32_t K_PLANE_BYTES = MT * RMNET_CK * 2;       // 16 KB: [64 cells][128 ch] as two 8 KB SW128 blocks
constexpr uint32_t K_STAGE_BYTES = 2 * K_PLANE_BYTES;       // hi + lo
constexpr uint32_t V_PLANE_BYTES = CVH * MT * 2;            // 32 KB: [256 ch][64 cells]
constexpr uint32_t V_STAGE_BYTES = 2 * V_PLANE_BYTES;
constexpr uint32_t SMEM_TILES = KST * K_STAGE_BYTES + VST * V_STAGE_BYTES;  // 224 KB
constexpr uint32_t SMEM_BYTES = SMEM_TILES + 1024 /*align slack*/ + 256 /*barriers*/;

// TMEM column map
constexpr uint32_t TM_O = 0, TM_S0 = 256, TM_Q_HI = 384, TM_Q_LO = 448;

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address      bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset  bits [32,46): 8 rows x 128 B
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A/B = bf16 (fmt 0) or f16 (fmt 1), both K-major, M = 128.
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int n) {
  const uint32_t ab = (fmt == 0) ? 1u : 0u;  // F16F32Format: F16 = 0, BF16 = 1
  return (1u << 4) | (ab << 7) | (ab << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TMEM_LD16(addr, r, o)                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),     \
                 "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]),   \
                 "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])                                   \
               : "r"(addr))
#define TMEM_ST16(addr, r, o)                                                                                         \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(addr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), \
                 "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]),         \
                 "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15])                                       \
               : "memory")

#define TMEM_ST8(addr, r, o)                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"                   \
               ::"r"(addr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]),    \
                 "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])                                              \
               : "memory")

__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2, rel. error 2^-22; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Two fp32 -> one 32-bit word of 16-bit values (x0 in the low half) and the word of the residuals x - hi.
// One cvt.rn.{bf16x2,f16x2}.f32 per word; the residual subtraction is exact in fp32.
template <int FMT, bool LO>
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  if (FMT == 0) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<uint32_t *>(&h);
    if (LO) {
      const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
      __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
      lo = *reinterpret_cast<uint32_t *>(&l);
    }
  } else {
    __half2 h = __floats2half2_rn(x0, x1);
    hi = *reinterpret_cast<uint32_t *>(&h);
    if (LO) {
      const float2 f = __half22float2(h);
      __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
      lo = *reinterpret_cast<uint32_t *>(&l);
    }
  }
}

struct Barriers {
  uint64_t k_full[KST], k_empty[KST], v_full[VST], v_empty[VST];
  uint64_t s_full[2], p_full[2], pv_done[2], q_ready;
  uint32_t tmem_base;
};

struct Piece {
  int o, qtile, half, tile_begin, n_it, slot, count, q_rows;
};
// The pieces of this CTA in execution order (the plan of sched.cuh, written by the launch before this one); every warp
// role walks the same list.  The first PLAN_SMEM_PIECES entries are staged in shared memory.
struct PieceIter {
  const int4 *cached, *all;
  int n, k;
  __device__ PieceIter(const int4 *cached_, const int4 *all_, int n_) : cached(cached_), all(all_), n(n_), k(0) {}
  __device__ bool next(Piece &p) {
    if (k >= n) return false;
    const int4 v = k < PLAN_SMEM_PIECES ? cached[k] : ld_dep(all + k);
    ++k;
    p.o = v.x & 255;
    p.qtile = (v.x >> 8) & 255;
    p.half = (v.x >> 16) & 15;
    p.slot = (v.x >> 20) & 15;
    p.q_rows = (int)((unsigned)v.x >> 24) + 1;  // rows beyond are padding of the query tile: computed, never stored
    p.tile_begin = v.y;
    p.n_it = v.z;
    p.count = v.w;
    return true;
  }
};

// USE_LO: the score product Q.K^T takes the 3-term hi/lo split (K lo plane loaded, Q lo plane in TMEM);
// PV_LO : so does the P.V product (V lo plane loaded, P split into hi/lo).  (true, true) = strict, (false, false) = fast,
// (true, false) = mixed: scores -- whose error is exponentiated -- keep 22 mantissa bits, P and V go through one product.
template <int FMT, bool USE_LO, bool PV_LO>
__global__ void __launch_bounds__(kThreads, 1)
memory_read_umma_kernel(const __grid_constant__ CUtensorMap map_khi, const __grid_constant__ CUtensorMap map_klo,
                        const __grid_constant__ CUtensorMap map_vhi, const __grid_constant__ CUtensorMap map_vlo,
                        const uint16_t *__restrict__ qhi, const uint16_t *__restrict__ qlo,
                        const int2 *__restrict__ plan_hdr, const int4 *__restrict__ plan_pieces,
                        float *__restrict__ opart, float *__restrict__ ml, int nq_pad, int n_obj,
                        float *__restrict__ dbg_arg, int dbg_flags_arg) {
  // Development instrumentation (S dump, clock64 stamps, the no-TMA experiment) exists only in -DRMNET_DEV builds
  // (`make DEV=1`); in the release build the hooks are compile-time nulls and every branch on them is dead code.
#ifdef RMNET_DEV
  float *const dbg = dbg_arg;
  const int dbg_flags = dbg_flags_arg;
#else
  constexpr float *dbg = nullptr;
  constexpr int dbg_flags = 0;
  (void)dbg_arg; (void)dbg_flags_arg;
#endif
  DEV_STAMP_MIN(5);
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  unsigned char *smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  Barriers *bars = reinterpret_cast<Barriers *>(smem_al + SMEM_TILES);
  const uint32_t k_smem = smem_base, v_smem = smem_base + KST * K_STAGE_BYTES;
  __shared__ int4 s_pieces[PLAN_SMEM_PIECES];
  __shared__ int2 s_hdr;

#ifdef RMNET_DEV
  long long *tstamp = dbg ? reinterpret_cast<long long *>(dbg + 8448) + (size_t)blockIdx.x * 32 : nullptr;
#else
  constexpr long long *tstamp = nullptr;
#endif
  if (tstamp && threadIdx.x == 128) { tstamp[0] = clock64(); tstamp[16] = (long long)globaltimer_ns(); }
  constexpr int fmt = FMT;
  constexpr bool use_lo = USE_LO;
  constexpr bool pv_lo = PV_LO;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup: barriers (warp 0) and TMEM (warp 1) before the dependency wait -- in a chained launch (PDL)
  //      this overlaps the tail of the pack kernel -- then the CTA's piece list (warp 3).
  if (tstamp && threadIdx.x == 96) tstamp[8] = clock64();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_khi); tma_prefetch_desc(&map_klo); tma_prefetch_desc(&map_vhi); tma_prefetch_desc(&map_vlo);
    for (int i = 0; i < KST; ++i) { mbar_init(smem_u32(&bars->k_full[i]), 1); mbar_init(smem_u32(&bars->k_empty[i]), 1); }
    for (int i = 0; i < VST; ++i) { mbar_init(smem_u32(&bars->v_full[i]), 1); mbar_init(smem_u32(&bars->v_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->p_full[i]), 128);
      mbar_init(smem_u32(&bars->pv_done[i]), 1);
    }
    mbar_init(smem_u32(&bars->q_ready), 128);
    fence_barrier_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM: the smem footprint guarantees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tstamp && threadIdx.x == 32) tstamp[11] = clock64();
  pdl_wait();  // the plan, the packed queries and the temporary frame all come from the launch before this one
  const uint16_t *const qhi_w = pdl_launder(qhi), *const qlo_w = pdl_launder(qlo);  // (common.cuh: loads stay behind the wait)
  if (warp == 3) {
    const int2 hd = ld_dep(plan_hdr + blockIdx.x);  // (everything the pack kernel wrote: ld_dep, common.cuh)
    if (lane == 0) s_hdr = hd;
    for (int k = lane; k < min(hd.x, (int)PLAN_SMEM_PIECES); k += 32) s_pieces[k] = ld_dep(plan_pieces + hd.y + k);
  }
  if (tstamp && threadIdx.x == 96) tstamp[9] = clock64();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (s_hdr.x == 0) {  // no pieces for this CTA (uniform): give the TMEM back and leave
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(bars->tmem_base), "r"(512u) : "memory");
    return;
  }
  const uint32_t tmem = bars->tmem_base;
  PieceIter iter(s_pieces, plan_pieces + s_hdr.y, s_hdr.x);
  Piece pc;
  if (tstamp && threadIdx.x == 128) tstamp[1] = clock64();

  if (warp == 0) {
    // ================= key-tile TMA producer =================
    if (lane == 0) {
      int kt = 0;  // key tiles issued by this CTA so far (ring position / parity run across pieces)
      while (iter.next(pc)) {
        for (int it = 0; it < pc.n_it; ++it, ++kt) {
          const int s = kt % KST;
          mbar_wait(smem_u32(&bars->k_empty[s]), ((kt / KST) & 1) ^ 1);
          const uint32_t full = smem_u32(&bars->k_full[s]);
          if ((dbg_flags & 1) && kt >= KST) { mbar_arrive(full); continue; }  // dev experiment: no TMA traffic after the ring fill (results are garbage)
          mbar_expect_tx(full, use_lo ? K_STAGE_BYTES : K_PLANE_BYTES);
          const uint32_t dst = k_smem + s * K_STAGE_BYTES;
          const int m0 = (pc.tile_begin + it) * MT;
          tma_load_3d(dst, &map_khi, full, 0, m0, pc.o);
          tma_load_3d(dst + K_PLANE_BYTES / 2, &map_khi, full, 64, m0, pc.o);
          if (use_lo) {
            tma_load_3d(dst + K_PLANE_BYTES, &map_klo, full, 0, m0, pc.o);
            tma_load_3d(dst + K_PLANE_BYTES + K_PLANE_BYTES / 2, &map_klo, full, 64, m0, pc.o);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================= value-tile TMA producer =================
    if (lane == 0) {
      int vt = 0;
      // At kernel start all 148 CTAs fill their rings and fetch their Q rows at once: 42 MB through L2 -> SM at its
      // throughput cap, ~7 k cycles before the first score MMA could issue.  The first score tile needs Q and one key
      // stage only, so the value ring (128 of the 224 KB) waits until this CTA's Q rows are in TMEM; its first tile is
      // due a score MMA chain and a softmax pass (~2.3 k cycles) later.
      mbar_wait(smem_u32(&bars->q_ready), 0);
      while (iter.next(pc)) {
        for (int it = 0; it < pc.n_it; ++it, ++vt) {
          const int s = vt % VST;
          mbar_wait(smem_u32(&bars->v_empty[s]), ((vt / VST) & 1) ^ 1);
          const uint32_t full = smem_u32(&bars->v_full[s]);
          if ((dbg_flags & 1) && vt >= VST) { mbar_arrive(full); continue; }
          mbar_expect_tx(full, pv_lo ? V_STAGE_BYTES : V_PLANE_BYTES);
          const uint32_t dst = v_smem + s * V_STAGE_BYTES;
          const int m0 = (pc.tile_begin + it) * MT;
          tma_load_3d(dst, &map_vhi, full, m0, pc.half * CVH, pc.o);
          if (pv_lo) tma_load_3d(dst + V_PLANE_BYTES, &map_vlo, full, m0, pc.half * CVH, pc.o);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc_qk = umma_idesc(fmt, MT), idesc_pv = umma_idesc(fmt, CVH);
    int gt = 0;       // tiles whose score MMA has been issued (global over pieces: S/P buffer + key ring position)
    auto issue_qk = [&]() {
      const int s = gt % KST, b = gt & 1;
      mbar_wait(smem_u32(&bars->k_full[s]), (gt / KST) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t kb = k_smem + s * K_STAGE_BYTES;
        const uint32_t d = tmem + TM_S0 + b * MT;
#pragma unroll
        for (int kk = 0; kk < RMNET_CK / 16; ++kk) {
          // 16 channels = 32 B inside the 128 B swizzle row; channels 64..127 live in the second 8 KB block
          const uint32_t off = (kk >> 2) * (K_PLANE_BYTES / 2) + (kk & 3) * 32;
          const uint64_t bh = umma_desc_sw128(kb + off);
          umma_ts(d, tmem + TM_Q_HI + kk * 8, bh, idesc_qk, kk > 0 ? 1u : 0u);
          if (use_lo) {
            const uint64_t bl = umma_desc_sw128(kb + K_PLANE_BYTES + off);
            umma_ts(d, tmem + TM_Q_HI + kk * 8, bl, idesc_qk, 1u);
            umma_ts(d, tmem + TM_Q_LO + kk * 8, bh, idesc_qk, 1u);
          }
        }
        umma_commit(smem_u32(&bars->k_empty[s]));  // key stage free once these MMAs retire
        umma_commit(smem_u32(&bars->s_full[b]));   // scores ready for the softmax warpgroup
      }
      __syncwarp();
      ++gt;
    };
    int pt = 0;       // tiles whose P.V MMA has been issued (global)
    int n_piece = 0;
    while (iter.next(pc)) {
      mbar_wait(smem_u32(&bars->q_ready), n_piece & 1);  // this piece's Q rows are in TMEM
      tc_fence_after();
      ++n_piece;
      // software pipeline, one code site per product: QK(0) QK(1) | PV(0) QK(2) | PV(1) QK(3) | ... | PV(n-2) | PV(n-1)
      // QK(s) overwrites the S/P buffer that PV(s-2) just consumed: the tensor pipe executes in issue order.
#pragma unroll 1
      for (int st = 0; st < pc.n_it + 2; ++st) {
        if (st >= 2) {
          const int it = st - 2;
          const int s = pt % VST, b = pt & 1;
          mbar_wait(smem_u32(&bars->p_full[b]), (pt >> 1) & 1);
          mbar_wait(smem_u32(&bars->v_full[s]), (pt / VST) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t vb = v_smem + s * V_STAGE_BYTES;
            const uint32_t p_hi = tmem + TM_S0 + b * MT, p_lo = p_hi + MT / 2;
#pragma unroll
            for (int kk = 0; kk < MT / 16; ++kk) {
              const uint64_t bh = umma_desc_sw128(vb + kk * 32);
              umma_ts(tmem + TM_O, p_hi + kk * 8, bh, idesc_pv, (it > 0 || kk > 0) ? 1u : 0u);
              if (pv_lo) {
                const uint64_t bl = umma_desc_sw128(vb + V_PLANE_BYTES + kk * 32);
                umma_ts(tmem + TM_O, p_hi + kk * 8, bl, idesc_pv, 1u);
                umma_ts(tmem + TM_O, p_lo + kk * 8, bh, idesc_pv, 1u);
              }
            }
            umma_commit(smem_u32(&bars->v_empty[s]));
            umma_commit(smem_u32(&bars->pv_done[b]));
          }
          __syncwarp();
          ++pt;
        }
        if (st < pc.n_it) issue_qk();
      }
    }
  } else if (warp >= 4) {
    // ================= softmax / correction / epilogue warpgroup: thread <-> query row <-> TMEM lane =================
    const int row = threadIdx.x - 128;                  // 0..127
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_base = tmem + lane_base;
    const float scale = 1.4426950408889634f * rsqrtf((float)RMNET_CK);
    int gt = 0;  // global tile counter (S/P buffer + barrier parity), runs across pieces
    bool first_piece = true;
    pdl_trigger();  // (after the wait above: the merge kernel's pre-wait part relies on the bank being final)
    // this row's 128 query-key channels as packed 16-bit pairs (hi and lo planes), written by the pack kernel's
    // query role in the TMEM column order (column c = channels 2c, 2c+1), 32 rows interleaved per 16 B chunk so that
    // each of the loads below is one coalesced 512 B access per warp
    uint32_t qh[RMNET_CK / 2], ql[USE_LO ? RMNET_CK / 2 : 1];
    auto fetch_q = [&](const Piece &p) {
      const size_t r = ((size_t)p.o * nq_pad + p.qtile * QT + (row & ~31)) * (RMNET_CK / 8) + (row & 31);  // uint4 units
      const uint4 *ph = reinterpret_cast<const uint4 *>(qhi_w) + r, *pl = reinterpret_cast<const uint4 *>(qlo_w) + r;
#pragma unroll
      for (int j = 0; j < RMNET_CK / 8; ++j) {
        const uint4 v = __ldg(ph + j * 32);
        qh[4 * j] = v.x; qh[4 * j + 1] = v.y; qh[4 * j + 2] = v.z; qh[4 * j + 3] = v.w;
      }
      if (USE_LO) {
#pragma unroll
        for (int j = 0; j < RMNET_CK / 8; ++j) {
          const uint4 v = __ldg(pl + j * 32);
          ql[4 * j] = v.x; ql[4 * j + 1] = v.y; ql[4 * j + 2] = v.z; ql[4 * j + 3] = v.w;
        }
      }
    };
    // One code site per phase.  Order per piece k:  [Q(k) -> TMEM]  [drain O(k-1)]  [tiles of k]  -- so the tensor pipe
    // already computes S(k, 0..1) while the numerators of piece k-1 leave TMEM, and the prefetched Q registers are dead
    // before the drain needs its own.
    Piece nxt;
    bool have_cur = false, have_nxt = iter.next(nxt);
    int gt_done = 0;  // tile counter at the end of the current piece (parity of its last pv_done)
    if (tstamp && row == 0) tstamp[12] = clock64();
    for (;;) {
      if (have_nxt) {
        // ---- Q rows of the next piece -> 16-bit hi/lo planes in TMEM (A operand of the score product).  All score
        //      MMAs of the current piece have retired (its last softmax pass waited on their commit) and its P.V
        //      MMAs do not read Q.
        fetch_q(nxt);
        if (tstamp && first_piece && row == 0) tstamp[13] = clock64();
#pragma unroll
        for (int c = 0; c < RMNET_CK / 2; c += 16) {
          TMEM_ST16(t_base + TM_Q_HI + c, qh, c);
          if (USE_LO) TMEM_ST16(t_base + TM_Q_LO + c, ql, c);
        }
        if (tstamp && first_piece && row == 0) tstamp[14] = clock64();
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->q_ready));
        if (tstamp && first_piece && row == 0) { tstamp[2] = clock64(); tstamp[21] = (long long)globaltimer_ns(); }
      }
      if (have_cur) {
        // ---- epilogue of the current piece: unnormalised numerators for merge.cu, partial slot pc.slot
        const int n = pc.qtile * QT + row;
        mbar_wait(smem_u32(&bars->pv_done[(gt_done - 1) & 1]), ((gt_done - 1) >> 1) & 1);
        tc_fence_after();
        if (tstamp && first_piece && row == 0) tstamp[5] = clock64();
        float *ob = opart + (((size_t)pc.slot * n_obj + pc.o) * RMNET_CV + pc.half * CVH) * nq_pad + n;
        // software pipeline over 32-column groups: the TMEM load of group g+1 is in flight while group g is stored
        uint32_t ra[32], rb[32];
        TMEM_LD16(t_base + TM_O, ra, 0);
        TMEM_LD16(t_base + TM_O + 16, ra, 16);
#pragma unroll 1
        for (int c = 0; c < CVH; c += 64) {
          tc_wait_ld();
          TMEM_LD16(t_base + TM_O + c + 32, rb, 0);
          TMEM_LD16(t_base + TM_O + c + 48, rb, 16);
          if (row < pc.q_rows) {
#pragma unroll
            for (int j = 0; j < 32; ++j) ob[(c + j) * nq_pad] = __uint_as_float(ra[j]);  // lanes run along queries: coalesced
          }
          tc_wait_ld();
          if (c + 64 < CVH) {
            TMEM_LD16(t_base + TM_O + c + 64, ra, 0);
            TMEM_LD16(t_base + TM_O + c + 80, ra, 16);
          }
          if (row < pc.q_rows) {
#pragma unroll
            for (int j = 0; j < 32; ++j) ob[(c + 32 + j) * nq_pad] = __uint_as_float(rb[j]);
          }
        }
        if (tstamp && first_piece && row == 0) tstamp[6] = clock64();
        first_piece = false;
      }
      if (!have_nxt) break;
      pc = nxt;
      have_cur = true;
      const int o = pc.o;
      const int n = pc.qtile * QT + row;
      const int count = pc.count;

      float m_ref = -INFINITY, l_sum = 0.f;
      for (int it = 0; it < pc.n_it; ++it, ++gt) {
        const int b = gt & 1;
        const uint32_t s_addr = t_base + TM_S0 + b * MT;
        mbar_wait(smem_u32(&bars->s_full[b]), (gt >> 1) & 1);
        tc_fence_after();
        uint32_t sr[MT];
        TMEM_LD16(s_addr, sr, 0);
        TMEM_LD16(s_addr + 16, sr, 16);
        TMEM_LD16(s_addr + 32, sr, 32);
        TMEM_LD16(s_addr + 48, sr, 48);
        tc_wait_ld();
        if (tstamp && first_piece && it == 0 && row == 0) tstamp[3] = clock64();
        long long t_s1 = 0;
        if (tstamp && first_piece && it == 1 && row == 0) t_s1 = clock64();  // dev: S of the second tile is in registers
        if (dbg && first_piece && it == 0 && blockIdx.x == 0) {
#pragma unroll
          for (int j = 0; j < MT; ++j) dbg[row * MT + j] = __uint_as_float(sr[j]);
        }
        const int valid = count - (pc.tile_begin + it) * MT;  // columns >= valid are beyond the stored cells (last tile only)
        if (valid < MT) {
#pragma unroll
          for (int j = 0; j < MT; ++j) if (j >= valid) sr[j] = 0xff800000u;  // -inf
        }
        float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]);
#pragma unroll
        for (int j = 2; j < MT; j += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[j]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[j + 1]));
        }
        const float mx = fmaxf(mx0, mx1) * scale;  // scale > 0: max commutes with the scaling
        if (it == 0) {
          m_ref = (mx == -INFINITY) ? 0.f : mx;
        } else if (__any_sync(0xffffffffu, mx > m_ref + kTau)) {
          // lazy rescale of the O accumulator (rare after the first tiles): PV(it-1) must have retired, PV(it) cannot
          // start before this warp arrives on p_full below.
          mbar_wait(smem_u32(&bars->pv_done[(gt - 1) & 1]), ((gt - 1) >> 1) & 1);
          tc_fence_after();
          const float m_new = fmaxf(m_ref, mx);
          const float f = exp2f(m_ref - m_new);
#pragma unroll 1
          for (int c = 0; c < CVH; c += 32) {
            uint32_t orr[32];
            TMEM_LD16(t_base + TM_O + c, orr, 0);
            TMEM_LD16(t_base + TM_O + c + 16, orr, 16);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * f);
            TMEM_ST16(t_base + TM_O + c, orr, 0);
            TMEM_ST16(t_base + TM_O + c + 16, orr, 16);
          }
          l_sum *= f;
          m_ref = m_new;
        }
        // P = 2^(s*scale - m_ref): one FFMA + one MUFU per element, then the 16-bit hi/lo split (packed cvt), stored
        // two cells per TMEM column over the S buffer: hi plane in columns [0,32), lo plane in [32,64)
        const float neg_m = -m_ref;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < MT; c += 16) {
          uint32_t ph[8], pl[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c + 2 * j]), scale, neg_m));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c + 2 * j + 1]), scale, neg_m));
            l0 += p0;
            l1 += p1;
            pl[j] = 0;
            split_pack2<FMT, PV_LO>(p0, p1, ph[j], pl[j]);
          }
          TMEM_ST8(s_addr + c / 2, ph, 0);
          if (PV_LO) TMEM_ST8(s_addr + MT / 2 + c / 2, pl, 0);
        }
        l_sum += l0 + l1;
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->p_full[b]));
        if (tstamp && first_piece && it == 1 && row == 0) tstamp[15] = clock64();
        if (tstamp && first_piece && it == 1 && row == 0) tstamp[14] = t_s1;
      }
      gt_done = gt;
      // (max, sum) statistics of the piece for merge.cu
      {
        float2 *dst = reinterpret_cast<float2 *>(ml) + (((size_t)pc.slot * n_obj + o) * 2 + pc.half) * nq_pad + n;
        *dst = make_float2(m_ref, l_sum);
      }
      if (dbg && first_piece && blockIdx.x == 0) dbg[QT * MT + row] = m_ref;
      if (tstamp && first_piece && row == 0) { tstamp[4] = clock64(); tstamp[7] = pc.n_it; }
      have_nxt = iter.next(nxt);
      if (tstamp && row == 0) { tstamp[19] += 1; tstamp[22] += pc.n_it; if (tstamp[19] == 1) tstamp[18] = pc.o | (pc.qtile << 8) | (pc.half << 16) | (pc.slot << 20); }
    }
    if (tstamp && row == 0) { tstamp[20] = clock64(); tstamp[17] = (long long)globaltimer_ns(); }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  DEV_STAMP_MAX(6);
}

// ---- host: TMA descriptors ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;  // resolved once; benign race (idempotent)
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_map(CUtensorMap *m, void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
             uint64_t stride2_bytes, uint32_t b0, uint32_t b1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return RMNET_E_CUDA; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return RMNET_E_CUDA; }
  return RMNET_OK;
}

thread_local float *g_dbg = nullptr;
thread_local int g_dbg_flags = 0;

}  // namespace

bool umma_supported(int cap_cells) { return cap_cells % 64 == 0; }

// persistent grid: one CTA per SM (224 KB of dynamic smem => 1 CTA/SM); merge.cu rebuilds the same schedule from it
int umma_grid_size() {
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n_sms = v;
    else n_sms = 148;
  }
  return n_sms;
}

// The work plan (W.plan_hdr / W.plan_pieces / W.sched) must have been written by the launch right before this one in the
// stream: the pack kernel of rmnet_frame_step or the query-side launch of rmnet_bank_memory_read (bank.cu: ROLE_PLAN).
int launch_memory_read_umma(const BankView &bank, int n_obj, int fmt, int precision, const ReadWorkspace &W, bool pdl,
                            cudaStream_t st) {
  RMNET_CHECK_ARG(bank.cap % 64 == 0, "tcgen05 path needs cap_cells %% 64 == 0 (got %d)", bank.cap);
  RMNET_CHECK_ARG(n_obj <= SCHED_MAX_OBJ, "tcgen05 path supports at most %d objects per call", (int)SCHED_MAX_OBJ);
  // The four tensor maps depend only on the bank (base pointers, capacity, slots): encode once per bank and thread.
  struct MapCache { const void *khi; int cap, n_slots; CUtensorMap m[4]; };
  static thread_local MapCache cache[4] = {};
  static thread_local int cache_next = 0;
  const MapCache *mc = nullptr;
  for (int i = 0; i < 4; ++i)
    if (cache[i].khi == bank.khi && cache[i].cap == bank.cap && cache[i].n_slots == bank.n_slots) mc = &cache[i];
  if (!mc) {
    MapCache &c = cache[cache_next];
    cache_next = (cache_next + 1) & 3;
    c.khi = nullptr;
    const uint64_t cap = bank.cap, ns = bank.n_slots;
    int rc;
    // keys  [slot][cell][128 ch] : box 64 ch x 64 cells (128 B rows)      values [slot][512 ch][cell] : box 64 cells x 256 ch
    if ((rc = make_map(&c.m[0], bank.khi, RMNET_CK, cap, ns, RMNET_CK * 2, cap * RMNET_CK * 2, 64, MT))) return rc;
    if ((rc = make_map(&c.m[1], bank.klo, RMNET_CK, cap, ns, RMNET_CK * 2, cap * RMNET_CK * 2, 64, MT))) return rc;
    if ((rc = make_map(&c.m[2], bank.vhi, cap, RMNET_CV, ns, cap * 2, cap * RMNET_CV * 2, MT, CVH))) return rc;
    if ((rc = make_map(&c.m[3], bank.vlo, cap, RMNET_CV, ns, cap * 2, cap * RMNET_CV * 2, MT, CVH))) return rc;
    c.khi = bank.khi; c.cap = bank.cap; c.n_slots = bank.n_slots;
    mc = &c;
  }
  const CUtensorMap &mkh = mc->m[0], &mkl = mc->m[1], &mvh = mc->m[2], &mvl = mc->m[3];
  const int n_sms = umma_grid_size();
  dim3 grid(n_sms);
  RMNET_CHECK_ARG(n_sms <= PLAN_HDR_CTAS && W.nq_pad / 128 <= 256, "tcgen05 path: %d SMs / %d query tiles exceed the plan's limits", n_sms, W.nq_pad / 128);
  const bool lo = precision != RMNET_PREC_SINGLE, plo = precision == RMNET_PREC_SPLIT3;
#define RMNET_LAUNCH_UMMA(F, L, P)                                                                                          \
  do {                                                                                                                   \
    static bool attr_set[64] = {};                                                                                       \
    int dev_ = 0;                                                                                                        \
    RMNET_CUDA(cudaGetDevice(&dev_));                                                                                    \
    if (dev_ < 0 || dev_ >= 64 || !attr_set[dev_]) {                                                                     \
      RMNET_CUDA(cudaFuncSetAttribute(memory_read_umma_kernel<F, L, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                      (int)SMEM_BYTES));                                                                 \
      if (dev_ >= 0 && dev_ < 64) attr_set[dev_] = true;                                                                 \
    }                                                                                                                    \
    RMNET_CUDA(launch_kernel(memory_read_umma_kernel<F, L, P>, grid, dim3(kThreads), SMEM_BYTES, st, pdl, mkh, mkl, mvh, mvl, \
                             W.qhi, W.qlo, reinterpret_cast<const int2 *>(W.plan_hdr),                                  \
                             reinterpret_cast<const int4 *>(W.plan_pieces), W.opart, W.ml, W.nq_pad, n_obj, g_dbg,      \
                             g_dbg_flags));                                                                              \
  } while (0)
  if (fmt == 0 && plo) RMNET_LAUNCH_UMMA(0, true, true);
  else if (fmt == 0 && lo) RMNET_LAUNCH_UMMA(0, true, false);
  else if (fmt == 0) RMNET_LAUNCH_UMMA(0, false, false);
  else if (plo) RMNET_LAUNCH_UMMA(1, true, true);
  else if (lo) RMNET_LAUNCH_UMMA(1, true, false);
  else RMNET_LAUNCH_UMMA(1, false, false);
#undef RMNET_LAUNCH_UMMA
  RMNET_LAUNCH_CHECK();
  return RMNET_OK;
}

}  // namespace rmnet

#ifdef RMNET_DEV
// development hooks, `make DEV=1` builds only (never in the release library, not part of the public header):
// dump S of the first tile of CTA (0,0,0) into `ptr` (128*64 + 128 floats)
extern "C" __attribute__((visibility("default"))) void rmnet_debug_set_umma_dump(float *ptr) { rmnet::g_dbg = ptr; }
// development hook: bit 0 = the TMA producers stop loading after the first ring fill (timing experiment only: results are garbage)
extern "C" __attribute__((visibility("default"))) void rmnet_debug_set_umma_flags(int flags) { rmnet::g_dbg_flags = flags; }
// chain stamps (common.cuh RMNET_DEV_STAMPS): a device array of 16 u64 (even slots = min-initialised starts, see tools/chain_timeline.py)
namespace rmnet {
void dev_set_chain_stamps_att_map(unsigned long long *);
void dev_set_chain_stamps_bank(unsigned long long *);
void dev_set_chain_stamps_merge(unsigned long long *);
}
extern "C" __attribute__((visibility("default"))) void rmnet_debug_set_chain_stamps(unsigned long long *p) {
  rmnet::dev_set_chain_stamps_att_map(p); rmnet::dev_set_chain_stamps_bank(p); rmnet::dev_set_chain_stamps_merge(p); rmnet::dev_set_chain_stamps_umma(p);
}
#endif  // RMNET_DEV
