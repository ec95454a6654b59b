// sched.cuh -- the work plan of the persistent tcgen05 read kernel (memory_read_umma.cu), built ON THE DEVICE from the
// actual cell counts by one CTA of the pack / query-side launch that precedes every read (bank.cu: ROLE_PLAN).
//
// A piece = (object o, query tile qt, Cv half, KV chunk [t0, t0 + len) of the object's tiles, partial slot j).  All
// 2 * nqt(o) pieces of one (o, chunk) -- a "group" -- go to CTAs that run side by side, so the group streams the SAME
// key / value tiles (one DRAM fetch, L2 hits for the rest).  The plan decides the chunking of every object and which
// CTA runs which pieces in which order; the read kernel just walks its list.  Two planners, the cheaper estimate wins:
//
//   * deal  : one chunk length c for all objects, ns(o) = ceil(nt(o) / c) balanced chunks, pieces dealt round-robin in
//             (o, chunk, half, qt) order.  Good when every CTA gets several pieces anyway (large banks / many objects).
//   * fill  : water-filling.  Objects in order of decreasing group width; every chunk takes the least-loaded CTAs and
//             is made as long as fits under a target level Lv = ideal makespan + margin (<= 64 tiles: accumulation-chain
//             bound, common.cuh), so CTAs end up level although group widths (2..2*nqt) and object sizes differ; the
//             short remainders of one object fill the room that the others left.  Three margins are tried in parallel
//             (one warp each).  At the headline size (480p, 5 objects, T = 20) the longest CTA drops from 43 tiles in 1
//             piece to 37 tiles in 1-2 pieces (kernel 88 -> 75 us).  tests/plan_model.py restates both planners in plain
//             Python; the GPU tests hold the device-built plan equal to it piece by piece (tests/test_gpu_plan.py).
//
// Cost model (cycles, measured with the DEV stamps: tools/umma_timeline.py): a KV tile, the first piece's prologue +
// drain, every further piece's restart.
#pragma once
#include "common.cuh"

namespace rmnet {

enum {
  PLAN_MAX_RECORDS = 160,    // group records of one water-filling run
  PLAN_MAX_SEGS = 64,        // CTA ranges of equal load
  PLAN_MAX_STEPS = 72,       // placement steps of one water-filling run (bounds its latency: ~1 200 cycles each on a shared SM)
  PLAN_FILL_STRIDE = 16,     // piece-list capacity per CTA of a water-filling plan
  PLAN_MIN_CHUNK = 4,        // tiles
  PLAN_FILL_MAX_LOAD = 96,   // water-filling only when the ideal load is at most this many tiles per CTA ...
  PLAN_FILL_MIN_LOAD = 12,   // ... and at least this many (below, the pieces are a few tiles long and dealing them is as good)
  PLAN_N_MARGINS = 3,
  PLAN_SMEM_PIECES = 64      // pieces of a CTA cached in the read kernel's shared memory
};

struct PlanCost { int tile, first, extra; };
__host__ __device__ inline PlanCost plan_cost(int precision) {
  PlanCost c;
  c.tile = precision == RMNET_PREC_SPLIT3 ? 2700 : (precision == RMNET_PREC_MIXED ? 1740 : 1490);
  c.first = 14000;
  c.extra = 6000;
  return c;
}

#ifdef __CUDACC__
// A record = one chunk [t0, t0 + len) of object o for the CTAs [cta_begin, cta_begin + n_ctas): CTA cta_begin + d runs
// unit unit0 + d % gw (unit = half * nqt + query tile) of chunk number d / gw, i.e. tiles [t0 + (d / gw) * len, ...) into
// partial slot slot + d / gw.  gw = n_ctas for a single chunk (possibly one part of a group that was split over two CTA
// ranges: unit0 > 0); gw = the group width when the record stands for several identical chunks placed side by side.
struct PlanRecord { int o_slot, t0, len, cta_begin, n_ctas, unit0, gw, pad; };  // o | slot << 8

struct PlanSmem {
  int nt[SCHED_MAX_OBJ], nqt[SCHED_MAX_OBJ], count[SCHED_MAX_OBJ], qcells[SCHED_MAX_OBJ];
  int ns[PLAN_N_MARGINS + 1][SCHED_MAX_OBJ];  // [0] = deal, [1 + m] = fill with margin m
  int ibase[SCHED_MAX_OBJ + 1];               // deal: first item of object o
  alignas(16) PlanRecord rec[PLAN_N_MARGINS][PLAN_MAX_RECORDS];
  int nrec[PLAN_N_MARGINS];
  unsigned cost[PLAN_N_MARGINS + 1];          // estimated makespan in cycles (0xffffffff = not applicable)
  int deal_items, winner;
};

// ceil(a / b) for 0 < b, a < 2^20 via one float multiply and a fix-up (a 32-bit integer division costs ~25 instructions)
__device__ __forceinline__ unsigned plan_ceil_div(unsigned a, unsigned b, float rcp_b) {
  unsigned q = (unsigned)((float)a * rcp_b);
  q += (q * b < a) ? 1u : 0u;
  q += (q * b < a) ? 1u : 0u;
  q -= (q > 0 && (q - 1u) * b >= a) ? 1u : 0u;
  return q;
}

// floor(a / b) for 0 <= a < 2^20, 0 < b < 2^20 (one float multiply and two fix-ups)
__device__ __forceinline__ int plan_floor_div(int a, int b) {
  int q = (int)((float)a * __frcp_rn((float)b));
  q -= (q * b > a) ? 1 : 0;
  q += ((q + 1) * b <= a) ? 1 : 0;
  return q;
}

// ---- planner "deal" (one full warp): chunk length c minimising rounds(c) * (longest chunk + overhead) ---------------
__device__ __forceinline__ void plan_deal(PlanSmem &S, int n_obj, int G, PlanCost C) {
  const int lane = threadIdx.x & 31;
  int max_nt = 0;
  for (int o = lane; o < n_obj; o += 32) max_nt = max(max_nt, S.nqt[o] > 0 ? S.nt[o] : 0);
  max_nt = __reduce_max_sync(0xffffffffu, max_nt);
  const unsigned c_min = max(1u, ((unsigned)max_nt + READ_MAX_SPLITS - 1) / READ_MAX_SPLITS);
  const unsigned cand[2] = {(unsigned)lane + 1u, (unsigned)lane + 33u};
  const float rcp[2] = {__frcp_rn((float)cand[0]), __frcp_rn((float)cand[1])};
  unsigned items[2] = {0u, 0u}, longest[2] = {0u, 0u};
#pragma unroll 2
  for (int o = 0; o < n_obj; ++o) {
    const unsigned nt = S.nt[o], w2 = 2u * (unsigned)S.nqt[o];
    if (nt > 0 && w2 > 0) {
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const unsigned ns = plan_ceil_div(nt, cand[rep], rcp[rep]);
        items[rep] += ns * w2;
        longest[rep] = max(longest[rep], plan_ceil_div(nt, ns, __frcp_rn((float)ns)));
      }
    }
  }
  unsigned best = 0xffffffffu, best_c = c_min;
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const unsigned c = cand[rep];
    if (c >= c_min && c <= MAX_TILES_PER_SPLIT && c <= (unsigned)max(max_nt, 1)) {
      const unsigned rounds = (items[rep] + G - 1) / (unsigned)G;
      // makespan bound in units of 64 cycles (fits 26 bits below the tie-break); ties -> fewer chunks
      const unsigned cyc = rounds * longest[rep] * (unsigned)C.tile + (unsigned)C.first + (rounds - 1u) * (unsigned)C.extra;
      const unsigned cost = min(cyc >> 6, 0x3ffffffu) * 64u + (64u - c);
      if (cost < best) { best = cost; best_c = c; }
    }
  }
  const unsigned bcast = __reduce_min_sync(0xffffffffu, best);
  const unsigned who = __ballot_sync(0xffffffffu, best == bcast);
  int c = (int)__shfl_sync(0xffffffffu, best_c, __ffs(who) - 1);
  if (max_nt > MAX_TILES_PER_SPLIT * READ_MAX_SPLITS) c = (int)c_min;  // huge banks: the slot bound wins over the chain bound
  if (lane == 0) {
    int acc = 0;
    for (int o = 0; o < n_obj; ++o) {
      const int nt = S.nt[o];
      // (with a huge object in the bank c exceeds the chain bound: the other objects keep theirs as far as their 16 slots go)
      const int ns = (nt > 0 && S.nqt[o] > 0) ? max((nt + c - 1) / c, min((int)READ_MAX_SPLITS, (nt + MAX_TILES_PER_SPLIT - 1) / MAX_TILES_PER_SPLIT)) : 0;
      S.ns[0][o] = ns;
      S.ibase[o] = acc;
      acc += ns * 2 * S.nqt[o];
    }
    S.ibase[n_obj] = acc;
    S.deal_items = acc;
    S.cost[0] = bcast == 0xffffffffu ? 0xfffffffeu : (bcast >> 6) * 64u;  // always applicable
#ifdef RMNET_DEV
    if (s_chain_stamps) { atomicMax(s_chain_stamps + 15, dev_globaltimer()); atomicMax(s_chain_stamps + 19, (unsigned long long)S.cost[0]); }
#endif
  }
  __syncwarp();
}

// ---- planner "fill" (one full warp per margin m): water-filling, see the file comment ------------------------------
// All control flow is warp-uniform.  The segment list (CTA ranges of equal load) lives in REGISTERS, two segments per lane
// (index lane and lane + 32): load and begin | len << 10 | pieces << 20; a step is one REDUX arg-min, a few shuffles and
// integer ALU work (~200 cycles) -- a shared-memory version of the same loop took ~1 200 cycles per step and made the
// plan role the last CTA of the pack kernel by 7 us.
__device__ __forceinline__ void plan_fill(PlanSmem &S, int m, int n_obj, int G, PlanCost C, long long total_units) {
  const int lane = threadIdx.x & 31;
  PlanRecord *rec = S.rec[m];
  unsigned cost = 0xffffffffu;
  int nrec = 0;
#ifdef RMNET_DEV
  const long long dev_t0 = clock64();
#endif
  do {
    if (total_units < (long long)G * PLAN_FILL_MIN_LOAD || total_units > (long long)G * PLAN_FILL_MAX_LOAD || G > 1023) break;
    const int ideal = (int)(((unsigned)total_units * (unsigned)C.tile + (unsigned)G * (unsigned)C.first) / (unsigned)G);  // < 2^31: <= 96 tiles per CTA
    const int Lv = ideal + (C.tile * 3 * (m + 2)) / 4;  // margins of 1.5, 2.25, 3 tiles
    const float inv_tile = __frcp_rn((float)C.tile);
    // this lane's two objects (sort keys) and two segments
    unsigned okey[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int o = lane + 32 * u;
      const bool live = o < n_obj && S.nqt[o] > 0 && S.nt[o] > 0;
      // widest group first, then most tiles, then lowest index
      okey[u] = live ? (((unsigned)min(S.nqt[o], 255) << 24) | ((unsigned)min(S.nt[o], 0x3ffff) << 6) | (unsigned)(63 - o)) : 0u;
      if (o < n_obj) S.ns[1 + m][o] = 0;
    }
    int sload[2] = {0, 0};
    unsigned smeta[2] = {lane == 0 ? ((unsigned)G << 10) : 0u, 0u};  // begin | len << 10 | pieces << 20
    int nseg = 1, steps = 0, maxload = 0, max_np = 0;
    bool ok = true;
    while (ok) {
      const unsigned key = __reduce_max_sync(0xffffffffu, max(okey[0], okey[1]));
      if (key == 0u) break;
      const int o = 63 - (int)(key & 63u);
      if (okey[0] == key) okey[0] = 0u;
      if (okey[1] == key) okey[1] = 0u;
      const int nqt = (int)(key >> 24), g = 2 * nqt;
      if (g > G || nqt >= 255) { ok = false; break; }
      int rem = S.nt[o], slot = 0, t0 = 0;
      while (rem > 0) {
        if (++steps > PLAN_MAX_STEPS) { ok = false; break; }
        // pass 1: the g least-loaded CTAs (whole segments in ascending load; the last one may be split).  The plan CTA
        // shares its SM's issue slots with three pack CTAs, so what counts here is the instruction count: the common
        // case -- the least-loaded segment alone has g CTAs -- needs one REDUX and two shuffles per step.
        unsigned long long taken = 0ull;
        int need = g, base = 0, last_seg = 0, last_k = 0, first_len = 0;
        unsigned first_meta = 0u;
        while (need > 0) {
          unsigned k0 = 0xffffffffu;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = lane + 32 * u;
            if (((smeta[u] >> 10) & 1023u) > 0u && !((taken >> i) & 1ull)) k0 = min(k0, ((unsigned)sload[u] << 6) | (unsigned)i);
          }
          k0 = __reduce_min_sync(0xffffffffu, k0);
          if (k0 == 0xffffffffu) { ok = false; break; }
          const int i = (int)(k0 & 63u);
          const unsigned ma = __shfl_sync(0xffffffffu, smeta[0], i & 31), mb = __shfl_sync(0xffffffffu, smeta[1], i & 31);
          const unsigned mi = (i >> 5) ? mb : ma;
          const int len_i = (int)((mi >> 10) & 1023u);
          const int k = min(need, len_i);
          if (need == g) { first_len = len_i; first_meta = mi; }
          taken |= 1ull << i;
          need -= k;
          base = (int)(k0 >> 6);
          last_seg = i;
          last_k = k;
        }
        if (!ok) break;
        const bool single = first_len >= g;  // the first pick covered the whole group
        // chunk length: up to the level, within the chain bound, leaving neither a tiny tail nor more than the
        // remaining partial slots can take
        const int avail = Lv - base - (base == 0 ? C.first : C.extra);
        int room = avail > 0 ? (int)((float)avail * inv_tile) : 0;
        room -= (room * C.tile > avail) ? 1 : 0;  // floor
        const int slots_left = READ_MAX_SPLITS - slot;
        int ln, nb = 1;  // nb: identical chunks placed side by side by this step
        if (slots_left == 1) {
          ln = rem;
        } else {
          const int ln0 = min(max(room, (int)PLAN_MIN_CHUNK), (int)MAX_TILES_PER_SPLIT);
          ln = min(ln0, rem);
          const int lo = rem - MAX_TILES_PER_SPLIT * (slots_left - 1);
          if (ln < lo) ln = lo;
          const int tail = rem - ln;
          if (tail > 0 && tail < PLAN_MIN_CHUNK) ln = (rem <= min((int)MAX_TILES_PER_SPLIT, room + PLAN_MIN_CHUNK)) ? rem : rem - PLAN_MIN_CHUNK;
          // The next chunks of this object would take the next g CTAs of the same segment (same load, same room) and get
          // the same length as long as none of the rules above modifies it: place them all in this step.
          if (ln == ln0 && first_len >= 2 * g) {
            nb = min(min(plan_floor_div(first_len, g), slots_left - 1), plan_floor_div(rem, ln));
            for (; nb >= 2; --nb) {  // the rules as chunk number nb - 1 would see them
              const int rem_j = rem - (nb - 1) * ln, left_j = slots_left - (nb - 1), tail_j = rem_j - ln;
              if (left_j >= 2 && rem_j - MAX_TILES_PER_SPLIT * (left_j - 1) <= ln && (tail_j == 0 || tail_j >= PLAN_MIN_CHUNK)) break;
            }
            nb = max(nb, 1);
            if (nb > 1) last_k = g * nb;
          }
        }
        if (ln > MAX_TILES_PER_SPLIT) { ok = false; break; }  // (slots ran out before the tiles did: the dealt plan handles it)
        // pass 2: records + segment updates
        int unit = 0;
        auto place = [&](int i, int ld, unsigned mt) {  // segment i (load ld, meta mt) gives its first k CTAs to this chunk
          const int src = i & 31, hi = i >> 5;
          const int bg = (int)(mt & 1023u), sl = (int)((mt >> 10) & 1023u), np = (int)(mt >> 20) + 1;
          const int k = (i == last_seg) ? last_k : sl;
          const int nl = ld + (ld == 0 ? C.first : C.extra) + ln * C.tile;
          maxload = max(maxload, nl);
          max_np = max(max_np, np);
          if (nrec >= PLAN_MAX_RECORDS || (k < sl && nseg >= PLAN_MAX_SEGS) || nl >= (1 << 25) || np > PLAN_FILL_STRIDE) { ok = false; return; }
          if (lane == 0) {
            int4 *r = reinterpret_cast<int4 *>(&rec[nrec]);
            r[0] = make_int4(o | (slot << 8), t0, ln, bg);
            r[1] = make_int4(k, unit, nb > 1 ? g : k, 0);
          }
          const unsigned taken_meta = (unsigned)bg | ((unsigned)k << 10) | ((unsigned)np << 20);
          if (k == sl) {  // the whole segment moves up
            if (lane == src) { sload[hi] = nl; smeta[hi] = taken_meta; }
          } else {        // split: the first k CTAs become a new segment, the rest keep their load
            if (lane == src) smeta[hi] = (unsigned)(bg + k) | ((unsigned)(sl - k) << 10) | ((unsigned)(np - 1) << 20);
            if (lane == (nseg & 31)) { sload[nseg >> 5] = nl; smeta[nseg >> 5] = taken_meta; }
            ++nseg;
          }
          ++nrec;
          unit += k;
        };
        if (single) {
          place(last_seg, base, first_meta);
        } else {
          unsigned long long tm = taken;
          while (tm && ok) {
            const int i = __ffsll((long long)tm) - 1;
            tm &= tm - 1ull;
            const int src = i & 31;
            const int l0 = __shfl_sync(0xffffffffu, sload[0], src), l1 = __shfl_sync(0xffffffffu, sload[1], src);
            const unsigned m0 = __shfl_sync(0xffffffffu, smeta[0], src), m1 = __shfl_sync(0xffffffffu, smeta[1], src);
            place(i, (i >> 5) ? l1 : l0, (i >> 5) ? m1 : m0);
          }
        }
        if (!ok) break;
        rem -= ln * nb; t0 += ln * nb; slot += nb;
      }
      if (!ok) break;
      if (lane == 0) S.ns[1 + m][o] = slot;
    }
    if (ok && max_np <= PLAN_FILL_STRIDE) cost = (unsigned)maxload;
  } while (false);
  if (lane == 0) { S.cost[1 + m] = cost; S.nrec[m] = nrec; }
#ifdef RMNET_DEV
  if (s_chain_stamps && lane == 0) {
    atomicMax(s_chain_stamps + 16 + m, dev_globaltimer());
    atomicMax(s_chain_stamps + 20 + m, (unsigned long long)nrec);
    atomicMax(s_chain_stamps + 24 + m, (unsigned long long)(clock64() - dev_t0));
    atomicMax(s_chain_stamps + 28 + m, (unsigned long long)cost);
  }
#endif
  __syncwarp();
}

// ---- the plan role: called by ALL threads of one CTA (>= 160 threads, >= 5 warps) -----------------------------------
// bank_meta: per-slot counters; q_rects: query cell rectangles (nullptr = dense); temp_rects: when non-null, the cell
// rectangles that are being stored as the temporary frame by this very launch (their cell count replaces META_CELLS_T,
// exactly as bank_pack_kernel derives it).  Outputs: ns_out[o] partial slots per object (for merge.cu), hdr[c] =
// (number of pieces, first piece) of CTA c, pieces[] = (o | qt << 8 | half << 16 | slot << 20 | (live query rows - 1) << 24,
// first tile, tiles, stored cells of the object).
__device__ __forceinline__ void plan_build(PlanSmem &S, const int *__restrict__ bank_meta, const int *__restrict__ q_rects,
                                           const int *__restrict__ temp_rects, int cap, int n_obj, int h, int w, int G,
                                           int precision, int *__restrict__ ns_out, int2 *__restrict__ hdr,
                                           int4 *__restrict__ pieces, int piece_cap) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const PlanCost C = plan_cost(precision);
  DEV_STAMP_MAX(12);
  for (int o = tid; o < n_obj; o += blockDim.x) {
    const int4 qr = q_rects ? ld_dep(reinterpret_cast<const int4 *>(q_rects) + o) : make_int4(0, w - 1, 0, h - 1);
    int count;
    if (temp_rects) {
      const int base = ld_dep(bank_meta + o * 8 + META_CELLS_C);
      const int r = rect_cells(ld_dep(reinterpret_cast<const int4 *>(temp_rects) + o));
      count = base + (base + r > cap ? 0 : r);  // bank_pack_kernel drops a frame that would overflow the bank
    } else {
      count = ld_dep(bank_meta + o * 8 + META_CELLS_C) + ld_dep(bank_meta + o * 8 + META_CELLS_T);
    }
    S.count[o] = count;
    S.nt[o] = (count + KV_TILE - 1) / KV_TILE;
    S.qcells[o] = rect_cells(qr);
    S.nqt[o] = (rect_cells(qr) + UMMA_QT - 1) / UMMA_QT;
  }
  __syncthreads();
  DEV_STAMP_MAX(13);
  long long total_units = 0;
  for (int o = 0; o < n_obj; ++o) total_units += 2ll * S.nqt[o] * S.nt[o];
  if (warp == 0) plan_deal(S, n_obj, G, C);
  else if (warp <= PLAN_N_MARGINS) plan_fill(S, warp - 1, n_obj, G, C, total_units);
  __syncthreads();
  DEV_STAMP_MAX(14);
  if (tid == 0) {
    int win = 0;
    unsigned best = S.cost[0];
    for (int m = 0; m < PLAN_N_MARGINS; ++m)
      if (S.cost[1 + m] < best) { best = S.cost[1 + m]; win = 1 + m; }
    if (win == 0 && (size_t)((S.deal_items + G - 1) / G) * G > (size_t)piece_cap) win = -1;  // cannot happen for a workspace of the documented size
    S.winner = win;
  }
  __syncthreads();
  const int win = S.winner;
  for (int o = tid; o < n_obj; o += blockDim.x) ns_out[o] = win < 0 ? 0 : S.ns[win][o];
  for (int c = tid; c < G; c += blockDim.x) {
    if (win <= 0) {
      // deal: items c, c + G, ... in (o, chunk, half, qt) order with qt fastest
      const int n_items = win < 0 ? 0 : S.deal_items;
      const int stride = (n_items + G - 1) / G;
      int n = 0;
      for (int item = c; item < n_items; item += G, ++n) {
        int o = 0;
        while (S.ibase[o + 1] <= item) ++o;
        const unsigned r = (unsigned)(item - S.ibase[o]);
        const unsigned nqt = S.nqt[o], nt = S.nt[o], ns = S.ns[0][o];
        const unsigned q = r / nqt, qt = r - q * nqt, half = q & 1u, j = q >> 1;
        const unsigned t0 = (j * nt) / ns, t1 = ((j + 1u) * nt) / ns;  // balanced partition of the nt tiles into ns chunks
        const unsigned rows = (unsigned)min((int)UMMA_QT, S.qcells[o] - (int)qt * UMMA_QT);  // live query rows of this tile
        pieces[(size_t)c * stride + n] = make_int4((int)((unsigned)o | (qt << 8) | (half << 16) | (j << 20) | ((rows - 1u) << 24)), (int)t0, (int)(t1 - t0), S.count[o]);
      }
      hdr[c] = make_int2(n, c * stride);
    } else {
      const PlanRecord *rec = S.rec[win - 1];
      const int nrec = S.nrec[win - 1];
      int n = 0;
      for (int i = 0; i < nrec; ++i) {
        const PlanRecord r = rec[i];
        const unsigned d = (unsigned)(c - r.cta_begin);
        if (d < (unsigned)r.n_ctas) {
          const unsigned grp = d / (unsigned)r.gw;
          const int o = r.o_slot & 255, slot = (r.o_slot >> 8) + (int)grp;
          const unsigned unit = (unsigned)r.unit0 + (d - grp * (unsigned)r.gw), nqt = S.nqt[o];
          const unsigned half = unit / nqt, qt = unit - half * nqt;
          const unsigned rows = (unsigned)min((int)UMMA_QT, S.qcells[o] - (int)qt * UMMA_QT);
          pieces[(size_t)c * PLAN_FILL_STRIDE + n] = make_int4((int)((unsigned)o | (qt << 8) | (half << 16) | ((unsigned)slot << 20) | ((rows - 1u) << 24)), r.t0 + (int)grp * r.len, r.len, S.count[o]);
          ++n;
        }
      }
      hdr[c] = make_int2(n, c * PLAN_FILL_STRIDE);
    }
  }
}
#endif  // __CUDACC__

}  // namespace rmnet
