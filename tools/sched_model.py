"""Offline model (no GPU) of the read kernel's work plan, rmnet_b200/csrc/sched.cuh: makespan of the two planners --
"deal" (one chunk length, pieces dealt round-robin) and "fill" (water-filling under a target level, three margins) --
against the ideal (perfectly divisible work), for random clip states shaped like the bench workloads and for the bench's
own C3 state.  The device code follows this model step for step (its per-margin costs, read back with the DEV stamps of
tools/chain_timeline.py, are the numbers this script prints for the C3 state: 117 200 / 117 200 / 119 900 cycles)."""
import math
import numpy as np

G, MAXS, MAXC, MINC = 148, 16, 64, 4
COST = {"strict": (2700, 14000, 6000), "mixed": (1740, 14000, 6000)}   # cycles: KV tile, first piece, every further piece
MARGINS = (1.5, 2.25, 3.0)                                             # tiles above the ideal level
FILL_MIN_LOAD, FILL_MAX_LOAD = 12, 96                                  # tiles per CTA


def fill(nt, nqt, tile, first, extra, margin):
    W = sum(2 * q * t for q, t in zip(nqt, nt))
    ideal = (W * tile + G * first) // G
    if not (G * FILL_MIN_LOAD <= W <= G * FILL_MAX_LOAD):
        return None, ideal
    Lv = ideal + int(tile * margin)
    segs = [[0, 0, G]]                                                 # load, first CTA, CTAs
    order = sorted([o for o in range(len(nt)) if nt[o] > 0 and nqt[o] > 0], key=lambda o: (-nqt[o], -nt[o], o))
    maxload, nrec = 0, 0
    for o in order:
        g, rem, slot = 2 * nqt[o], nt[o], 0
        while rem > 0:
            need, taken, base = g, [], 0
            for i in sorted(range(len(segs)), key=lambda i: (segs[i][0], i)):
                if need == 0:
                    break
                k = min(need, segs[i][2]); taken.append((i, k)); need -= k; base = segs[i][0]
            room = max(0, Lv - base - (first if base == 0 else extra)) // tile
            left = MAXS - slot
            if left == 1:
                ln = rem
            else:
                ln = min(max(room, MINC), MAXC, rem)
                ln = max(ln, rem - MAXC * (left - 1))
                if 0 < rem - ln < MINC:
                    ln = rem if rem <= min(MAXC, room + MINC) else rem - MINC
            for i, k in taken:
                s = segs[i]
                nl = s[0] + (first if s[0] == 0 else extra) + ln * tile
                maxload = max(maxload, nl); nrec += 1
                if k == s[2]:
                    s[0] = nl
                else:
                    segs.append([nl, s[1], k]); s[1] += k; s[2] -= k
            rem -= ln; slot += 1
    return maxload, ideal


def deal(nt, nqt, tile, first, extra):
    max_nt = max(t for t, q in zip(nt, nqt) if q > 0)
    cmin = max(1, math.ceil(max_nt / MAXS))
    best = None
    for c in (range(cmin, min(MAXC, max_nt) + 1) if cmin <= MAXC else [cmin]):
        ns = [math.ceil(t / c) if q > 0 else 0 for t, q in zip(nt, nqt)]
        items = sum(s * 2 * q for s, q in zip(ns, nqt))
        longest = max(math.ceil(t / s) for t, s in zip(nt, ns) if s > 0)
        rounds = math.ceil(items / G)
        cost = (rounds * longest * tile + first + (rounds - 1) * extra, 64 - c)
        if best is None or cost < best[0]:
            best = (cost, ns)
    ns = best[1]
    lens = [ln for t, s, q in zip(nt, ns, nqt) for j in range(s) for ln in [(j + 1) * t // s - j * t // s] * (2 * q)]
    loads = [0] * G
    for i, ln in enumerate(lens):
        loads[i % G] += ln * tile + (first if i < G else extra)
    return max(loads)


if __name__ == "__main__":
    for mode, (tile, first, extra) in COST.items():
        rng = np.random.default_rng(0)
        cols = []
        for name, n_obj, T, N in (("c2-like", 3, 5, 1620), ("c3-like", 5, 20, 1620), ("8 obj T=20", 8, 20, 1620), ("5 obj T=10", 5, 10, 1620), ("c4-like", 10, 40, 3600)):
            e_deal, e_best = [], []
            for _ in range(150):
                f_m = rng.uniform(0.15, 0.45, n_obj) * rng.uniform(0.15, 0.45, n_obj) / 0.09 * 0.3   # region fraction per object
                f_q = np.clip(f_m * rng.uniform(0.7, 1.3, n_obj), 0.03, 1.0)
                nt = [max(1, math.ceil(T * N * f / 64)) for f in np.clip(f_m, 0.03, 1.0)]
                nqt = [max(1, math.ceil(N * f / 128)) for f in f_q]
                d = deal(nt, nqt, tile, first, extra)
                best, ideal = d, None
                for m in MARGINS:
                    f, ideal = fill(nt, nqt, tile, first, extra, m)
                    if f is not None and f < best:
                        best = f
                e_deal.append(ideal / d); e_best.append(ideal / best)
            cols.append(f"{name}: deal {np.mean(e_deal):.3f} -> plan {np.mean(e_best):.3f}")
        print(f"{mode}: ideal / makespan  |  " + "  |  ".join(cols))
    nt, nqt = [129, 200, 134, 189, 98], [4, 2, 4, 5, 3]                 # the bench's C3 state (tools/umma_timeline.py)
    tile, first, extra = COST["strict"]
    print("bench C3 state: deal", deal(nt, nqt, tile, first, extra), "cycles; fill", [fill(nt, nqt, tile, first, extra, m)[0] for m in MARGINS],
          "; ideal", fill(nt, nqt, tile, first, extra, 1.5)[1])
