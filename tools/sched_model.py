"""Offline model of the read kernel's schedule (no GPU): makespan in KV-tile units of
  (a) the shipped scheduler (one chunk length per object, items dealt round-robin; common.cuh sched_build),
  (b) an exact stream-K split (every CTA gets U/G tile-units; a CTA's range may span unit boundaries -> extra pieces),
for random clip states shaped like the bench workloads.  piece overhead = tiles of prologue / drain per extra piece."""
import math, sys
import numpy as np

G, MAXC, MAXS = 148, 64, 16


def shipped(nt, nqt, ov_first=5.0):
    best = None
    max_nt = max(nt)
    c_min = max(1, math.ceil(max_nt / MAXS))
    cands = range(c_min, min(MAXC, max_nt) + 1) if c_min <= MAXC else [c_min]   # huge banks: the slot bound wins
    for c in cands:
        ns = [math.ceil(t / c) for t in nt]
        items = sum(s * 2 * q for s, q in zip(ns, nqt))
        longest = max(math.ceil(t / s) for t, s in zip(nt, ns))
        rounds = math.ceil(items / G)
        cost = rounds * (longest + 5) * 128 + (64 - c)
        if best is None or cost < best[0]:
            best = (cost, c, ns, items, longest, rounds)
    _, c, ns, items, longest, rounds = best
    # actual makespan: items dealt round-robin in (o, j, half, qt) order; CTA k runs items k, k+G, ...
    lens = []
    for t, s, q in zip(nt, ns, nqt):
        for j in range(s):
            ln = (j + 1) * t // s - j * t // s
            lens += [ln] * (2 * q)
    load = np.zeros(G)
    for i, ln in enumerate(lens):
        load[i % G] += ln + (ov_first if i < G else PIECE_OV)
    return load.max(), items


def streamk(nt, nqt, ov_first=5.0):
    units = []
    for t, q in zip(nt, nqt):
        units += [t] * (2 * q)
    U = sum(units)
    L = math.ceil(U / G)
    # CTA k covers [k*L, (k+1)*L): count the unit boundaries inside -> pieces
    bounds = np.cumsum(units)
    worst = 0.0
    for k in range(G):
        a, b = k * L, min((k + 1) * L, U)
        if a >= b:
            continue
        pieces = 1 + int(((bounds > a) & (bounds < b)).sum())
        worst = max(worst, (b - a) + ov_first + (pieces - 1) * PIECE_OV)
    return worst, U / G


PIECE_OV = float(sys.argv[1]) if len(sys.argv) > 1 else 2.5
rng = np.random.default_rng(0)
for name, n_obj, T, N in (("c2-like", 3, 5, 1620), ("c3-like", 5, 20, 1620), ("c4-like", 10, 40, 3600)):
    ratios, eff_s, eff_k = [], [], []
    for _ in range(200):
        f_m = rng.uniform(0.15, 0.45, n_obj) * rng.uniform(0.15, 0.45, n_obj) / 0.09 * 0.3   # region fraction per object
        f_q = np.clip(f_m * rng.uniform(0.7, 1.3, n_obj), 0.03, 1.0)
        nt = [max(1, math.ceil(T * N * f / 64)) for f in np.clip(f_m, 0.03, 1.0)]
        nqt = [max(1, math.ceil(N * f / 128)) for f in f_q]
        ideal = sum(t * 2 * q for t, q in zip(nt, nqt)) / G
        a, _ = shipped(nt, nqt)
        b, _ = streamk(nt, nqt)
        ratios.append(b / a); eff_s.append(ideal / a); eff_k.append(ideal / b)
    print(f"{name}: piece overhead {PIECE_OV} tiles: shipped schedule reaches {np.mean(eff_s):.2f} of the ideal makespan (p10 {np.percentile(eff_s, 10):.2f}), "
          f"exact stream-K {np.mean(eff_k):.2f} (p10 {np.percentile(eff_k, 10):.2f}); stream-K / shipped makespan = {np.mean(ratios):.2f}")
