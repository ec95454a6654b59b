"""Offline view (no GPU) of the read kernel's work plan, rmnet_b200/csrc/sched.cuh, through its plain-Python restatement
tests/plan_model.py (which the GPU tests hold equal to the device-built plan piece by piece): for random clip states shaped
like the bench workloads and for the bench's own C3 state, the makespan of the dealt plan alone and of the plan the device
picks (deal or water-filling with three margins), against the ideal (perfectly divisible work), under the cycle costs
measured with tools/umma_timeline.py (KV tile 2 544, first piece 17 600, every further piece 5 195 cycles)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import plan_model as pm  # noqa: E402

TRUE = (2544, 17600, 5195)


def makespan(lists):
    tile, first, extra = TRUE
    return max((sum(p[5] for p in pcs) * tile + (first + (len(pcs) - 1) * extra if pcs else 0)) for pcs in lists)


def ideal(lists):
    tile, first, _ = TRUE
    return sum(sum(p[5] for p in pcs) for pcs in lists) * tile / len(lists) + first


def deal_only(counts, q_cells, precision):
    saved = pm.N_MARGINS
    pm.N_MARGINS = 0
    try:
        return pm.build_plan(counts, q_cells, precision=precision)
    finally:
        pm.N_MARGINS = saved


if __name__ == "__main__":
    for mode, prec in (("strict", pm.PREC_SPLIT3), ("mixed", pm.PREC_MIXED)):
        rng = np.random.default_rng(0)
        cols = []
        for name, n_obj, T, N in (("c2-like", 3, 5, 1620), ("c3-like", 5, 20, 1620), ("8 obj T=20", 8, 20, 1620), ("5 obj T=10", 5, 10, 1620), ("c4-like", 10, 40, 3600)):
            e_deal, e_plan, fills = [], [], 0
            for _ in range(150):
                f_m = rng.uniform(0.15, 0.45, n_obj) * rng.uniform(0.15, 0.45, n_obj) / 0.09 * 0.3   # region fraction per object
                f_q = np.clip(f_m * rng.uniform(0.7, 1.3, n_obj), 0.03, 1.0)
                counts = [int(T * N * f) for f in np.clip(f_m, 0.03, 1.0)]
                q_cells = [max(1, int(N * f)) for f in f_q]
                _, _, l_deal = deal_only(counts, q_cells, prec)
                win, _, l_plan = pm.build_plan(counts, q_cells, precision=prec)
                fills += win > 0
                e_deal.append(ideal(l_deal) / makespan(l_deal)); e_plan.append(ideal(l_plan) / makespan(l_plan))
            cols.append(f"{name}: deal {np.mean(e_deal):.3f} -> plan {np.mean(e_plan):.3f} ({fills} of 150 water-filled)")
        print(f"{mode}: ideal / makespan  |  " + "  |  ".join(cols))
    counts, q_cells = [8215, 12790, 8523, 12059, 6272], [414, 140, 462, 630, 300]      # the bench's C3 state (tools/umma_timeline.py)
    _, _, l_deal = deal_only(counts, q_cells, pm.PREC_SPLIT3)
    win, ns, l_plan = pm.build_plan(counts, q_cells)
    nt = [math.ceil(c / 64) for c in counts]
    print(f"bench C3 state (KV tiles {nt}): dealt plan {makespan(l_deal)} cycles, busiest CTA {max(sum(p[5] for p in l) for l in l_deal)} tiles; "
          f"device's choice (planner {win}) {makespan(l_plan)} cycles, busiest CTA {max(sum(p[5] for p in l) for l in l_plan)} tiles in "
          f"{max(len(l) for l in l_plan)} pieces, chunks per object {ns}; ideal {ideal(l_plan):.0f}")
