"""CPU tests of the read kernel's planning rules (tests/plan_model.py restates rmnet_b200/csrc/sched.cuh; the GPU tests of
tests/test_gpu_plan.py compare the device-built plan with this restatement piece by piece): thousands of random bank states,
every plan must cover the work exactly once within the partial-slot, chain and piece-list bounds."""
import numpy as np
import pytest

import plan_model as pm


def _random_state(rng, regime):
    n = int(rng.integers(1, 11))
    if regime == "small":          # a few tiles per CTA: the dealt plan
        T, N = int(rng.integers(1, 6)), 1620
    elif regime == "wave":         # 12-96 tiles per CTA: water-filling applies
        T, N = int(rng.integers(6, 30)), 1620
    else:                          # hundreds of tiles per CTA (720p, long clips)
        T, N = int(rng.integers(20, 60)), 3600
    counts, q_cells = [], []
    for _ in range(n):
        if rng.uniform() < 0.1:
            counts.append(0 if rng.uniform() < 0.5 else int(rng.integers(1, 200)))
            q_cells.append(0 if rng.uniform() < 0.5 else int(rng.integers(1, 300)))
            continue
        f = rng.uniform(0.02, 1.0) * rng.uniform(0.1, 1.0)
        counts.append(int(T * N * f))
        q_cells.append(max(1, int(N * min(1.0, f * rng.uniform(0.5, 1.5)))))
    return counts, q_cells


@pytest.mark.parametrize("regime", ["small", "wave", "huge"])
@pytest.mark.parametrize("precision", [pm.PREC_SPLIT3, pm.PREC_MIXED], ids=["split3", "mixed"])
def test_plans_of_random_bank_states_cover_the_work_exactly_once(regime, precision):
    rng = np.random.default_rng({"small": 1, "wave": 2, "huge": 3}[regime] * 10 + precision)
    wins = [0, 0, 0, 0]
    eff = []
    for _ in range(400 if regime != "huge" else 120):
        counts, q_cells = _random_state(rng, regime)
        win, ns, lists = pm.build_plan(counts, q_cells, precision=precision)
        wins[win] += 1
        pm.check_plan(ns, lists, counts, q_cells, max_pieces=pm.FILL_STRIDE if win else None)
        loads = [sum(p[5] for p in pcs) for pcs in lists]
        if sum(loads) >= 12 * len(lists):
            eff.append(sum(loads) / len(lists) / max(loads))
    print(f"{regime}: planner chosen (deal, fill 1.5, 2.25, 3.0 tiles) = {wins}; mean tiles / max tiles per CTA {np.mean(eff) if eff else float('nan'):.3f}")
    if regime == "wave":
        assert sum(wins[1:]) > wins[0] // 4, "water-filling should win a good share of the single-wave states"


def test_the_bench_states_get_the_documented_plans():
    # the C3 state of tools/umma_timeline.py: water-filling, 37 tiles on the busiest CTA instead of 43
    counts, q_cells = [8215, 12790, 8523, 12059, 6272], [414, 140, 462, 630, 300]
    win, ns, lists = pm.build_plan(counts, q_cells)
    pm.check_plan(ns, lists, counts, q_cells, max_pieces=pm.FILL_STRIDE)
    assert win >= 1 and max(sum(p[5] for p in pcs) for pcs in lists) <= 38 and max(len(p) for p in lists) <= 2
    # the C2 state: 4-5 tiles per CTA, dealt
    counts, q_cells = [2121, 1368, 2120], [440, 270, 462]
    win, ns, lists = pm.build_plan(counts, q_cells)
    pm.check_plan(ns, lists, counts, q_cells)
    assert win == 0 and max(len(p) for p in lists) == 1
