"""GPU tests of the read kernel's device-built work plan (rmnet_b200/csrc/sched.cuh): for random bank states the plan
read back through the C ABI must cover every (object, query tile, Cv half, KV tile) exactly once within the partial-slot
and chain bounds, and the tcgen05 read that follows it must agree with the FFMA kernel (which has its own static split)."""
import numpy as np
import pytest
import torch

import plan_model
import rmnet_b200
from plan_model import check_plan
from rmnet_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _random_rects(rng, n, h, w, lo, hi, p_empty=0.0):
    out = np.zeros((n, 4), np.int32)
    for o in range(n):
        if rng.uniform() < p_empty:
            out[o] = (0, -1, 0, -1)
            continue
        bw = max(1, int(rng.uniform(lo, hi) * w))
        bh = max(1, int(rng.uniform(lo, hi) * h))
        x0 = int(rng.integers(0, w - bw + 1))
        y0 = int(rng.integers(0, h - bh + 1))
        out[o] = (x0, x0 + bw - 1, y0, y0 + bh - 1)
    return out


def _cells(r):
    return max(0, int(r[1]) - int(r[0]) + 1) * max(0, int(r[3]) - int(r[2]) + 1)


CASES = [
    # seed, n_obj, T, h, w, rect fraction range, p(empty rect)
    (1, 5, 20, 30, 54, (0.35, 0.7), 0.0),     # the headline shape (480p, 5 objects, T = 20)
    (2, 3, 5, 30, 54, (0.3, 0.7), 0.0),       # configs[1]
    (3, 1, 3, 30, 54, (0.8, 1.0), 0.0),       # one big object
    (4, 8, 12, 30, 54, (0.1, 0.5), 0.2),      # many small objects, some absent
    (5, 10, 24, 45, 80, (0.3, 0.8), 0.1),     # 720p, multi-round regime
    (6, 2, 2, 8, 12, (0.1, 0.4), 0.0),        # tiny
    (7, 6, 16, 30, 54, (0.05, 1.0), 0.1),     # very uneven objects
]


@pytest.mark.parametrize("case", CASES, ids=[f"seed{c[0]}_n{c[1]}_T{c[2]}_{c[3]}x{c[4]}" for c in CASES])
@pytest.mark.parametrize("prec", [rmnet_b200.RMNET_PREC_SPLIT3, rmnet_b200.RMNET_PREC_MIXED], ids=["split3", "mixed"])
def test_plan_covers_the_work_and_the_read_matches_the_ffma_kernel(case, prec):
    seed, n, T, h, w, (lo, hi), p_empty = case
    rng = np.random.default_rng(seed)
    g = torch.Generator(device=DEV).manual_seed(seed)
    bank = ops.MemoryBank(n, h, w, T, DEV)
    for t in range(T):
        rects = _random_rects(rng, n, h, w, lo, hi, p_empty)
        k4 = torch.randn((n, 128, h, w), generator=g, device=DEV) * 0.5
        v4 = torch.randn((n, 512, h, w), generator=g, device=DEV)
        bank.memorize(k4, v4, torch.from_numpy(rects).to(DEV), commit=(t < T - 1))   # the last frame stays temporary
    q_rects = _random_rects(rng, n, h, w, lo, hi, p_empty)
    qk = torch.randn((128, h, w), generator=g, device=DEV) * 0.5
    qv = torch.randn((512, h, w), generator=g, device=DEV)
    qr = torch.from_numpy(q_rects).to(DEV)
    got = bank.read(qk, qv, qr, n, precision=prec, impl=rmnet_b200.RMNET_IMPL_UMMA)
    ns, lists = bank.read_plan(n)
    st = bank.stats()
    counts = [int(st[o, 0] + st[o, 1]) for o in range(n)]
    q_cells = [_cells(r) for r in q_rects]
    check_plan(ns, lists, counts, q_cells)
    # the device-built plan against the plain-Python restatement of the planner (tests/plan_model.py), piece by piece
    win, ns_model, lists_model = plan_model.build_plan(counts, q_cells, G=len(lists), precision=prec)
    assert ns.tolist() == ns_model, f"partial slots per object: device {ns.tolist()}, model {ns_model} (planner {win})"
    for c, (dev_pcs, mod_pcs) in enumerate(zip(lists, lists_model)):
        assert [tuple(p) for p in dev_pcs] == [p[:7] for p in mod_pcs], f"CTA {c}: device {dev_pcs}, model {mod_pcs} (planner {win})"
    loads = [sum(p[5] for p in pcs) for pcs in lists]
    total = sum(loads)
    if total:
        print(f"plan (planner {win}): {sum(1 for l in loads if l)} CTAs busy, tiles/CTA max {max(loads)} mean {total / len(loads):.1f}, "
              f"pieces/CTA max {max(len(p) for p in lists)}, ns {ns.tolist()}")
    ref = bank.read(qk, qv, qr, n, precision=rmnet_b200.RMNET_PREC_SPLIT3, impl=rmnet_b200.RMNET_IMPL_SIMT)
    err = (got - ref).abs().max().item()
    tol = 2e-4 if prec == rmnet_b200.RMNET_PREC_SPLIT3 else 2e-3
    print(f"max-abs vs the FFMA kernel {err:.3e} (tolerance {tol})")
    assert err <= tol


def test_plan_is_rebuilt_by_every_frame_step():
    """rmnet_frame_step builds the plan in its pack launch (temporary frame counted from its cell rectangle): the plan after
    a step must describe the bank INCLUDING the frame that step stored, with or without commit."""
    import synth
    n, T, H, W = 3, 4, 240, 432
    rng = np.random.default_rng(11)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T + 1, device=DEV)
    h, w = rm.bank.h, rm.bank.w
    g = torch.Generator(device=DEV).manual_seed(5)
    for t in range(T):
        lab = synth.rect_label_map(rng, n, H, W)
        mask = torch.from_numpy(synth.onehot(lab, n + 1)).to(DEV)[None]
        flow = torch.from_numpy(synth.flow_field(rng, H, W, 2.0)).to(DEV)[None]
        k4 = torch.randn((n, 128, h, w), generator=g, device=DEV)
        v4 = torch.randn((n, 512, h, w), generator=g, device=DEV)
        qk = torch.randn((128, h, w), generator=g, device=DEV)
        qv = torch.randn((512, h, w), generator=g, device=DEV)
        commit = t % 2 == 0
        out = rm.step(k4, v4, mask, flow, qk, qv, commit=commit)
        boxes = out[1] if isinstance(out, tuple) else None
        ns, lists = rm.bank.read_plan(n)
        st = rm.bank.stats()
        counts = [int(st[o, 0] + st[o, 1]) for o in range(n)]
        got_counts = {}
        for pcs in lists:
            for p in pcs:
                got_counts[p[0]] = p[6]
        for o, cnt in got_counts.items():
            assert cnt == counts[o], f"frame {t} (commit={commit}): plan saw {cnt} cells of object {o}, the bank holds {counts[o]}"
        for o in range(n):
            if counts[o] and ns[o]:
                assert o in got_counts


@pytest.mark.parametrize("prec", [rmnet_b200.RMNET_PREC_SPLIT3, rmnet_b200.RMNET_PREC_MIXED], ids=["split3", "mixed"])
def test_device_plan_equals_the_model_on_random_states(prec):
    """Many bank states, cheaply: the per-object cell counters of an (otherwise empty) bank are poked directly, the query
    stage of the read (pack kernel: query roles + the plan role) is run alone, and the plan it leaves in the workspace is
    compared piece by piece with tests/plan_model.py -- single-wave states (water-filling, with its multi-segment takes,
    splits and bulk records), small ones (dealt) and one-object extremes."""
    n_max, T, h, w = 10, 40, 30, 54
    N = h * w
    rng = np.random.default_rng(100 + prec)
    bank = ops.MemoryBank(n_max, h, w, T, DEV)
    off = bank.ptr - bank.blob.data_ptr()
    meta = bank.blob[off:off + n_max * 32].view(torch.int32).view(n_max, 8)
    qk = torch.randn((128, h, w), device=DEV)
    qv = torch.randn((512, h, w), device=DEV)
    wins = [0, 0, 0, 0]
    for it in range(160):
        n = int(rng.integers(1, n_max + 1))
        t_eff = int(rng.integers(1, T + 1)) if it % 3 else int(rng.integers(1, 6))
        counts = [0 if rng.uniform() < 0.08 else int(min(bank.cap, t_eff * N * rng.uniform(0.02, 1.0) * rng.uniform(0.1, 1.0))) for _ in range(n)]
        q_rects = _random_rects(rng, n, h, w, 0.05, 1.0, p_empty=0.08)
        m = torch.zeros((n_max, 8), dtype=torch.int32)
        m[:n, 0] = torch.tensor(counts, dtype=torch.int32)
        meta.copy_(m.to(DEV))
        bank.read(qk, qv, torch.from_numpy(q_rects).to(DEV), n, precision=prec, impl=rmnet_b200.RMNET_IMPL_UMMA, stages=4)
        ns, lists = bank.read_plan(n)
        q_cells = [_cells(r) for r in q_rects]
        win, ns_model, lists_model = plan_model.build_plan(counts, q_cells, G=len(lists), precision=prec)
        wins[win] += 1
        assert ns.tolist() == ns_model, f"state {it}: counts {counts} q {q_cells}: device ns {ns.tolist()}, model {ns_model} (planner {win})"
        for c, (dev_pcs, mod_pcs) in enumerate(zip(lists, lists_model)):
            assert [tuple(p) for p in dev_pcs] == [p[:7] for p in mod_pcs], f"state {it} CTA {c}: device {dev_pcs}, model {mod_pcs} (planner {win}; counts {counts}, q {q_cells})"
        check_plan(ns, lists, counts, q_cells)
    print("planner chosen (deal, fill 1.5, 2.25, 3.0 tiles):", wins)
    assert sum(wins[1:]) >= 10, "the random states should exercise the water-filling planner"
