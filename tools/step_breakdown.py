"""Per-op device time of one bench step (development aid)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
from rmnet_b200 import ops
dev = torch.device("cuda:0")
for wlname in ("c2", "c3"):
    wl = bench.WORKLOADS[wlname]
    n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    pool = bench.make_pool(wl, 1234, 2)
    h, w, lw, Wp = pool["h"], pool["w"], pool["lw"], pool["Wp"]
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    fr = pool["frames"]
    D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    for t in range(T - 1):
        d = D(fr[t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
    d = D(fr[T - 1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rects = ops.regional_boxes(d["mask"][None], None, True)[1][0, 1:n + 1].contiguous()
    rq = ops.regional_boxes(d["mask"][None], d["flow"][None], False)[1][0, 1:n + 1].contiguous()
    out = torch.empty((n, 1024, h, w), device=dev)
    opsd = {
        "boxes+rects (memorise side)": lambda: ops.regional_boxes(d["mask"][None], None, True),
        "bank.memorize(temp)": lambda: rm.bank.memorize(d["k4"], d["v4"], rects, False),
        "warp+boxes+rects (segment)": lambda: ops.regional_boxes(d["mask"][None], d["flow"][None], False),
        "read: attention kernel": lambda: rm.bank.read(d["qk"], d["qv"], rq, n, stages=1, out=out),
        "read: merge kernel": lambda: rm.bank.read(d["qk"], d["qv"], rq, n, stages=2, out=out),
        "whole step": lambda: (rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=False), rm.read(d["qk"], d["qv"], d["mask"][None], d["flow"][None])),
    }
    print("==", wlname, "cells/object", (rm.bank.stats()[:n, 0] + rm.bank.stats()[:n, 1]).tolist())
    for name, fn in opsd.items():
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"   {name:28s} {np.median(ts):8.1f} us (min {min(ts):.1f})")
