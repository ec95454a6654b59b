"""Run a few frames of rm.step for c2 and c3 (for ncu launch lists)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
dev = torch.device("cuda:0")
for wlname in sys.argv[1:] or ["c2", "c3"]:
    wl = bench.WORKLOADS[wlname]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    pool = bench.make_pool(wl, 1234, 2)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    for t in range(T - 1):
        d = D(pool["frames"][t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
    d = D(pool["frames"][T - 1])
    torch.cuda.synchronize()
    for _ in range(4):
        rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False)
    torch.cuda.synchronize()
    print(wlname, "done")
