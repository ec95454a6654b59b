"""Dev aid: chained (PDL) frame step vs RMNET_DISABLE_PDL=1, with / without the L2 flush, per workload."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for wlname in ("c2", "c3"):
    wl = bench.WORKLOADS[wlname]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    pool = bench.make_pool(wl, 1234, 2)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    for t in range(T - 1):
        d = D(pool["frames"][t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
    d = D(pool["frames"][T - 1])
    step = lambda: rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False)
    for pdl in ("0", "1"):
        os.environ["RMNET_DISABLE_PDL"] = pdl
        for do_flush in (True, False):
            for _ in range(3): step()
            ts = []
            for _ in range(20):
                if do_flush: flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); step(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            print(f"{wlname} disable_pdl={pdl} flush={do_flush}: median {np.median(ts):7.1f} us  min {min(ts):7.1f}  max {max(ts):7.1f}")
