"""Dev aid: bench-style timed loop (no sync between steps), PDL on/off, same frame vs cycling frames."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for wlname in ("c2", "c3"):
    wl = bench.WORKLOADS[wlname]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    pool = bench.make_pool(wl, 1234, 8)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    for t in range(T - 1):
        d = D(pool["frames"][t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
    dfr = [D(f) for f in pool["frames"][T - 1:]]
    step = lambda d: rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False)
    for pdl in ("0", "1"):
        os.environ["RMNET_DISABLE_PDL"] = pdl
        for cycle in (False, True):
            for i in range(5): step(dfr[i % len(dfr) if cycle else 0])
            torch.cuda.synchronize()
            evs = []
            t0 = time.perf_counter()
            for i in range(30):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); step(dfr[i % len(dfr) if cycle else 0]); b.record()
                evs.append((a, b))
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            ts = [a.elapsed_time(b) * 1e3 for a, b in evs]
            print(f"{wlname} disable_pdl={pdl} cycle={cycle}: median {np.median(ts):7.1f} us  min {min(ts):7.1f}  max {max(ts):7.1f}  enqueue {1e6*(t1-t0)/30:.0f} us/step total {1e6*(t2-t0)/30:.0f}")
            print("    ", " ".join(f"{t:.0f}" for t in ts))
