"""Dev aid: where do multi-ms outlier steps come from?  bench-style loop, 3 repetitions, with / without preallocated out."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
wl = bench.WORKLOADS["c2"]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
pool = bench.make_pool(wl, 1234, 8)
rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
for t in range(T - 1):
    d = D(pool["frames"][t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
dfr = [D(f) for f in pool["frames"][T - 1:]]
outb = torch.empty((n, 1024, rm.h, rm.w), device=dev)
for prealloc in (False, True, False, True):
    step = lambda d: rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False, out=outb if prealloc else None)
    for i in range(10): step(dfr[i % 8])
    torch.cuda.synchronize()
    evs = []; cpu = []
    for i in range(200):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record(); step(dfr[i % 8]); b.record(); cpu.append(time.perf_counter() - t0)
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = np.array([a.elapsed_time(b) * 1e3 for a, b in evs]); cpu = np.array(cpu) * 1e6
    bad = np.nonzero(ts > 3 * np.median(ts))[0]
    print(f"prealloc={prealloc}: mean {ts.mean():.1f} median {np.median(ts):.1f} max {ts.max():.1f} us; outliers (idx: gpu_us / cpu_us):",
          " ".join(f"{i}:{ts[i]:.0f}/{cpu[i]:.0f}" for i in bad), "| cpu median", np.median(cpu), "max", cpu.max(), "at", cpu.argmax())
