import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
from rmnet_b200 import ops
dev = torch.device("cuda:0")
wl = bench.WORKLOADS["c2"]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
pool = bench.make_pool(wl, 1234, 8)
rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
fr = pool["frames"]
D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
for t in range(T - 1):
    d = D(fr[t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
ds = [D(f) for f in fr[T - 1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def T_(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return f"{np.median(ts):8.1f} us (min {min(ts):.1f})"
d = ds[0]
print("frame_regions (both)   ", T_(lambda: ops.frame_regions(d["mask"][None], d["flow"][None])))
print("regional_boxes warp    ", T_(lambda: ops.regional_boxes(d["mask"][None], d["flow"][None], False)))
print("regional_boxes direct  ", T_(lambda: ops.regional_boxes(d["mask"][None], None, True)))
print("rm.step                ", T_(lambda: rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False)))
for i, dd in enumerate(ds[:4]):
    print(f"rm.step frame {i}       ", T_(lambda: rm.step(dd["k4"], dd["v4"], dd["mask"][None], dd["flow"][None], dd["qk"], dd["qv"], commit=False)), rm.bank.stats()[:n, :2].tolist())
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(200):
    dd = ds[i % len(ds)]; rm.step(dd["k4"], dd["v4"], dd["mask"][None], dd["flow"][None], dd["qk"], dd["qv"], commit=False)
torch.cuda.synchronize(); print("back-to-back wall per step (no flush): %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
t0 = time.perf_counter()
for i in range(200):
    dd = ds[i % len(ds)]; ops.frame_regions(dd["mask"][None], dd["flow"][None])
print("host-side cost of frame_regions call: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6)); torch.cuda.synchronize()
