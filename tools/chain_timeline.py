"""Dev aid (`make -C rmnet_b200/csrc DEV=1` build): %globaltimer stamps of the frame-step chain, first CTA start / last CTA
end per kernel, relative to the region kernel's start (microseconds; median over repetitions, L2 flushed before each)."""
import ctypes, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
L = rmnet_b200.lib()
L.rmnet_debug_set_chain_stamps.argtypes = [ctypes.c_void_p]; L.rmnet_debug_set_chain_stamps.restype = None
dev = torch.device("cuda:0")
NAMES = ["regions start", "regions end", "pack start", "pack after wait", "pack end", "read start", "read end", "merge start",
         "merge gather after wait", "merge gather end", "merge fill end", "plan role end (pack kernel)", "plan role start", "plan: counts known", "plan: planners done"]
MIN_SLOTS = (0, 2, 3, 5, 7, 8, 37)
for wlname in sys.argv[1:] or ("c2", "c3"):
    wl = bench.WORKLOADS[wlname]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    pool = bench.make_pool(wl, 1234, 2)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    for t in range(T - 1):
        d = D(pool["frames"][t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
    d = D(pool["frames"][T - 1])
    step = lambda: rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False)
    for _ in range(3): step()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    init = torch.zeros(40, dtype=torch.int64); init[list(MIN_SLOTS)] = 2 ** 62
    rows = []
    for rep in range(12):
        st = init.to(dev)
        flush.zero_(); torch.cuda.synchronize()
        L.rmnet_debug_set_chain_stamps(st.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        L.rmnet_debug_set_chain_stamps(None)
        s = st.cpu().numpy().astype(np.float64)
        raw = st.cpu().numpy()
        rows.append(np.append((s[:15] - s[0]) / 1e3, e0.elapsed_time(e1) * 1e3))
    r = np.median(np.array(rows[2:]), axis=0)
    print(f"== {wlname}: step by CUDA events {r[15]:.1f} us")
    for i, nm in enumerate(NAMES): print(f"   {nm:26s} {r[i]:8.2f} us")
    print("   regions: first CTA leaves the pixel loop %.2f us, last CTA %.2f us, last ticket taken %.2f us" % tuple((raw[i] - raw[0]) / 1e3 for i in (37, 36, 38)))
    print(f"   plan: deal done {(raw[15] - raw[0]) / 1e3:.2f} us (cost {raw[19]}); fill margins done at", [round((raw[16 + m] - raw[0]) / 1e3, 2) for m in range(4)],
          "records", [int(raw[20 + m]) for m in range(4)], "cycles", [int(raw[24 + m]) for m in range(4)], "cost", [int(raw[28 + m]) for m in range(4)])
