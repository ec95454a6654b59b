"""Dev aid: per-CTA phase timestamps of the tcgen05 kernel (first item of every CTA)."""
import ctypes, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
L = rmnet_b200.lib()
L.rmnet_debug_set_umma_dump.argtypes = [ctypes.c_void_p]; L.rmnet_debug_set_umma_dump.restype = None
L.rmnet_debug_set_umma_flags.argtypes = [ctypes.c_int]; L.rmnet_debug_set_umma_flags.restype = None
NO_TMA = "no_tma" in sys.argv   # experiment: producers stop loading after the ring fill -> is the tile time bound by the tile traffic?
dev = torch.device("cuda:0")
for wlname in ("c2", "c3"):
    wl = bench.WORKLOADS[wlname]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    pool = bench.make_pool(wl, 1234, 2)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    D = lambda f: {k: torch.from_numpy(v).to(dev) for k, v in f.items()}
    for t in range(T - 1):
        d = D(pool["frames"][t]); rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
    d = D(pool["frames"][T - 1])
    for _ in range(3): rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False)
    dbg = torch.zeros(8448 + 148 * 64 + 64, device=dev)
    torch.cuda.synchronize()
    standalone = len(sys.argv) > 1 and sys.argv[1] == "standalone"
    if standalone:   # the read kernel launched alone (no PDL chain): what bench.py's roofline times
        _, rq0 = rmnet_b200.ops.regional_boxes(d["mask"][None], d["flow"][None], padded_frame=False, k_scan=n + 1)
        rq0 = rq0[0, 1:n + 1].contiguous()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()
        torch.cuda.synchronize()
    L.rmnet_debug_set_umma_dump(dbg.data_ptr())
    if NO_TMA: L.rmnet_debug_set_umma_flags(1)
    if standalone:
        rm.bank.read(d["qk"], d["qv"], rq0, n, stages=1)
    else:
        rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False)
    torch.cuda.synchronize()
    L.rmnet_debug_set_umma_flags(0)
    L.rmnet_debug_set_umma_dump(None)
    ts = dbg[8448:8448 + 148 * 64].view(torch.int64).view(148, 32).cpu().numpy()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    np.save(os.path.join(ROOT, 'gpurun_out', f"timeline_{wlname}_{'standalone' if standalone else 'chained'}.npy"), ts)
    live = ts[:, 6] > 0
    t = ts[live]
    rel = (t[:, 1:7] - t[:, :1]).astype(np.float64)
    names = ["setup(sched,alloc,barriers)", "Q in TMEM", "first S loaded", "last P stored", "last PV retired", "epilogue stored"]
    print(f"== {wlname}: {live.sum()} CTAs with work; tiles/item median {np.median(t[:,7])}; cycles from kernel entry (median / max):")
    prev = 0
    for i, nm in enumerate(names):
        print(f"   {nm:30s} {np.median(rel[:, i]):9.0f} {rel[:, i].max():9.0f}   (+{np.median(rel[:, i]) - prev:7.0f})")
        prev = np.median(rel[:, i])
    ex = (t[:, 8:15] - t[:, :1]).astype(np.float64)
    for i, nm in enumerate(["setup begin", "plan loaded (after the dependency wait)", "-", "TMEM allocated", "softmax warpgroup starts", "Q loads issued", "Q converted+stored (before wait::st)"]):
        print(f"      .. {nm:38s} {np.median(ex[:, i]):9.0f}")
    order = np.argsort(rel[:, 5])
    print("   tiles of the first item, by CTA end time:", " ".join(f"{int(a)}" for a in t[order, 7][::4]))
    print("   end cycles (every 4th CTA):            ", " ".join(f"{int(a/1000)}" for a in rel[order, 5][::4]))
    _, rq = rmnet_b200.ops.regional_boxes(d["mask"][None], d["flow"][None], padded_frame=False)
    rqn = rq[0, 1:n + 1].cpu().numpy()
    st = rm.bank.stats()
    print("   cells/object", (st[:n, 0] + st[:n, 1]).tolist(), "query cells/object", [int(max(0, r[1]-r[0]+1) * max(0, r[3]-r[2]+1)) for r in rqn])
    wg = (t[:, 15] - t[:, 14]).astype(np.float64)
    wg = wg[(t[:, 15] > 0) & (t[:, 14] > 0)]
    if len(wg): print(f"   softmax warpgroup, second tile: S in registers -> P stored and signalled: median {np.median(wg):.0f} cycles (max {wg.max():.0f})")
    per_tile = (rel[:, 3] - rel[:, 2]) / np.maximum(t[:, 7] - 1, 1)
    print(f"   steady state cycles per tile (median) {np.median(per_tile):.0f}")
