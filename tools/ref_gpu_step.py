"""GPU-vs-GPU bar: the REFERENCE's composition of the frame step (tests/ref_composition.py: torch's CUDA ops + the reference's
unmodified CUDA extension from oracle/_ref) against RegionalMemory.step on the same B200, same inputs, L2 flushed."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rmnet_b200
from ref_composition import ReferenceClip
from test_gpu_parity import _ref_generator
gen = _ref_generator()
assert gen is not None, "oracle/_ref/reg_att_map_generator*.so is missing (make -C oracle ref)"
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), out


for wlname in ("c2", "c3"):
    wl = bench.WORKLOADS[wlname]; n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    pool = bench.make_pool(wl, 1234, 1)
    fr = [{k: torch.from_numpy(v).to(dev) for k, v in f.items()} for f in pool["frames"]]
    ref = ReferenceClip(gen, n, bench.K_CH, H, W)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    for t in range(T - 1):
        ref.commit(fr[t])
        rm.memorize(fr[t]["k4"], fr[t]["v4"], fr[t]["mask"][None], commit=True)
    cur = fr[T - 1]
    t_ref, (m_ref, _, _) = timed(lambda: ref.step(cur), 10)
    t_ours, (m_ours, _, _) = timed(lambda: rm.step(cur["k4"], cur["v4"], cur["mask"][None], cur["flow"][None], cur["qk"], cur["qv"], commit=False), 20)
    err = (m_ours - m_ref).abs().max().item()
    print(f"{wlname}: reference composition on this GPU {t_ref:9.1f} us/frame   rmnet_b200 step {t_ours:7.1f} us/frame   "
          f"ratio {t_ref / t_ours:6.1f}x   max|mem_val - reference| {err:.2e}")
