"""Dev aid: per-op device times of the scan kernels (generator, warp, fused warp+bbox, flow-affine) next to what the
reference runs for the same call on this GPU, with the HBM-roofline fraction of each (SURVEY 8d byte counts)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import json, synth
from rmnet_b200 import ops
from ref_composition import torch_warp
from test_gpu_parity import _ref_generator
gen = _ref_generator()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timed(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


for H, W in ((480, 854), (720, 1280)):
    K = 11
    rng = np.random.default_rng(3)
    lab = synth.rect_label_map(rng, 5, H, W)
    mask = torch.from_numpy(synth.soft_masks(rng, lab, K)[None]).to(dev)
    flow = torch.from_numpy(synth.flow_field(rng, H, W, 2.0)[None]).to(dev)
    of = torch.from_numpy(np.ascontiguousarray(np.moveaxis(synth.flow_field(rng, H, W, 3.0), 0, -1))).to(dev)
    m1, m2 = synth.affine_pair(rng)
    px = H * W
    rows = [
        ("generator, boxes only", lambda: ops.reg_att_map_forward(mask, want_att=False), 4 * (K - 1) * px),
        ("generator, boxes + full-res att_map", lambda: ops.reg_att_map_forward(mask), 4 * (K - 1) * px + 4 * K * px),
        ("warp (img1 + valid, 11 channels)", lambda: ops.warp(mask, flow), 4 * K * px + 8 * px + 8 * K * px),
        ("fused warp + bbox (get_att_map with flow), boxes only", lambda: ops.warp_att_map_forward(mask, flow, want_att=False), 4 * (K - 1) * px + 8 * px),
        ("update_optical_flow (device tensors)", lambda: ops.update_optical_flow_cuda(of, m1, m2), 16 * px),
    ]
    print(f"== {H}x{W}, K = {K} (HBM peak {PEAK:.0f} GB/s)")
    for name, fn, byts in rows:
        t = timed(fn)
        print(f"   {name:56s} {t:8.1f} us   {byts / 1e6:6.1f} MB  -> {byts / t / 1e3:7.1f} GB/s = {byts / t / 1e3 / PEAK:5.2f} of peak")
    if gen is not None:
        print(f"   {'reference generator kernel (oracle/_ref), same call':56s} {timed(lambda: gen.forward(mask, 0.5, 10, 64), 5):8.1f} us")
    print(f"   {'reference warp (torch ops, models/rmnet.py:252-278)':56s} {timed(lambda: torch_warp(mask, flow), 5):8.1f} us")
    gw = lambda: gen.forward(torch_warp(mask, flow)[0].contiguous(), 0.5, 10, 64)
    if gen is not None:
        print(f"   {'reference get_att_map(prev_mask, flow) = warp + generator':56s} {timed(gw, 5):8.1f} us")
