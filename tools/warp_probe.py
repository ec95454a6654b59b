"""Diagnostic: which floating-point evaluation order does torch's CUDA backend use for RMNet.warp?"""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle, synth
from rmnet_b200 import ops
f32 = np.float32
DEV = "cuda:0"
rng = np.random.default_rng(31)
H, W, C = 48, 64, 3
img = rng.random((1, C, H, W)).astype(f32)
flow = (rng.standard_normal((1, 2, H, W)) * 2).astype(f32)
ti, tf = torch.from_numpy(img).to(DEV), torch.from_numpy(flow).to(DEV)
xs = torch.arange(W, device=DEV).float().view(1, -1).repeat(H, 1)
ys = torch.arange(H, device=DEV).float().view(-1, 1).repeat(1, W)
vx, vy = xs + tf[0, 0], ys + tf[0, 1]
gx_t = (2.0 * vx / max(W - 1, 1) - 1.0).cpu().numpy()
gy_t = (2.0 * vy / max(H - 1, 1) - 1.0).cpu().numpy()
vxn, vyn = vx.cpu().numpy(), vy.cpu().numpy()
def variants(v, d):
    return {"div": (f32(2) * v) / f32(d) - f32(1), "rcp": (f32(2) * v) * (f32(1) / f32(d)) - f32(1),
            "rcp64": ((f32(2) * v).astype(np.float64) * (1.0 / d)).astype(f32) - f32(1),
            "fma_rcp": ((f32(2) * v).astype(np.float64) * np.float64(f32(1) / f32(d)) - 1.0).astype(f32),
            "two_over_d": (v * f32(2.0 / d)) - f32(1),
            "fma_div": ((f32(2) * v).astype(np.float64) / d - 1.0).astype(f32)}
for nm, arr in variants(vxn, W - 1).items():
    print("norm x", nm, int((arr != gx_t).sum()))
for nm, arr in variants(vyn, H - 1).items():
    print("norm y", nm, int((arr != gy_t).sum()))
# grid_sample given torch's own normalised grid
vgrid = torch.stack([torch.from_numpy(gx_t), torch.from_numpy(gy_t)], -1)[None].to(DEV)
gs = F.grid_sample(ti, vgrid, align_corners=True).cpu().numpy()[0]
ones = F.grid_sample(torch.ones_like(ti), vgrid, align_corners=True).cpu().numpy()[0]
def fma(a, b, c): return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)
def sample(unnorm, order, wform):
    if unnorm == "A": ix = ((gx_t + f32(1)) / f32(2)) * f32(W - 1); iy = ((gy_t + f32(1)) / f32(2)) * f32(H - 1)
    if unnorm == "B": ix = fma(gx_t, np.full_like(gx_t, (W - 1) / 2), np.full_like(gx_t, (W - 1) / 2)); iy = fma(gy_t, np.full_like(gy_t, (H - 1) / 2), np.full_like(gy_t, (H - 1) / 2))
    if unnorm == "C": ix = ((gx_t.astype(np.float64) + 1) / 2 * (W - 1)).astype(f32); iy = ((gy_t.astype(np.float64) + 1) / 2 * (H - 1)).astype(f32)
    fx, fy = np.floor(ix), np.floor(iy)
    if wform == "se":
        wx0, wy0, wx1, wy1 = (fx + 1) - ix, (fy + 1) - iy, ix - fx, iy - fy
    else:
        wx1, wy1 = ix - fx, iy - fy; wx0, wy0 = f32(1) - wx1, f32(1) - wy1
    nw, ne, sw, se = wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1
    x0, y0 = fx.astype(int), fy.astype(int)
    outs = []
    for c in range(C):
        def g(yy, xx):
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            return np.where(ok, img[0, c][np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0)).astype(f32)
        a, b, cc, d = g(y0, x0), g(y0, x0 + 1), g(y0 + 1, x0), g(y0 + 1, x0 + 1)
        if order == "fma": o = fma(d, se, fma(cc, sw, fma(b, ne, a * nw)))
        if order == "nofma": o = ((a * nw + b * ne) + cc * sw) + d * se
        outs.append(o)
    return np.stack(outs)
for un in "ABC":
    for order in ("fma", "nofma"):
        for wf in ("se", "one"):
            print("sample", un, order, wf, int((sample(un, order, wf) != gs).sum()), "of", gs.size)
# whole pipeline comparisons
t1 = None
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import _torch_warp
r1, rm = _torch_warp(ti, tf)
mine, mv = ops.warp(ti, tf)
o_cuda, _ = oracle.warp(img, flow, arith="cuda"); o_cpu, _ = oracle.warp(img, flow, arith="cpu")
print("mine vs torch", int((mine != r1).sum()), "valid", int((mv != rm).sum()), "maxdiff", float((mine - r1).abs().max()))
print("oracle cuda vs torch", int((o_cuda != r1.cpu().numpy()).sum()), " oracle cpu vs torch", int((o_cpu != r1.cpu().numpy()).sum()))
print("mine vs oracle cuda", int((mine.cpu().numpy() != o_cuda).sum()))
