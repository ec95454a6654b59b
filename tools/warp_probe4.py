"""Where does the cuDNN-order pipeline still differ from torch (cuDNN enabled)?"""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle, synth
from rmnet_b200 import ops
from test_gpu_parity import _torch_warp, _warp_cases
f32 = np.float32
DEV = "cuda:0"
for name, img, flow in _warp_cases()[:3]:
    ti, tf = torch.from_numpy(img).to(DEV), torch.from_numpy(flow).to(DEV)
    ref1, refm = _torch_warp(ti, tf)
    mine, mv = ops.warp(ti, tf)
    o1, om = oracle.warp(img, flow, arith="cuda_cudnn")
    r = ref1.cpu().numpy(); m = mine.cpu().numpy()
    bad = np.argwhere(m != r)
    print(name, "mine!=torch:", len(bad), "oracle!=torch:", int((o1 != r).sum()), "mine!=oracle:", int((m != o1).sum()), "of", r.size)
    B, C, H, W = img.shape
    # recompute coordinates
    xs = np.arange(W, dtype=f32)[None, :].repeat(H, 0); ys = np.arange(H, dtype=f32)[:, None].repeat(W, 1)
    vx, vy = xs + flow[0, 0], ys + flow[0, 1]
    gx = (f32(2) * vx) * (f32(1) / f32(W - 1)) - f32(1); gy = (f32(2) * vy) * (f32(1) / f32(H - 1)) - f32(1)
    ix = ((gx + f32(1)) / f32(2)) * f32(W - 1); iy = ((gy + f32(1)) / f32(2)) * f32(H - 1)
    nb = 0
    for (b, c, y, x) in bad[:12]:
        fx, fy = np.floor(ix[y, x]), np.floor(iy[y, x])
        inb = (0 <= fx < W - 1) and (0 <= fy < H - 1)
        nb += inb
        print(f"   c={c} y={y} x={x} ix={ix[y,x]:.6f} iy={iy[y,x]:.6f} inbounds4={inb} torch={r[b,c,y,x]:.9g} mine={m[b,c,y,x]:.9g} valid={refm[b,c,y,x].item()}")
    allin = sum(1 for (b, c, y, x) in bad if (0 <= np.floor(ix[y, x]) < W - 1) and (0 <= np.floor(iy[y, x]) < H - 1))
    print("   mismatches with all 4 taps in bounds:", allin, "of", len(bad))
    # is torch's result for these equal to the ATen order instead?
    o2, _ = oracle.warp(img, flow, arith="cuda_native")
    print("   torch == aten-order oracle at mismatches:", int(sum(o2[tuple(i)] == r[tuple(i)] for i in bad)), "of", len(bad))
    # direct check: grid_sample alone with cudnn on the pipeline's grid
    vgrid = torch.stack([torch.from_numpy(gx), torch.from_numpy(gy)], -1)[None].to(DEV)
    gs = F.grid_sample(ti, vgrid, align_corners=True).cpu().numpy()
    ones = F.grid_sample(torch.ones_like(ti), vgrid, align_corners=True).cpu().numpy()
    # emulate
    def fma(a, b, c): return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)
    fx, fy = np.floor(ix), np.floor(iy); x0, y0 = fx.astype(int), fy.astype(int)
    dx, dy = ix - fx, iy - fy; ex, ey = f32(1) - dx, f32(1) - dy
    wt = {"nw": ex * ey, "ne": dx * ey, "sw": ex * dy, "se": dx * dy}
    cnt = {}
    for c in range(C):
        def g(yy, xx):
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            return np.where(ok, img[0, c][np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0)).astype(f32)
        v = {"nw": g(y0, x0), "ne": g(y0, x0 + 1), "sw": g(y0 + 1, x0), "se": g(y0 + 1, x0 + 1)}
        import itertools
        for order in itertools.permutations(["nw", "ne", "sw", "se"]):
            a, b2, c2, d = order
            o = fma(v[d], wt[d], fma(v[c2], wt[c2], fma(v[b2], wt[b2], v[a] * wt[a])))
            cnt[order] = cnt.get(order, 0) + int((o != gs[0, c]).sum())
    best = sorted(cnt.items(), key=lambda kv: kv[1])[:4]
    print("   grid_sample-only emulation mismatches by order:", best)
