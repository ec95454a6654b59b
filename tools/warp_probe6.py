import itertools, os, sys
import numpy as np, torch
import torch.nn.functional as F
f32 = np.float32
DEV = "cuda:0"
rng = np.random.default_rng(11)
H, W, C = 64, 80, 1
img = rng.random((1, C, H, W)).astype(f32)
ixw = (rng.random((H, W)) * (W - 3) + 1).astype(f32)
iyw = (rng.random((H, W)) * (H - 3) + 1).astype(f32)
ixw[:, :20] = rng.random((H, 20)).astype(f32)
iyw[:20, :] = rng.random((20, W)).astype(f32)
ixw[:, 70:] = (rng.random((H, 10)) * 3 - 2 + np.where(rng.random((H, 10)) < 0.5, 0, W - 1)).astype(f32)   # partially out of bounds
iyw[56:, :] = (rng.random((8, W)) * 3 - 2 + np.where(rng.random((8, W)) < 0.5, 0, H - 1)).astype(f32)
gx = (f32(2) * ixw / f32(W - 1) - f32(1)).astype(f32)
gy = (f32(2) * iyw / f32(H - 1) - f32(1)).astype(f32)
ti = torch.from_numpy(img).to(DEV)
vgrid = torch.stack([torch.from_numpy(gx), torch.from_numpy(gy)], -1)[None].to(DEV)
gs = F.grid_sample(ti, vgrid, align_corners=True).cpu().numpy()[0, 0]
ones = F.grid_sample(torch.ones_like(ti), vgrid, align_corners=True).cpu().numpy()[0, 0]
def fma(a, b, c): return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)
ix = ((gx + f32(1)) / f32(2)) * f32(W - 1); iy = ((gy + f32(1)) / f32(2)) * f32(H - 1)
fx, fy = np.floor(ix), np.floor(iy); x0, y0 = fx.astype(int), fy.astype(int)
def g(yy, xx, im):
    ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
    return np.where(ok, im[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0)).astype(f32)
def wpair(x, f, kind):
    if kind == "A": d = x - f; return f32(1) - d, d            # (w0, w1)
    if kind == "B": e = (f + f32(1)) - x; return e, f32(1) - e
    if kind == "C": d = x - f; e = (f + f32(1)) - x; return e, d
res = []
for kx, ky in itertools.product("ABC", "ABC"):
    wx0, wx1 = wpair(ix, fx, kx); wy0, wy1 = wpair(iy, fy, ky)
    wt = {"nw": wx0 * wy0, "ne": wx1 * wy0, "sw": wx0 * wy1, "se": wx1 * wy1}
    for im, target, nm in ((img[0, 0], gs, "img"), (np.ones((H, W), f32), ones, "ones")):
        v = {"nw": g(y0, x0, im), "ne": g(y0, x0 + 1, im), "sw": g(y0 + 1, x0, im), "se": g(y0 + 1, x0 + 1, im)}
        o = fma(v["se"], wt["se"], fma(v["sw"], wt["sw"], fma(v["nw"], wt["nw"], v["ne"] * wt["ne"])))
        bad = (o != target)
        inb = (fx >= 0) & (fx < W - 1) & (fy >= 0) & (fy < H - 1)
        res.append((int(bad.sum()), int((bad & inb).sum()), kx, ky, nm, float(np.abs(o - target).max())))
for r in sorted(res): print(r)
