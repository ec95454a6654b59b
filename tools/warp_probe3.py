"""Broader search for cuDNN's bilinear sampler arithmetic (cudnnSpatialTfSamplerForward)."""
import itertools, os, sys
import numpy as np, torch
import torch.nn.functional as F
f32 = np.float32
DEV = "cuda:0"
rng = np.random.default_rng(5)
H, W, C = 96, 128, 1
img = rng.random((1, C, H, W)).astype(f32)
gx = (rng.random((H, W)) * 1.9 - 0.95).astype(f32)
gy = (rng.random((H, W)) * 1.9 - 0.95).astype(f32)
ti = torch.from_numpy(img).to(DEV)
vgrid = torch.stack([torch.from_numpy(gx), torch.from_numpy(gy)], -1)[None].to(DEV)
gs = F.grid_sample(ti, vgrid, align_corners=True).cpu().numpy()[0, 0]
def fma(a, b, c): return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)
def U(g, size, kind):
    s1 = f32(size - 1)
    if kind == "U1": return ((g + f32(1)) / f32(2)) * s1
    if kind == "U3": return fma(g, np.full_like(g, s1 / 2), np.full_like(g, s1 / 2))
    if kind == "U4": return g * f32(s1 / 2) + f32(s1 / 2)
    if kind == "U7": return (g + f32(1)) * s1 * f32(0.5)
    if kind == "U8": return (g * s1 + s1) * f32(0.5)
    if kind == "U9": return fma(g, np.full_like(g, s1), np.full_like(g, s1)) * f32(0.5)
res = []
for un in ("U1", "U3", "U4", "U7", "U8", "U9"):
    ix, iy = U(gx, W, un), U(gy, H, un)
    fx, fy = np.floor(ix), np.floor(iy)
    x0, y0 = fx.astype(int), fy.astype(int)
    def g(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return np.where(ok, img[0, 0][np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0)).astype(f32)
    v = {"nw": g(y0, x0), "ne": g(y0, x0 + 1), "sw": g(y0 + 1, x0), "se": g(y0 + 1, x0 + 1)}
    dx, dy = ix - fx, iy - fy
    ex, ey = f32(1) - dx, f32(1) - dy
    wsets = {"prod": {"nw": ex * ey, "ne": dx * ey, "sw": ex * dy, "se": dx * dy}}
    se = dx * dy
    wsets["sub"] = {"se": se, "ne": dx - se, "sw": dy - se, "nw": ((f32(1) - dx) - dy) + se}
    wsets["sub2"] = {"se": se, "ne": dx - se, "sw": dy - se, "nw": (f32(1) - dx) - (dy - se)}
    wsets["subfma"] = {"se": se, "ne": fma(-dx, dy, dx), "sw": fma(-dx, dy, dy), "nw": fma(-ex, dy, ex)}
    for wn, wt in wsets.items():
        for order in itertools.permutations(["nw", "ne", "sw", "se"]):
            a, b, c, d = order
            o1 = fma(v[d], wt[d], fma(v[c], wt[c], fma(v[b], wt[b], v[a] * wt[a])))
            o2 = ((v[a] * wt[a] + v[b] * wt[b]) + v[c] * wt[c]) + v[d] * wt[d]
            res.append((int((o1 != gs).sum()), un, wn, "fma", order))
            res.append((int((o2 != gs).sum()), un, wn, "nofma", order))
        for (a, b), (c, d) in [(("nw", "ne"), ("sw", "se")), (("nw", "sw"), ("ne", "se")), (("nw", "se"), ("ne", "sw"))]:
            res.append((int(((fma(v[b], wt[b], v[a] * wt[a]) + fma(v[d], wt[d], v[c] * wt[c])) != gs).sum()), un, wn, "pairfma", (a, b, c, d)))
            res.append((int((((v[a] * wt[a] + v[b] * wt[b]) + (v[c] * wt[c] + v[d] * wt[d])) != gs).sum()), un, wn, "pair", (a, b, c, d)))
    # separable forms
    top = fma(dx, v["ne"], ex * v["nw"]); bot = fma(dx, v["se"], ex * v["sw"])
    res.append((int((fma(dy, bot, ey * top) != gs).sum()), un, "sepx", "fma", ()))
    top = ex * v["nw"] + dx * v["ne"]; bot = ex * v["sw"] + dx * v["se"]
    res.append((int(((ey * top + dy * bot) != gs).sum()), un, "sepx", "nofma", ()))
    res.append((int((fma(dy, bot, ey * top) != gs).sum()), un, "sepx", "mixed", ()))
    top = fma(dx, v["ne"] - v["nw"], v["nw"]); bot = fma(dx, v["se"] - v["sw"], v["sw"])
    res.append((int((fma(dy, bot - top, top) != gs).sum()), un, "lerpx", "fma", ()))
    l = fma(dy, v["sw"] - v["nw"], v["nw"]); r = fma(dy, v["se"] - v["ne"], v["ne"])
    res.append((int((fma(dx, r - l, l) != gs).sum()), un, "lerpy", "fma", ()))
    l = fma(dy, v["sw"], ey * v["nw"]); r = fma(dy, v["se"], ey * v["ne"])
    res.append((int((fma(dx, r, ex * l) != gs).sum()), un, "sepy", "fma", ()))
print("total", gs.size)
for r in sorted(res, key=lambda r: r[0])[:12]:
    print(r)
