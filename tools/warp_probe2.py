"""Diagnostic: identify the arithmetic of the bilinear sampler torch dispatches to on CUDA for
grid_sample(bilinear, zeros, align_corners=True) -- cudnnSpatialTfSamplerForward when cuDNN is enabled."""
import itertools, os, sys
import numpy as np, torch
import torch.nn.functional as F
f32 = np.float32
DEV = "cuda:0"
rng = np.random.default_rng(5)
H, W, C = 96, 128, 2
img = rng.random((1, C, H, W)).astype(f32)
gx = (rng.random((H, W)) * 2.2 - 1.1).astype(f32)
gy = (rng.random((H, W)) * 2.2 - 1.1).astype(f32)
ti = torch.from_numpy(img).to(DEV)
vgrid = torch.stack([torch.from_numpy(gx), torch.from_numpy(gy)], -1)[None].to(DEV)
gs_cudnn = F.grid_sample(ti, vgrid, align_corners=True).cpu().numpy()[0]
torch.backends.cudnn.enabled = False
gs_native = F.grid_sample(ti, vgrid, align_corners=True).cpu().numpy()[0]
torch.backends.cudnn.enabled = True
print("cudnn vs native differ:", int((gs_cudnn != gs_native).sum()), "of", gs_native.size)

def fma(a, b, c): return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)
def unnorm(g, size, kind):
    s1 = f32(size - 1)
    if kind == "U1": return ((g + f32(1)) / f32(2)) * s1
    if kind == "U3": return fma(g, np.full_like(g, s1 / 2), np.full_like(g, s1 / 2))
    if kind == "U4": return g * f32(s1 / 2) + f32(s1 / 2)
    if kind == "U6": return (g + f32(1)) * f32(s1 / 2)
def taps(ix, iy, c):
    fx, fy = np.floor(ix), np.floor(iy)
    x0, y0 = fx.astype(int), fy.astype(int)
    def g(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return np.where(ok, img[0, c][np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0)).astype(f32)
    return fx, fy, g(y0, x0), g(y0, x0 + 1), g(y0 + 1, x0), g(y0 + 1, x0 + 1)
def accumulate(kind, ix, iy, c):
    fx, fy, a, b, cc, d = taps(ix, iy, c)
    dx, dy = ix - fx, iy - fy
    ex, ey = f32(1) - dx, f32(1) - dy
    nw, ne, sw, se = ex * ey, dx * ey, ex * dy, dx * dy
    if kind == "chain_fma": return fma(d, se, fma(cc, sw, fma(b, ne, a * nw)))
    if kind == "chain_nofma": return ((a * nw + b * ne) + cc * sw) + d * se
    if kind == "chain_rev_fma": return fma(a, nw, fma(b, ne, fma(cc, sw, d * se)))
    if kind == "pairs_fma": return fma(b, ne, a * nw) + fma(d, se, cc * sw)
    if kind == "pairs_nofma": return (a * nw + b * ne) + (cc * sw + d * se)
    if kind == "lerp_fma":
        top = fma(dx, b - a, a); bot = fma(dx, d - cc, cc); return fma(dy, bot - top, top)
    if kind == "lerp_nofma":
        top = a + dx * (b - a); bot = cc + dx * (d - cc); return top + dy * (bot - top)
    if kind == "sep_x_fma":
        top = fma(dx, b, ex * a); bot = fma(dx, d, ex * cc); return fma(dy, bot, ey * top)
    if kind == "sep_x_nofma":
        top = ex * a + dx * b; bot = ex * cc + dx * d; return ey * top + dy * bot
    if kind == "sep_y_fma":
        l = fma(dy, cc, ey * a); r = fma(dy, d, ey * b); return fma(dx, r, ex * l)
    if kind == "sep_y_nofma":
        l = ey * a + dy * cc; r = ey * b + dy * d; return ex * l + dx * r
    if kind == "lerp_y_fma":
        l = fma(dy, cc - a, a); r = fma(dy, d - b, b); return fma(dx, r - l, l)
    if kind == "lerp_y_nofma":
        l = a + dy * (cc - a); r = b + dy * (d - b); return l + dx * (r - l)
    if kind == "w3_fma":   # weights as ((1-dx)*(1-dy)) but value*wx then *wy
        return fma(d * dx, dy, fma(cc * ex, dy, fma(b * dx, ey, (a * ex) * ey)))
    if kind == "w3_nofma":
        return (((a * ex) * ey + (b * dx) * ey) + (cc * ex) * dy) + (d * dx) * dy
    if kind == "chain_fma_order2": return fma(d, se, fma(b, ne, fma(cc, sw, a * nw)))
    if kind == "chain_fma_order3": return fma(cc, sw, fma(d, se, fma(b, ne, a * nw)))
KINDS = ["chain_fma", "chain_nofma", "chain_rev_fma", "pairs_fma", "pairs_nofma", "lerp_fma", "lerp_nofma", "sep_x_fma",
         "sep_x_nofma", "sep_y_fma", "sep_y_nofma", "lerp_y_fma", "lerp_y_nofma", "w3_fma", "w3_nofma", "chain_fma_order2", "chain_fma_order3"]
res = []
for un in ("U1", "U3", "U4", "U6"):
    ix, iy = unnorm(gx, W, un), unnorm(gy, H, un)
    for k in KINDS:
        o = np.stack([accumulate(k, ix, iy, c) for c in range(C)])
        res.append((int((o != gs_cudnn).sum()), int((o != gs_native).sum()), un, k, float(np.abs(o - gs_cudnn).max())))
for r in sorted(res)[:14]:
    print("cudnn-mismatch %6d native-mismatch %6d  %s %-16s maxdiff %.2e" % r)
print("best for native:", sorted(res, key=lambda r: r[1])[:3])
