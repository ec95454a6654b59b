import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from rmnet_b200 import ops
from test_gpu_parity import _torch_warp, _warp_cases
f32 = np.float32
name, img, flow = _warp_cases()[5]
ti, tf = torch.from_numpy(img).to("cuda:0"), torch.from_numpy(flow).to("cuda:0")
ref1, refm = _torch_warp(ti, tf)
mine, mv = ops.warp(ti, tf)
r, m = ref1.cpu().numpy(), mine.cpu().numpy()
bad = np.argwhere(m != r)
print(name, "mismatches", len(bad), "of", r.size, " valid mismatches", int((mv != refm).sum().item()))
B, C, H, W = img.shape
xs = np.arange(W, dtype=f32)[None, :].repeat(H, 0); ys = np.arange(H, dtype=f32)[:, None].repeat(W, 1)
vx, vy = xs + flow[0, 0], ys + flow[0, 1]
gx = (f32(2) * vx) * (f32(1) / f32(W - 1)) - f32(1); gy = (f32(2) * vy) * (f32(1) / f32(H - 1)) - f32(1)
ix = ((gx + f32(1)) / f32(2)) * f32(W - 1); iy = ((gy + f32(1)) / f32(2)) * f32(H - 1)
for (b, c, y, x) in bad[:16]:
    print(f"  c={c} y={y} x={x} ix={ix[y,x]!r} iy={iy[y,x]!r} torch={r[b,c,y,x]!r} mine={m[b,c,y,x]!r} valid={refm[b,c,y,x].item()}")
