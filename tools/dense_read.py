"""Dense C3-shaped read (5 objects, T=20, 480p) for profiling the steady state of the tcgen05 kernel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmnet_b200
from rmnet_b200 import ops
dev = "cuda:0"
n, T, h, w = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (5, 20, 30, 54)))
g = torch.Generator(device=dev).manual_seed(1)
bank = ops.MemoryBank(n, h, w, T, dev)
dense = torch.tensor([[0, w - 1, 0, h - 1]] * n, dtype=torch.int32, device=dev)
for t in range(T):
    bank.memorize(torch.randn((n, 128, h, w), device=dev, generator=g) * 0.5, torch.randn((n, 512, h, w), device=dev, generator=g), dense, commit=True)
qk = torch.randn((128, h, w), device=dev, generator=g) * 0.5
qv = torch.randn((512, h, w), device=dev, generator=g)
for _ in range(3):
    out = bank.read(qk, qv, dense, n)
torch.cuda.synchronize()
print("ok", float(out.sum()))
import math, numpy as np
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
N, M = h * w, T * h * w
flops = 1280.0 * M * N * n
def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))
t_att = timed(lambda: bank.read(qk, qv, dense, n, stages=1, out=out))
t_mrg = timed(lambda: bank.read(qk, qv, dense, n, stages=2, out=out))
t_all = timed(lambda: bank.read(qk, qv, dense, n, out=out))
print(f"dense n={n} T={T} {h}x{w}: attention {t_att:.1f} us = {3 * flops / t_att / 1e6:.0f} TF/s-equivalent (3 passes), merge {t_mrg:.1f} us, whole read {t_all:.1f} us")
# the reference's MemoryReader composition (models/rmnet.py:147-165) with torch's CUDA ops on the same GPU
mk = torch.randn((n, 128, T, h, w), device=dev, generator=g) * 0.5
mv = torch.randn((n, 512, T, h, w), device=dev, generator=g)
qk4, qv4 = qk[None].expand(n, -1, -1, -1).contiguous(), qv[None].expand(n, -1, -1, -1).contiguous()
def ref():
    mi = torch.transpose(mk.view(n, 128, M), 1, 2); qi = qk4.view(n, 128, N)
    p = torch.softmax(torch.bmm(mi, qi) / math.sqrt(128), dim=1)
    return torch.cat([torch.bmm(mv.view(n, 512, M), p).view(n, 512, h, w), qv4], dim=1)
for _ in range(2): ref()
print(f"reference MemoryReader ops on this GPU (fp32 cuBLAS + ATen): {timed(ref, 5):.0f} us")
