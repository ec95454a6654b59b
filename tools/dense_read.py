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
