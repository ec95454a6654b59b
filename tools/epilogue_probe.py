"""Dev aid: rmnet_mask_epilogue_forward vs the reference's torch composition (models/rmnet.py:368-380, :289-302, :450) on CUDA."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth, oracle
from rmnet_b200 import ops
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import _torch_mask_epilogue
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, K, H, W in ((3, 11, 480, 854), (5, 11, 480, 854), (10, 11, 720, 1280)):
    rng = np.random.default_rng(5)
    x = torch.from_numpy(synth.decoder_logits(rng, n, H, W)).to(dev)
    modes = [0] * K
    Hp, Wp = x.shape[-2:]
    byts = 8 * n * Hp * Wp + 2 * 4 * K * H * W
    for name, fn in (("ours (logit + est_mask)", lambda: ops.mask_epilogue(x, K, (H, W), modes, None)),
                     ("ours (est_mask only)", lambda: ops.mask_epilogue(x, K, (H, W), modes, None, want_logit=False)),
                     ("torch ops (reference)", lambda: _torch_mask_epilogue(x, K, H, W, modes, None))):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        t = float(np.median(ts))
        print(f"n={n} K={K} {H}x{W}: {name:26s} {t:8.1f} us   ({byts / t / 1e3:7.1f} GB/s of {byts / 1e6:.1f} MB algorithmic)")
