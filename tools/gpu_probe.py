"""Per-op timing probe on the GPU box (development aid; bench.py is the contract).  Writes gpurun_out/probe.json."""
import glob
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rmnet_b200  # noqa: E402
import synth  # noqa: E402
from rmnet_b200 import ops  # noqa: E402

DEV = "cuda:0"
out = {"gpu": torch.cuda.get_device_name(0), "cpus": os.cpu_count(), "has_reference_tree": os.path.isdir("/root/reference"),
       "umma": ops.umma_available()}
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return {"best_us": min(ts), "median_us": float(np.median(ts))}


def ref_generator():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "reg_att_map_generator*.so"))
    if not so:
        return None
    sys.path.insert(0, os.path.dirname(so[0]))
    import reg_att_map_generator
    return reg_att_map_generator


def torch_reader(m_key, m_val, q_key, q_val):
    B, D_e, T, H, W = m_key.size()
    D_o = m_val.size(1)
    mi = torch.transpose(m_key.view(B, D_e, T * H * W), 1, 2)
    qi = q_key.view(B, D_e, H * W)
    p = torch.bmm(mi, qi) / math.sqrt(D_e)
    p = torch.softmax(p, dim=1)
    mem = torch.bmm(m_val.view(B, D_o, T * H * W), p).view(B, D_o, H, W)
    return torch.cat([mem, q_val], dim=1)


rng = np.random.default_rng(0)
H, W, K = 480, 854, 11
lab = synth.rect_label_map(rng, 5, H, W)
mask = torch.from_numpy(synth.soft_masks(rng, lab, K)[None]).to(DEV)
flow = torch.from_numpy(synth.flow_field(rng, H, W, 3.0)[None]).to(DEV)
maskp = torch.nn.functional.pad(mask, (5, 5, 0, 0)).contiguous()

out["generator_480x864_bbox_only"] = timeit(lambda: ops.reg_att_map_forward(maskp, want_att=False))
out["generator_480x864_with_att"] = timeit(lambda: ops.reg_att_map_forward(maskp))
gen = ref_generator()
if gen is not None:
    out["reference_generator_480x864"] = timeit(lambda: gen.forward(maskp, 0.5, 10, 64))
out["warp_bbox_fused_480x854"] = timeit(lambda: ops.warp_att_map_forward(mask, flow, want_att=False))
out["warp_literal_480x854"] = timeit(lambda: ops.warp(mask, flow))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import _torch_warp  # noqa: E402
out["torch_warp_480x854"] = timeit(lambda: _torch_warp(mask, flow))
of = torch.from_numpy(np.ascontiguousarray(np.moveaxis(synth.flow_field(rng, H, W, 3.0), 0, -1))).to(DEV)
m1, m2 = synth.affine_pair(rng)
out["flow_affine_480x854_device"] = timeit(lambda: ops.update_optical_flow_cuda(of, m1, m2))

for name, (n, T) in {"c2_n3_T5": (3, 5), "c3_n5_T20": (5, 20)}.items():
    h, w = 30, 54
    N = h * w
    g = torch.Generator(device=DEV).manual_seed(1)
    ks = torch.randn((n, 128, T, h, w), device=DEV, generator=g) * 0.5
    vs = torch.randn((n, 512, T, h, w), device=DEV, generator=g)
    qk = torch.randn((128, h, w), device=DEV, generator=g) * 0.5
    qv = torch.randn((512, h, w), device=DEV, generator=g)
    dense = torch.tensor([[0, w - 1, 0, h - 1]] * n, dtype=torch.int32, device=DEV)
    # regional boxes: ~24 % of the cells
    reg = torch.tensor([[10, 10 + 26, 5, 5 + 14]] * n, dtype=torch.int32, device=DEV)
    for label, rect in (("dense", dense), ("regional_f0.25", reg)):
        bank = ops.MemoryBank(n, h, w, T + 1, DEV)
        k_t = [ks[:, :, t].contiguous() for t in range(T)]
        v_t = [vs[:, :, t].contiguous() for t in range(T)]
        for t in range(T):
            bank.memorize(k_t[t], v_t[t], rect, commit=True)
        out[f"pack_one_frame_{name}_{label}"] = timeit(lambda: bank.memorize(k_t[0], v_t[0], rect, commit=False))
        for impl_name, impl in (("simt", rmnet_b200.RMNET_IMPL_SIMT), ("umma", rmnet_b200.RMNET_IMPL_UMMA)):
            if impl == rmnet_b200.RMNET_IMPL_UMMA and not ops.umma_available():
                continue
            for pname, prec in (("split3", 0), ("single", 1)):
                out[f"read_{name}_{label}_{impl_name}_{pname}"] = timeit(
                    lambda: bank.read(qk, qv, rect, n, precision=prec, impl=impl), iters=5)
        del bank
    qk_n = qk[None].expand(n, -1, -1, -1).contiguous()
    qv_n = qv[None].expand(n, -1, -1, -1).contiguous()
    out[f"torch_reference_reader_{name}"] = timeit(lambda: torch_reader(ks, vs, qk_n, qv_n), iters=5)
    out[f"literal_reader_{name}_simt"] = timeit(lambda: ops.memory_reader_forward(ks, vs, qk_n, qv_n, impl=rmnet_b200.RMNET_IMPL_SIMT), iters=3)
    t0 = time.time()
    cpu_in = [x.cpu() for x in (ks, vs, qk_n, qv_n)]
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    torch_reader(*cpu_in)
    out[f"torch_cpu_reader_{name}_s"] = time.time() - t0

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
