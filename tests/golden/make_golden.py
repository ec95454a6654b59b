"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE in the build container.

Run from the repo root:  python tests/golden/make_golden.py   (needs /root/reference and `make -C oracle ref`)

The reference's own tests hold no golden vector for this path, so these fixtures are the
pin: outputs of models/rmnet.py (MemoryReader.forward, RMNet.warp), utils/helpers.py (pad_divide_by),
torch's F.interpolate as called at models/rmnet.py:245/:356, and the reference NumPy extension built
from extensions/flow_affine_transformation/flow_affine_transformation.cpp.  Inputs are regenerated
from seeds by tests/synth.py; only seeds, shapes, input checksums and reference OUTPUTS are stored.
The reference CUDA extension (reg_att_map_generator) cannot execute without a GPU; it is pinned live
on the GPU box from oracle/_ref instead (tests/test_gpu_parity.py).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RMNET_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
# import-time stand-in for the CUDA-only extension module (never called here)
sys.modules.setdefault("reg_att_map_generator", types.ModuleType("reg_att_map_generator"))

import synth  # noqa: E402
import utils.helpers as ref_helpers  # noqa: E402  (reference)
from models.rmnet import MemoryReader, RMNet  # noqa: E402  (reference)
import flow_affine_transformation as ref_fat  # noqa: E402  (reference, built into oracle/_ref)

torch.set_num_threads(1)


def csum(*arrs):
    return np.array([float(np.asarray(a, np.float64).sum()) for a in arrs])


def golden_memory_read():
    cases = []
    for seed, n, T, h, w, scale in [(11, 2, 3, 5, 7, 1.0), (12, 1, 1, 4, 6, 3.0), (13, 3, 2, 6, 9, 0.3)]:
        mk, mv, qk, qv = synth.memory_read_inputs(seed, n, T, h, w, scale)
        mem_val, p = MemoryReader()(*(torch.from_numpy(a) for a in (mk, mv, qk, qv)))
        cases.append(dict(seed=seed, n=n, T=T, h=h, w=w, scale=scale, insum=csum(mk, mv, qk, qv),
                          mem=mem_val[:, :synth.CV].numpy(), p=p.numpy()))
    np.savez_compressed(os.path.join(HERE, "memory_read.npz"),
                        **{f"c{i}_{k}": v for i, c in enumerate(cases) for k, v in c.items()}, n_cases=len(cases))


def golden_regional_read():
    """models/rmnet.py:245-248 + :355-361 composed with the reference's own ops."""
    seed, n, T, H, W = 21, 2, 3, 96, 144
    h, w = H // 16, W // 16
    rng = np.random.default_rng(seed)
    mk, mv, qk, qv = synth.memory_read_inputs(seed + 1, n, T, h, w, 1.0)
    boxes_m = np.zeros((n, T, 4), np.int32)
    att_m = np.zeros((n, T, H, W), np.float32)
    for o in range(n):
        for t in range(T):
            x0, y0 = int(rng.integers(0, W // 2)), int(rng.integers(0, H // 2))
            x1, y1 = int(rng.integers(x0, W)), int(rng.integers(y0, H))
            boxes_m[o, t] = (x0, x1, y0, y1)
            att_m[o, t, y0:y1 + 1, x0:x1 + 1] = 1
    boxes_q = np.zeros((n, 4), np.int32)
    att_q = np.zeros((n, 1, H, W), np.float32)
    for o in range(n):
        x0, y0 = int(rng.integers(0, W // 2)), int(rng.integers(0, H // 2))
        x1, y1 = int(rng.integers(x0, W)), int(rng.integers(y0, H))
        boxes_q[o] = (x0, x1, y0, y1)
        att_q[o, 0, y0:y1 + 1, x0:x1 + 1] = 1
    a16m = F.interpolate(torch.from_numpy(att_m), scale_factor=1 / 16)          # :245
    k = torch.from_numpy(mk) * a16m[:, None]                                    # :247
    v = torch.from_numpy(mv) * a16m[:, None]                                    # :248
    a16q = F.interpolate(torch.from_numpy(att_q), scale_factor=1 / 16)          # :356
    # one query frame expanded over objects (:332-333)
    qk1, qv1 = torch.from_numpy(qk[0]), torch.from_numpy(qv[0])
    k4e = qk1.expand(n, -1, -1, -1) * a16q                                      # :357
    v4e = qv1.expand(n, -1, -1, -1) * a16q                                      # :358
    mem_val, _ = MemoryReader()(k.contiguous(), v.contiguous(), k4e.contiguous(), v4e.contiguous())  # :361
    np.savez_compressed(os.path.join(HERE, "regional_read.npz"), seed=seed, n=n, T=T, H=H, W=W,
                        boxes_m=boxes_m, boxes_q=boxes_q, att16_m=a16m.numpy(), att16_q=a16q.numpy(),
                        insum=csum(mk, mv, qk, qv), mem_val=mem_val.numpy())


def golden_warp():
    out = {}
    cases = [(31, 3, 40, 56, 2.0, False, "onehot"), (32, 4, 33, 47, 6.0, True, "onehot"),
             (33, 3, 48, 64, 1.0, False, "soft")]
    for i, (seed, K, H, W, sigma, half, kind) in enumerate(cases):
        rng = np.random.default_rng(seed)
        lab = synth.rect_label_map(rng, K - 1, H, W)
        img = synth.onehot(lab, K) if kind == "onehot" else synth.soft_masks(rng, lab, K)
        flow = synth.flow_field(rng, H, W, sigma, half)
        img1, mask = RMNet.warp(None, torch.from_numpy(img[None]), torch.from_numpy(flow[None]))
        out.update({f"c{i}_seed": seed, f"c{i}_K": K, f"c{i}_H": H, f"c{i}_W": W, f"c{i}_sigma": sigma,
                    f"c{i}_half": half, f"c{i}_kind": kind, f"c{i}_insum": csum(img, flow),
                    f"c{i}_img1": img1[0].numpy(), f"c{i}_valid": mask[0, 0].numpy().astype(np.uint8)})
    np.savez_compressed(os.path.join(HERE, "warp.npz"), n_cases=len(cases), **out)


def golden_pad_downsample():
    sizes = [(480, 854), (240, 432), (720, 1280), (100, 70), (33, 47), (16, 16), (481, 865)]
    pads, a16 = [], {}
    for i, (H, W) in enumerate(sizes):
        x = torch.arange(H * W, dtype=torch.float32).view(1, 1, H, W) + 1
        (xp,), pad = ref_helpers.pad_divide_by([x], 16, (H, W))
        pads.append(pad)
        a16[f"ds{i}"] = F.interpolate(xp, scale_factor=1 / 16)[0, 0].numpy()   # models/rmnet.py:245 on a ramp
    np.savez_compressed(os.path.join(HERE, "pad_downsample.npz"), sizes=np.array(sizes), pads=np.array(pads), **a16)


def golden_flow_affine():
    out = {}
    cases = [(41, 48, 64, 3.0), (42, 37, 53, 10.0), (43, 120, 90, 0.5)]
    for i, (seed, H, W, sigma) in enumerate(cases):
        rng = np.random.default_rng(seed)
        of = np.ascontiguousarray(np.moveaxis(synth.flow_field(rng, H, W, sigma), 0, -1))
        m1, m2 = synth.affine_pair(rng)
        res = ref_fat.update_optical_flow(of, m1, m2)
        out.update({f"c{i}_seed": seed, f"c{i}_H": H, f"c{i}_W": W, f"c{i}_sigma": sigma,
                    f"c{i}_insum": csum(of, m1, m2), f"c{i}_out": res.astype(np.float32)})
    np.savez_compressed(os.path.join(HERE, "flow_affine.npz"), n_cases=len(cases), **out)


def golden_mask_epilogue():
    """The tail after the decoder: models/rmnet.py:368-370 (2-class softmax), the reference's own soft_aggregation
    (:289-302), the un-pad (:376-380), the channel overrides (:436-448, restated with the same torch expressions: they
    are inline in RMNet.forward) and the final softmax (:450)."""
    out = {}
    #        seed  n  K   H    W   modes (0 keep, 1 absent, 2 new)
    cases = [(51, 2, 4, 40, 56, None), (52, 3, 11, 33, 47, [0, 0, 0, 2, 1] + [0] * 6), (53, 5, 11, 48, 70, [0, 1, 0, 0, 2, 0] + [0] * 5)]
    for i, (seed, n, K, H, W, modes) in enumerate(cases):
        rng = np.random.default_rng(seed)
        x = synth.decoder_logits(rng, n, H, W)
        new_mask = (synth.onehot(synth.rect_label_map(rng, K - 1, H, W), K)).astype(np.int32)
        logits = torch.from_numpy(x)
        ps = F.softmax(logits, dim=1)[:, 1]                                     # :368-370
        logit = RMNet.soft_aggregation(None, ps, K, [n])                        # :373 -> :289-302 (reference code)
        (_,), pad = ref_helpers.pad_divide_by([torch.zeros(1, 1, H, W)], 16, (H, W))
        if pad[2] + pad[3] > 0:
            logit = logit[:, :, pad[2]:-pad[3], :]                              # :376-377
        if pad[0] + pad[1] > 0:
            logit = logit[:, :, :, pad[0]:-pad[1]]                              # :379-380
        logit = logit.clone()
        masks_t = torch.from_numpy(new_mask)
        for j in range(K):
            if modes is not None and modes[j] == 2:
                logit[0, j] = masks_t[j].float() * 32.0605 - 16.1181            # :442
            if modes is not None and modes[j] == 1:
                logit[0, j] = -16.1181                                          # :448
        est = F.softmax(logit, dim=1)                                           # :450
        out.update({f"c{i}_seed": seed, f"c{i}_n": n, f"c{i}_K": K, f"c{i}_H": H, f"c{i}_W": W,
                    f"c{i}_modes": np.array(modes if modes is not None else [0] * K), f"c{i}_insum": csum(x, new_mask),
                    f"c{i}_logit": logit.numpy(), f"c{i}_est": est.numpy()})
    np.savez_compressed(os.path.join(HERE, "mask_epilogue.npz"), n_cases=len(cases), **out)


if __name__ == "__main__":
    golden_mask_epilogue()
    golden_memory_read()
    golden_regional_read()
    golden_warp()
    golden_pad_downsample()
    golden_flow_affine()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
