"""GPU parity tests (pytest -m gpu) at the HEADLINE sizes and in the large-score regime.

  * C3 (480x854, 5 objects, T = 20) and C4 (720x1280, 10 objects, T = 40): RegionalMemory.step against the reference's
    own composition of the frame step on this GPU (tests/ref_composition.py: torch's CUDA ops + the UNMODIFIED reference
    CUDA kernel for both get_att_map calls) -- bounding boxes bit-exact, mem_val within TOL_STRICT.  At C4 the reference's
    reader runs one object at a time (its p is 2 GB per object).
  * score magnitude: the dense reader against the fp64 oracle for key scales that give |scaled score| from ~6 to ~600
    (default-init RMNet weights give 230-640), with the tolerance each regime actually meets.
  * bank overflow is reported, not silent; the NumPy drop-in of update_optical_flow works in a forked worker of a
    process that has CUDA initialised.
"""
import glob
import math
import os
import sys

import numpy as np
import pytest
import torch

import oracle
import rmnet_b200
import synth
from rmnet_b200 import ops

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"
TOL_STRICT = 2e-4     # on mem_val, values O(1)


def _ref_generator():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "reg_att_map_generator*.so"))
    if not so:
        return None
    if os.path.dirname(so[0]) not in sys.path:
        sys.path.insert(0, os.path.dirname(so[0]))
    try:
        import reg_att_map_generator
    except (ImportError, OSError):
        return None
    if not hasattr(reg_att_map_generator, "forward") or "oracle" not in (getattr(reg_att_map_generator, "__file__", "") or ""):
        return None
    return reg_att_map_generator


def _gpu_frame(g, n, K, H, W, h, w, t):
    """One synthetic frame generated on the device (SURVEY 8d shapes): drifting rectangles -> soft masks, N(0,2px) flow,
    N(0, 0.5) keys, N(0,1) values."""
    lab = torch.zeros((H, W), dtype=torch.long, device=DEV)
    for o in range(1, n + 1):
        bh, bw = int(H * (0.15 + 0.03 * o)), int(W * (0.15 + 0.025 * o))
        y0 = (37 * o + 3 * t) % (H - bh)
        x0 = (91 * o + 5 * t) % (W - bw)
        lab[y0:y0 + bh, x0:x0 + bw] = o
    logits = torch.randn((K, H, W), device=DEV, generator=g)
    logits += 8.0 * torch.nn.functional.one_hot(lab, K).permute(2, 0, 1)
    return dict(mask=torch.softmax(logits, 0).contiguous(),
                flow=(torch.randn((2, H, W), device=DEV, generator=g) * 2.0).contiguous(),
                k4=torch.randn((n, 128, h, w), device=DEV, generator=g) * 0.5,
                v4=torch.randn((n, 512, h, w), device=DEV, generator=g),
                qk=torch.randn((128, h, w), device=DEV, generator=g) * 0.5,
                qv=torch.randn((512, h, w), device=DEV, generator=g))


@pytest.mark.parametrize("name,H,W,n,T,per_object", [("c3_480p_5obj_T20", 480, 854, 5, 20, False), ("c4_720p_10obj_T40", 720, 1280, 10, 40, True)],
                         ids=["c3", "c4"])
def test_step_matches_the_reference_composition_at_headline_size(name, H, W, n, T, per_object):
    gen = _ref_generator()
    if gen is None:
        pytest.skip("oracle/_ref/reg_att_map_generator*.so not built (make -C oracle ref)")
    from ref_composition import ReferenceClip
    K = 11
    h, w = (H + 15) // 16, (W + 15) // 16
    g = torch.Generator(device=DEV).manual_seed(1000 + n)
    torch.cuda.empty_cache()
    ref = ReferenceClip(gen, n, K, H, W, per_object_reader=per_object)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T + 1, device=DEV)
    for t in range(T - 1):
        f = _gpu_frame(g, n, K, H, W, h, w, t)
        ref.commit(f)
        rm.memorize(f["k4"], f["v4"], f["mask"][None], commit=True)
    worst = 0.0
    for t, commit in ((T - 1, True), (T, False)):       # the second step reads T + 1 frames (T committed + the temporary one)
        cur = _gpu_frame(g, n, K, H, W, h, w, t)
        m_ref, pb_ref, cb_ref = ref.step(cur)
        m, pb, cb = rm.step(cur["k4"], cur["v4"], cur["mask"][None], cur["flow"][None], cur["qk"], cur["qv"], commit=commit)
        assert torch.equal(pb, pb_ref) and torch.equal(cb, cb_ref), "bounding boxes differ"
        assert torch.equal(m[:, 512:], m_ref[:, 512:]), "q_val passthrough is not bit-exact"
        err = float((m - m_ref).abs().max())
        worst = max(worst, err)
        assert err <= TOL_STRICT, f"{name}: mem_val differs from the reference composition by {err:.2e}"
        if commit:
            ref.commit(cur)
    st = rm.bank.stats()
    print(f"\n[{name}] max |mem_val - reference composition| = {worst:.2e}; in-region fraction {float(st[:, 0].sum()) / (n * T * h * w):.2f}")
    del ref, rm
    torch.cuda.empty_cache()


@pytest.mark.parametrize("scale,tol", [(0.35, 2e-6), (1.0, 5e-6), (2.0, 6e-5), (3.5, 1.2e-4), (6.0, 3e-4), (10.0, 1e-3)],
                         ids=["score~1", "score~5", "score~22", "score~66", "score~194", "score~540"])
def test_dense_reader_vs_fp64_oracle_across_score_magnitudes(scale, tol):
    """Strict (fp16 hi/lo planes x3) reader vs the fp64 oracle when the scaled scores k.q/sqrt(128) grow from O(1) to
    O(500): the operands carry 22 mantissa bits, so the score error grows with |score|.  `tol` is the max-abs
    bound on mem_val each regime is held to (values are N(0,1)) -- about 2.5x what was measured on the B200 (2.3e-7,
    8.1e-7, 1.8e-5, 3.6e-5, 1.1e-4 for scales 0.35 .. 6), which is also what the reference's own fp32 reader is away from
    fp64 (printed).  north_star's bound is 1e-3 on the LOGIT map; the real-decoder test (tests/test_gpu_rmnet.py) checks
    that at |score| ~ 640."""
    n, T, h, w = 2, 4, 30, 54
    ins = synth.memory_read_inputs(4242, n, T, h, w, scale)
    ref64, _ = oracle.memory_read(*ins, dtype=np.float64)
    ref32, _ = oracle.memory_read(*ins, dtype=np.float32)
    smax = float(np.abs(np.einsum("ncm,ncq->nmq", ins[0].reshape(n, 128, -1)[:, :, :2048], ins[2].reshape(n, 128, -1))).max()) / math.sqrt(128)
    got = ops.memory_reader_forward(*(torch.from_numpy(x).to(DEV) for x in ins)).cpu().numpy()
    err = float(np.abs(got[:, :512] - ref64[:, :512]).max())
    floor = float(np.abs(ref32[:, :512] - ref64[:, :512]).max())
    print(f"\n[key scale {scale}] max |scaled score| ~ {smax:.0f}: max-abs vs fp64 {err:.2e} (fp32 reference reader vs fp64: {floor:.2e})")
    assert err <= tol


def test_frame_step_is_bit_reproducible_from_run_to_run():
    """Two fresh banks, the same frames: mem_val and the boxes must be bit-identical (the reference pins cuDNN to
    deterministic algorithms, runner.py:73-74).  The per-channel value sums behind the uniform rows are accumulated with
    integer atomics on a fixed-point image (associative), everything else has a fixed order."""
    import bench
    wl = bench.WORKLOADS["c2"]
    n, T, H, W = wl["n"], 3, wl["H"], wl["W"]
    pool = bench.make_pool(dict(wl, T=T), 77, 2)
    fr = [{k: torch.from_numpy(v).to(DEV) for k, v in f.items()} for f in pool["frames"]]
    outs = []
    for rep in range(3):
        rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T + 1, device=DEV)
        for t in range(T - 1):
            rm.memorize(fr[t]["k4"], fr[t]["v4"], fr[t]["mask"][None], commit=True)
        res = []
        for t, commit in ((T - 1, True), (T, False)):
            c = fr[t]
            m, pb, cb = rm.step(c["k4"], c["v4"], c["mask"][None], c["flow"][None], c["qk"], c["qv"], commit=commit)
            res.append((m.clone(), pb.clone(), cb.clone()))
        outs.append(res)
    for rep in outs[1:]:
        for (m, pb, cb), (m0, pb0, cb0) in zip(rep, outs[0]):
            assert torch.equal(pb, pb0) and torch.equal(cb, cb0)
            assert torch.equal(m, m0), f"mem_val differs between runs by {float((m - m0).abs().max()):.2e}"


def test_bank_overflow_is_reported_by_stats():
    """A frame that does not fit behind the committed cells is dropped and flagged (bank.cu META_OVERFLOW); the Python
    wrappers guard the frame count, so force it through the bookkeeping and check that stats() raises."""
    n, h, w = 1, 8, 8
    bank = ops.MemoryBank(n, h, w, 1, DEV)
    dense = torch.tensor([[0, w - 1, 0, h - 1]], dtype=torch.int32, device=DEV)
    k, v = torch.randn((n, 128, h, w), device=DEV), torch.randn((n, 512, h, w), device=DEV)
    bank.memorize(k, v, dense, commit=True)
    assert bank.stats()[0, 6] == 0
    bank.frames_committed = 0                      # bypass the host-side guard: the slot is full on the device
    bank.memorize(k, v, dense, commit=False)
    with pytest.raises(RuntimeError, match="overflow"):
        bank.stats()
    bank.reset()
    assert bank.stats()[0, 6] == 0


class _FlowDataset:
    def __len__(self):
        return 2

    def __getitem__(self, i):
        rng = np.random.default_rng(70 + i)
        of = rng.standard_normal((48, 64, 2)).astype(np.float32) * 3
        m1, m2 = synth.affine_pair(rng)
        return rmnet_b200.update_optical_flow(of, m1, m2), of, np.stack([m1, m2])


def test_update_optical_flow_in_a_forked_worker_of_a_cuda_process():
    """utils/data_transforms.py:293-302 calls the op inside DataLoader workers forked from a process that already holds a
    CUDA context: the drop-in must take its plain-C path there (a forked child cannot use CUDA) and stay bit-exact; in
    the parent, with CUDA initialised, it runs on the GPU -- same bits."""
    torch.zeros(1, device=DEV)                     # CUDA is initialised in this (parent) process
    loader = torch.utils.data.DataLoader(_FlowDataset(), batch_size=1, num_workers=2, multiprocessing_context="fork")
    for out, of, ms in loader:
        of, m1, m2 = of[0].numpy(), ms[0, 0].numpy(), ms[0, 1].numpy()
        np.testing.assert_array_equal(out[0].numpy(), oracle.update_optical_flow(of, m1, m2))
        np.testing.assert_array_equal(rmnet_b200.update_optical_flow(of, m1, m2), out[0].numpy())          # GPU path in the parent
        np.testing.assert_array_equal(rmnet_b200.update_optical_flow(of, m1, m2, device="cpu"), out[0].numpy())
