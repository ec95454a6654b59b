"""CPU tests: pin the oracle (oracle/) against the golden vectors produced by the reference
(tests/golden/make_golden.py) and against the unmodified reference extension in oracle/_ref."""
import glob
import os
import sys

import numpy as np
import pytest

import oracle
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _csum(*arrs):
    return np.array([float(np.asarray(a, np.float64).sum()) for a in arrs])


def test_memory_read_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "memory_read.npz")
    for i in range(int(g["n_cases"])):
        a = {k: g[f"c{i}_{k}"] for k in ("seed", "n", "T", "h", "w", "scale", "insum", "mem", "p")}
        ins = synth.memory_read_inputs(int(a["seed"]), int(a["n"]), int(a["T"]), int(a["h"]), int(a["w"]), float(a["scale"]))
        np.testing.assert_allclose(_csum(*ins), a["insum"], rtol=0, atol=0)  # inputs regenerate bit-identically
        mem32, p32 = oracle.memory_read(*ins, dtype=np.float32, want_p=True)
        mem64, _ = oracle.memory_read(*ins, dtype=np.float64)
        memc, pc = oracle.memory_read_f64(*ins, want_p=True)
        ref = a["mem"].reshape(mem32[:, :synth.CV].shape)
        # floating point: the reference itself is fp32 SGEMM; 2e-5 abs on O(1) values is its own noise floor
        for got in (mem32, mem64, memc):
            np.testing.assert_allclose(got[:, :synth.CV], ref, rtol=0, atol=2e-5)
            np.testing.assert_array_equal(got[:, synth.CV:], ins[3])       # q_val passthrough is bit-exact
        np.testing.assert_allclose(p32, a["p"], rtol=1e-4, atol=1e-7)   # exp() of O(50) scores: relative
        np.testing.assert_allclose(pc, a["p"], rtol=1e-4, atol=1e-7)


def test_regional_read_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "regional_read.npz")
    n, T, H, W = int(g["n"]), int(g["T"]), int(g["H"]), int(g["W"])
    h, w = H // 16, W // 16
    mk, mv, qk, qv = synth.memory_read_inputs(int(g["seed"]) + 1, n, T, h, w, 1.0)
    np.testing.assert_array_equal(_csum(mk, mv, qk, qv), g["insum"])
    att_m = np.zeros((n, T, H, W), np.float32)
    for o in range(n):
        for t in range(T):
            x0, x1, y0, y1 = g["boxes_m"][o, t]
            att_m[o, t, y0:y1 + 1, x0:x1 + 1] = 1
    att_q = np.zeros((n, H, W), np.float32)
    for o in range(n):
        x0, x1, y0, y1 = g["boxes_q"][o]
        att_q[o, y0:y1 + 1, x0:x1 + 1] = 1
    # the /16 nearest rule (models/rmnet.py:245,:356) is bit-exact
    np.testing.assert_array_equal(oracle.downsample16(att_m), g["att16_m"])
    np.testing.assert_array_equal(oracle.downsample16(att_q)[:, None], g["att16_q"])
    got = oracle.regional_memory_read(mk, mv, att_m, qk[0], qv[0], att_q)
    np.testing.assert_allclose(got[:, :synth.CV], g["mem_val"][:, :synth.CV], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(got[:, synth.CV:], g["mem_val"][:, synth.CV:])


def test_warp_matches_reference_golden_bit_exact(golden_dir):
    g = _load(golden_dir, "warp.npz")
    for i in range(int(g["n_cases"])):
        seed, K, H, W = (int(g[f"c{i}_{k}"]) for k in ("seed", "K", "H", "W"))
        rng = np.random.default_rng(seed)
        lab = synth.rect_label_map(rng, K - 1, H, W)
        img = synth.onehot(lab, K) if str(g[f"c{i}_kind"]) == "onehot" else synth.soft_masks(rng, lab, K)
        flow = synth.flow_field(rng, H, W, float(g[f"c{i}_sigma"]), bool(g[f"c{i}_half"]))
        np.testing.assert_array_equal(_csum(img, flow), g[f"c{i}_insum"])
        img1, valid = oracle.warp(img[None], flow[None], arith="cpu")   # reference ran on torch's CPU backend
        np.testing.assert_array_equal(img1[0], g[f"c{i}_img1"])
        np.testing.assert_array_equal(valid[0, 0].astype(np.uint8), g[f"c{i}_valid"])
        # the CUDA-arithmetic variant differs by at most a few ulp of the coordinate (reciprocal multiply)
        img1c, _ = oracle.warp(img[None], flow[None], arith="cuda")
        assert np.abs(img1c[0] - g[f"c{i}_img1"]).max() < 1e-3


def test_warp_randomised_bit_exact_against_torch_cpu_ops():
    """oracle.warp(arith='cpu') against RMNet.warp restated with torch's own CPU ops (tests/ref_composition.torch_warp,
    models/rmnet.py:252-278), live, on 24 random cases: odd sizes, B > 1, half-pixel flows, large flows that leave the
    frame, soft and one-hot images.  Bit-exact, like the golden cases."""
    import torch
    from ref_composition import torch_warp
    torch.set_num_threads(1)
    rng = np.random.default_rng(77)
    for case in range(24):
        B, C = int(rng.integers(1, 3)), int(rng.integers(1, 5))
        H, W = int(rng.integers(2, 70)), int(rng.integers(2, 90))
        lab = synth.rect_label_map(rng, max(C - 1, 1), H, W)
        img = np.stack([synth.soft_masks(rng, lab, C) if case % 2 else synth.onehot(lab, C) for _ in range(B)])
        flow = np.stack([synth.flow_field(rng, H, W, float(rng.choice([0.5, 2.0, 15.0])), half_pixel=(case % 3 == 0)) for _ in range(B)])
        ref_img, ref_mask = torch_warp(torch.from_numpy(img), torch.from_numpy(flow))
        img1, valid = oracle.warp(img, flow, arith="cpu")
        np.testing.assert_array_equal(img1, ref_img.numpy(), err_msg=f"case {case} {B}x{C}x{H}x{W}")
        np.testing.assert_array_equal(valid, ref_mask.numpy(), err_msg=f"case {case}")


def test_pad_and_downsample_match_reference_golden(golden_dir):
    g = _load(golden_dir, "pad_downsample.npz")
    for i, (H, W) in enumerate(g["sizes"]):
        assert oracle.pad_amounts(int(H), int(W)) == tuple(int(v) for v in g["pads"][i])
        x = (np.arange(H * W, dtype=np.float32) + 1).reshape(1, H, W)
        xp, pad = oracle.pad_divide_by(x)
        np.testing.assert_array_equal(oracle.downsample16(xp)[0], g[f"ds{i}"])


def test_flow_affine_matches_reference_golden_bit_exact(golden_dir):
    g = _load(golden_dir, "flow_affine.npz")
    for i in range(int(g["n_cases"])):
        seed, H, W = (int(g[f"c{i}_{k}"]) for k in ("seed", "H", "W"))
        rng = np.random.default_rng(seed)
        of = np.ascontiguousarray(np.moveaxis(synth.flow_field(rng, H, W, float(g[f"c{i}_sigma"])), 0, -1))
        m1, m2 = synth.affine_pair(rng)
        np.testing.assert_array_equal(_csum(of, m1, m2), g[f"c{i}_insum"])
        np.testing.assert_array_equal(oracle.update_optical_flow(of, m1, m2), g[f"c{i}_out"])


def test_flow_affine_matches_live_reference_extension():
    """oracle vs the UNMODIFIED reference .so (oracle/_ref, built by `make -C oracle ref`)."""
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "flow_affine_transformation*.so"))
    if not so:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    sys.path.insert(0, os.path.dirname(so[0]))
    import flow_affine_transformation as ref
    rng = np.random.default_rng(5)
    # the reference's own smoke test shape/distribution (extensions/flow_affine_transformation/test.py:15-18) ...
    of = rng.random((480, 640, 2)).astype(np.float32)
    m1, m2 = rng.random((2, 3)).astype(np.float32), rng.random((2, 3)).astype(np.float32)
    np.testing.assert_array_equal(oracle.update_optical_flow(of, m1, m2), ref.update_optical_flow(of, m1, m2))
    # ... and realistic affine pairs, incl. exact .5 ties for round-half-away
    for s in range(6):
        rng = np.random.default_rng(100 + s)
        H, W = int(rng.integers(20, 200)), int(rng.integers(20, 200))
        of = np.ascontiguousarray(np.moveaxis(synth.flow_field(rng, H, W, 4.0, half_pixel=(s % 2 == 0)), 0, -1))
        m1, m2 = synth.affine_pair(rng)
        if s == 0:
            m1 = np.array([[1, 0, 0.5], [0, 1, -0.5]], np.float32)
            m2 = np.array([[1, 0, 0], [0, 1, 0]], np.float32)
        np.testing.assert_array_equal(oracle.update_optical_flow(of, m1, m2), ref.update_optical_flow(of, m1, m2))


def test_generator_semantics_edge_cases():
    """reg_att_map_generator.cu:31-92 edge cases: empty object -> full frame, n_points 9/10,
    loosening clamps at x_min in {63,64,65} and x_max+64 in {W-1,W}, NaN pixels, channel 0 untouched."""
    H, W, K = 80, 200, 6
    m = np.zeros((1, K, H, W), np.float32)
    m[0, 0] = 1.0                                   # background channel: never scanned
    m[0, 1, 10, 63] = m[0, 1, 11, 63:72] = 1.0      # 10 points, x_min = 63 -> 0
    m[0, 2, 70, 64:74] = 1.0                        # x_min = 64 -> 0 (<=), x_max = 73
    m[0, 3, 5, 65:74] = 1.0                         # 9 points -> full frame
    m[0, 4, 40, 65:75] = 0.5                        # == threshold counts; x_min = 65 -> 1
    m[0, 4, 41, 135] = 0.5                          # x_max + 64 = 199 = W-1 -> 199
    m[0, 5, 20:30, 100:136] = np.nan                # NaN never counts
    m[0, 5, 50, 126:137] = 0.75                     # x_max = 136, +64 = 200 >= W -> W-1
    att, bb = oracle.reg_att_map(m)
    np.testing.assert_array_equal(bb[0, 0], [0, 0, 0, 0])
    np.testing.assert_array_equal(bb[0, 1], [0, 135, 0, 75])
    np.testing.assert_array_equal(bb[0, 2], [0, 137, 6, 79])
    np.testing.assert_array_equal(bb[0, 3], [0, W - 1, 0, H - 1])
    np.testing.assert_array_equal(bb[0, 4], [1, 199, 0, 79])
    np.testing.assert_array_equal(bb[0, 5], [62, 199, 0, 79])
    assert att[0, 0].sum() == 0
    for i in range(1, K):
        x0, x1, y0, y1 = bb[0, i]
        ref = np.zeros((H, W), np.float32)
        ref[y0:y1 + 1, x0:x1 + 1] = 1
        np.testing.assert_array_equal(att[0, i], ref)


def test_generator_randomised_against_an_independent_numpy_restatement():
    """oracle.reg_att_map (C) against an independent numpy restatement of reg_att_map_generator.cu:31-92 on 60 random
    masks: sparse / dense foregrounds, NaNs, objects touching the borders, other thresholds and loosening widths."""
    rng = np.random.default_rng(123)

    def numpy_generator(mask, thr, n_pts, loose):
        B, K, H, W = mask.shape
        att = np.zeros_like(mask)
        bb = np.zeros((B, K, 4), np.int32)
        for b in range(B):
            for i in range(1, K):
                with np.errstate(invalid="ignore"):
                    ys, xs = np.nonzero(mask[b, i] >= thr)                     # :42 (NaN compares false)
                if len(xs) < n_pts:                                            # :57-61
                    x0, x1, y0, y1 = 0, W - 1, 0, H - 1
                else:                                                          # :63-74
                    x0 = 0 if xs.min() <= loose else xs.min() - loose
                    x1 = W - 1 if xs.max() + loose >= W else xs.max() + loose
                    y0 = 0 if ys.min() <= loose else ys.min() - loose
                    y1 = H - 1 if ys.max() + loose >= H else ys.max() + loose
                bb[b, i] = (x0, x1, y0, y1)
                att[b, i, y0:y1 + 1, x0:x1 + 1] = 1                            # :81-92
        return att, bb

    for case in range(60):
        B, K = int(rng.integers(1, 3)), int(rng.integers(2, 7))
        H, W = int(rng.integers(8, 150)), int(rng.integers(8, 220))
        thr = float(rng.choice([0.5, 0.3, 0.9]))
        n_pts = int(rng.choice([10, 1, 25]))
        loose = int(rng.choice([64, 0, 7, 300]))
        m = rng.random((B, K, H, W)).astype(np.float32) * float(rng.choice([0.6, 1.0, 1.4]))
        if case % 3 == 0:                                                      # sparse: a few pixels above the threshold
            m *= (rng.random((B, K, H, W)) < 0.002)
        if case % 4 == 1:
            m[rng.random((B, K, H, W)) < 0.01] = np.nan
        if case % 5 == 2:                                                      # an object touching two borders
            m[:, 1, :3, -3:] = 1.0
        att, bb = oracle.reg_att_map(m, thr, n_pts, loose)
        att_ref, bb_ref = numpy_generator(m, thr, n_pts, loose)
        np.testing.assert_array_equal(bb, bb_ref, err_msg=f"case {case}")
        np.testing.assert_array_equal(att, att_ref, err_msg=f"case {case}")


def test_mask_epilogue_matches_reference_golden(golden_dir):
    """oracle.mask_epilogue vs the reference's soft_aggregation + torch ops (models/rmnet.py:368-380, :289-302, :436-450).
    est_mask (what the frame loop stores, :450) is held to 1e-5; the logit map to 1e-3 (north_star) wherever the
    reference's own formula is well conditioned, see synth.epilogue_logit_tolerance for the saturated pixels."""
    g = _load(golden_dir, "mask_epilogue.npz")
    for i in range(int(g["n_cases"])):
        seed, n, K, H, W = (int(g[f"c{i}_{k}"]) for k in ("seed", "n", "K", "H", "W"))
        modes = [int(m) for m in g[f"c{i}_modes"]]
        rng = np.random.default_rng(seed)
        x = synth.decoder_logits(rng, n, H, W)
        new_mask = synth.onehot(synth.rect_label_map(rng, K - 1, H, W), K).astype(np.int32)
        np.testing.assert_allclose(_csum(x, new_mask), g[f"c{i}_insum"], rtol=1e-12)
        tol = synth.epilogue_logit_tolerance(x, K, H, W)
        assert (tol <= 1.1e-3).mean() > 0.25             # a good share of the pixels is held to (essentially) the plain 1e-3
        logit, est = oracle.mask_epilogue(x, K, (H, W), modes, new_mask)
        assert logit.shape == (1, K, H, W) and est.shape == (1, K, H, W)
        assert (np.abs(logit - g[f"c{i}_logit"]) <= tol).all()
        assert np.abs(est - g[f"c{i}_est"]).max() <= 1e-5
        # override channels, the clamp ceiling (15.9424, the reference's own comment at :441) and the floor are exact
        for j, m in enumerate(modes):
            if m != oracle.CH_KEEP:
                np.testing.assert_array_equal(logit[0, j], g[f"c{i}_logit"][0, j])
        assert np.float32(logit.max()) == g[f"c{i}_logit"].max()
        if K > n + 1 and modes[n + 1] == oracle.CH_KEEP:
            assert (logit[0, n + 1] == np.float32(np.log(np.float32(1e-7) / (np.float32(1) - np.float32(1e-7))))).all()


def test_mask_epilogue_and_memory_read_randomised_against_torch_cpu_ops():
    """Live (no stored vectors): oracle.mask_epilogue against the torch composition of models/rmnet.py:368-380, :289-302,
    :436-450 and oracle.memory_read against MemoryReader's ops (:147-165), both on torch's CPU backend, random shapes."""
    import math
    import torch
    from ref_composition import torch_mask_epilogue
    torch.set_num_threads(2)
    rng = np.random.default_rng(131)
    for case in range(8):
        n, K = int(rng.integers(1, 6)), 11
        H, W = int(rng.integers(17, 70)), int(rng.integers(17, 90))
        x = synth.decoder_logits(rng, n, H, W)
        modes = [0] * K
        if case % 2:
            modes[int(rng.integers(1, n + 1))] = oracle.CH_ABSENT
        if case % 3 == 0:
            modes[n] = oracle.CH_NEW
        new_mask = synth.onehot(synth.rect_label_map(rng, K - 1, H, W), K).astype(np.int32)
        lt, et = torch_mask_epilogue(torch.from_numpy(x), K, H, W, modes, torch.from_numpy(new_mask))
        lo, eo = oracle.mask_epilogue(x, K, (H, W), modes, new_mask)
        assert (np.abs(lo - lt.numpy()) <= synth.epilogue_logit_tolerance(x, K, H, W)).all(), f"case {case}"
        assert np.abs(eo - et.numpy()).max() <= 1e-5
    for case in range(6):
        n, T, h, w = int(rng.integers(1, 4)), int(rng.integers(1, 5)), int(rng.integers(1, 9)), int(rng.integers(1, 11))
        mk, mv, qk, qv = synth.memory_read_inputs(200 + case, n, T, h, w, float(rng.choice([0.3, 1.0, 2.0])))
        M, N = T * h * w, h * w
        mi = torch.transpose(torch.from_numpy(mk).view(n, synth.CK, M), 1, 2)
        p = torch.softmax(torch.bmm(mi, torch.from_numpy(qk).view(n, synth.CK, N)) / math.sqrt(synth.CK), dim=1)
        mem = torch.bmm(torch.from_numpy(mv).view(n, synth.CV, M), p).view(n, synth.CV, h, w)
        ref = torch.cat([mem, torch.from_numpy(qv)], dim=1).numpy()
        got, p_o = oracle.memory_read(mk, mv, qk, qv, want_p=True)
        assert np.abs(got - ref).max() <= 2e-5 and np.abs(p_o - p.numpy()).max() <= 2e-6, f"case {case}"
