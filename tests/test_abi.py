"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/rmnet_b200.h declares
(no compute calls without a GPU), argument validation returns error codes, host-side helpers agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
import rmnet_b200
from rmnet_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rmnet_b200.h")).read()
    return sorted(set(re.findall(r"RMNET_API[^;(]*?\b(rmnet_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 19
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"librmnet_b200.so does not export {n}"
    # and the ctypes prototypes cover exactly the header
    assert sorted(_lib.PROTOTYPES) == names


def test_abi_version_and_sizes():
    L = rmnet_b200.lib()
    assert L.rmnet_abi_version() == 1
    assert L.rmnet_reg_att_map_workspace_bytes(1, 11) == 16 * 2 * 12 * 8 * 4   # 16 copies of two accumulator sets (warped, direct)
    b1, b2 = L.rmnet_bank_bytes(3, 8128), L.rmnet_bank_bytes(3, 2 * 8128)
    assert b1 > 3 * 8128 * 640 * 4 and b2 > b1
    assert L.rmnet_bank_bytes(0, 64) == 0
    assert L.rmnet_memory_reader_workspace_bytes(3, 5, 30, 54) > L.rmnet_bank_bytes(3, 5 * 1620)
    # split-KV partials: at least one [n_obj,512,nq_pad] f32 block per split; grows with the bank capacity
    w1, w2 = L.rmnet_memory_read_workspace_bytes(3, 30, 54, 8128), L.rmnet_memory_read_workspace_bytes(3, 30, 54, 32448)
    assert w1 >= 3 * 512 * 1664 * 4 and w2 >= w1
    # ... plus the read kernel's work plan: a header per persistent CTA and the piece lists (16 bytes per piece)
    assert w1 >= 16 * 3 * 512 * 1664 * 4 + 256 * 8 + 256 * 16 * 16


def test_argument_validation_returns_error_codes_without_touching_the_gpu():
    L = rmnet_b200.lib()
    assert L.rmnet_reg_att_map_forward(None, 1, 11, 8, 8, 0.5, 10, 64, None, None, None, 0, None) == -1
    assert b"null" in L.rmnet_last_error()
    assert L.rmnet_warp_forward(None, None, 1, 1, 8, 8, 0, None, None, None) == -1
    assert L.rmnet_update_optical_flow(None, None, None, 4, 4, None, None) == -1
    assert L.rmnet_bank_reset(None, 0, 1, 64, None) == -1
    buf = ctypes.create_string_buffer(4096)
    # K must be >= 2 (channel 0 is the background), bad shapes are rejected before any launch
    addr = (ctypes.addressof(buf) + 15) // 16 * 16
    assert L.rmnet_reg_att_map_forward(addr, 1, 1, 8, 8, 0.5, 10, 64, addr, None, addr, 4096, None) == -1
    assert L.rmnet_reg_att_map_forward(addr, 1, 11, 8, 8, 0.5, 10, 64, addr, None, addr, 8, None) == -3  # workspace too small
    # the region workspace is 16 copies of two accumulator sets and must be 16-byte aligned
    assert L.rmnet_reg_att_map_forward(addr, 1, 2, 8, 8, 0.5, 10, 64, addr, None, addr + 4, 4096, None) == -1
    assert b"aligned" in L.rmnet_last_error()
    # plan read-back: null pointers / bad sizes are rejected before any CUDA call
    assert L.rmnet_memory_read_plan_host(None, 3, 30, 54, None, None, None, 16, None, None) == -1
    n_ctas = ctypes.c_int(0)
    assert L.rmnet_memory_read_plan_host(addr, 0, 30, 54, addr, addr, addr, 16, ctypes.byref(n_ctas), None) == -1
    assert L.rmnet_memory_read_plan_host(addr, 65, 30, 54, addr, addr, addr, 16, ctypes.byref(n_ctas), None) == -1   # > 64 objects: no tcgen05 plan


def test_wrappers_reject_cpu_tensors_like_the_reference_check_input():
    import torch
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):   # reg_att_map_generator_cuda.cpp:14,29
        ops.reg_att_map_forward(torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        rmnet_b200.MemoryReader()(torch.zeros(1, 128, 1, 2, 2), torch.zeros(1, 512, 1, 2, 2),
                                  torch.zeros(1, 128, 2, 2), torch.zeros(1, 512, 2, 2))
    with pytest.raises(ValueError):
        rmnet_b200.update_optical_flow(np.zeros((4, 4), np.float32), np.eye(2, 3), np.eye(2, 3))


def test_pad_amounts_match_oracle():
    for h, w in [(480, 854), (240, 432), (720, 1280), (33, 47), (481, 865), (16, 16)]:
        assert ops.pad_amounts(h, w) == oracle.pad_amounts(h, w)


def test_cell_rect_closed_form_equals_pad_then_downsample():
    """The closed form the CUDA path uses (rmnet_cell_rects_from_bboxes) == the reference's pad + interpolate(1/16)
    on a rectangular att map (models/rmnet.py:245, :307, :356), checked on the host restatement of the formula."""
    rng = np.random.default_rng(3)
    for _ in range(300):
        H, W = int(rng.integers(17, 200)), int(rng.integers(17, 200))
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        x1, y1 = int(rng.integers(x0, W)), int(rng.integers(y0, H))
        att = np.zeros((1, H, W), np.float32)
        att[0, y0:y1 + 1, x0:x1 + 1] = 1
        attp, (lw, uw, lh, uh) = oracle.pad_divide_by(att)
        a16 = oracle.downsample16(attp)[0]
        h, w = a16.shape
        cx0, cx1 = max(0, (x0 + lw + 15) >> 4), min(w - 1, (x1 + lw) >> 4)
        cy0, cy1 = max(0, (y0 + lh + 15) >> 4), min(h - 1, (y1 + lh) >> 4)
        ref = np.zeros_like(a16)
        if cx0 <= cx1 and cy0 <= cy1:
            ref[cy0:cy1 + 1, cx0:cx1 + 1] = 1
        np.testing.assert_array_equal(a16, ref)


def test_frame_loop_host_logic_matches_the_reference_rules():
    """rmnet_b200.frame_loop: the memorise schedule (models/rmnet.py:405-408, :424) and the per-frame channel overrides
    (:436-448), against a literal restatement of the reference's loops."""
    from rmnet_b200.frame_loop import CH_ABSENT, CH_KEEP, CH_NEW, channel_modes, memorize_schedule
    for n_frames, every, nobj in ((16, 5, [1] * 16), (12, 3, [2] * 4 + [3] * 5 + [4] * 3), (7, 1, [1, 1, 2, 2, 2, 3, 3]), (3, 10, [2, 2, 2])):
        to_mem, new_at, commit, n_commits = memorize_schedule(n_frames, every, nobj)
        ref_to_mem = [j for j in range(0, n_frames, every)]
        ref_new = [j for j in range(1, n_frames) if nobj[j] != nobj[j - 1]]
        assert sorted(to_mem) == ref_to_mem and sorted(new_at) == ref_new
        T = 0   # the reference's bank length: `keys` grows by one frame whenever t-1 is committed (:424-426)
        for t in range(1, n_frames):
            ref_commit = (t - 1 in ref_to_mem) or (t - 1 in ref_new)
            assert commit[t] == ref_commit
            T += int(ref_commit)
        assert T == n_commits
    # with memorize_every = 5, T(t) = 1 + #{j in {0,5,10,...} : j < t-1} ... the bank reaches 20 committed frames at t = 96
    _, _, commit, _ = memorize_schedule(100, 5, [1] * 100)
    assert sum(commit[t] for t in range(1, 97)) == 20
    K, n_max = 11, 3
    existing = [0, 1]
    assert channel_modes(K, n_max, existing, None) == [CH_KEEP, CH_KEEP, CH_ABSENT, CH_ABSENT] + [CH_KEEP] * 7
    m = channel_modes(K, n_max, existing, [0, 1, 3])          # object 3 is first annotated in this frame
    assert m == [CH_KEEP, CH_KEEP, CH_ABSENT, CH_NEW] + [CH_KEEP] * 7 and existing == [0, 1, 3]
    assert channel_modes(K, n_max, existing, [0, 1, 3]) == [CH_KEEP, CH_KEEP, CH_ABSENT, CH_KEEP] + [CH_KEEP] * 7


# ---------------------------------------------------------------------------------------------------------------
# update_optical_flow without a GPU: the library's plain-C entry point (the one a forked DataLoader worker gets)
# ---------------------------------------------------------------------------------------------------------------
def _flow_case(seed, H, W, half):
    import synth
    rng = np.random.default_rng(seed)
    of = np.ascontiguousarray(np.moveaxis(synth.flow_field(rng, H, W, 4.0, half_pixel=half), 0, -1))
    m1, m2 = synth.affine_pair(rng)
    return of, m1, m2


def test_update_optical_flow_cpu_entry_point_bit_exact_vs_golden_oracle_and_reference_extension(golden_dir):
    """rmnet_update_optical_flow_cpu (flow_affine_transformation.cpp:63-83 in plain C, -ffp-contract=off) against the
    golden vectors the unmodified reference extension produced, the C oracle, and the reference extension live."""
    import glob
    import sys
    import synth
    g = np.load(os.path.join(golden_dir, "flow_affine.npz"))
    for i in range(int(g["n_cases"])):
        seed, H, W = (int(g[f"c{i}_{k}"]) for k in ("seed", "H", "W"))
        rng = np.random.default_rng(seed)
        of = np.ascontiguousarray(np.moveaxis(synth.flow_field(rng, H, W, float(g[f"c{i}_sigma"])), 0, -1))
        m1, m2 = synth.affine_pair(rng)
        np.testing.assert_array_equal(ops.update_optical_flow(of, m1, m2, device="cpu"), g[f"c{i}_out"])
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "flow_affine_transformation*.so"))
    ref = None
    if so:
        sys.path.insert(0, os.path.dirname(so[0]))
        sys.modules.pop("flow_affine_transformation", None)
        import flow_affine_transformation as ref
    for s in range(8):
        of, m1, m2 = _flow_case(300 + s, 37 + 31 * s, 50 + 17 * s, s % 2 == 0)
        if s == 0:   # exact .5 ties for round-half-away
            m1, m2 = np.array([[1, 0, 0.5], [0, 1, -0.5]], np.float32), np.array([[1, 0, 0], [0, 1, 0]], np.float32)
        out = ops.update_optical_flow(of, m1, m2, device="cpu")
        np.testing.assert_array_equal(out, oracle.update_optical_flow(of, m1, m2))
        if ref is not None:
            np.testing.assert_array_equal(out, ref.update_optical_flow(of, m1, m2))
    # the reference's float64-zeros hazard (utils/data_loaders.py:54-55): converted, not reinterpreted
    z = ops.update_optical_flow(np.zeros((8, 9, 2)), np.eye(2, 3), np.eye(2, 3), device="cpu")
    assert z.dtype == np.float32 and not z.any()


class _FlowDataset:
    """A dataset whose __getitem__ calls the drop-in exactly where the reference does (RandomAffine.__call__,
    utils/data_transforms.py:293-302): inside a DataLoader worker process."""

    def __len__(self):
        return 4

    def __getitem__(self, i):
        import flow_affine_transformation            # the drop-in module of that name (rmnet_b200/dropin on sys.path)
        of, m1, m2 = _flow_case(900 + i, 64, 80, False)
        return flow_affine_transformation.update_optical_flow(of, m1, m2)


def test_update_optical_flow_dropin_runs_inside_forked_dataloader_workers():
    import sys
    import torch
    dropin = os.path.join(ROOT, "rmnet_b200", "dropin")
    sys.path.insert(0, dropin)
    sys.modules.pop("flow_affine_transformation", None)
    try:
        loader = torch.utils.data.DataLoader(_FlowDataset(), batch_size=1, num_workers=2, multiprocessing_context="fork")
        outs = [b[0].numpy() for b in loader]
    finally:
        sys.path.remove(dropin)
        sys.modules.pop("flow_affine_transformation", None)
    for i, out in enumerate(outs):
        of, m1, m2 = _flow_case(900 + i, 64, 80, False)
        np.testing.assert_array_equal(out, oracle.update_optical_flow(of, m1, m2))
