"""GPU test (pytest -m gpu): the reference's LITERAL driver -- `core.inference.inference_net(cfg)` (core/inference.py:21-71:
its DataLoader, its DataParallel wrapping, its checkpoint loading, utils.helpers.multi_scale_inference, its PNG writer),
imported unmodified from baseline/_ref -- on a synthetic DAVIS-layout dataset on disk (SURVEY 8c shim iii), once as is and
once after `rmnet_b200.install(models.rmnet)`.  Nothing of the reference is edited; the second run must execute this
library's kernels and write the same segmentation overlays."""
import copy
import json
import os
import sys

import numpy as np
import pytest
import torch

import baseline
import rmnet_b200

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_davis_tree(root, name, n_frames, H, W, n_objects, seed):
    from PIL import Image
    rng = np.random.default_rng(seed)
    img_dir, ann_dir = os.path.join(root, "JPEGImages", "480p", name), os.path.join(root, "Annotations", "480p", name)
    os.makedirs(img_dir), os.makedirs(ann_dir)
    base = rng.integers(0, 255, (H // 8, W // 8, 3), dtype=np.uint8)
    for i in range(n_frames):
        frame = np.kron(np.roll(base, 2 * i, axis=1), np.ones((8, 8, 1), np.uint8))            # a drifting blocky texture
        Image.fromarray(frame).save(os.path.join(img_dir, "%05d.jpg" % i), quality=95)
    lab = np.zeros((H, W), np.uint8)
    for o in range(1, n_objects + 1):
        y0, x0 = 20 + 50 * o, 30 + 90 * o
        lab[y0:y0 + 70, x0:x0 + 110] = o
    pal = Image.fromarray(lab, mode="P")
    pal.putpalette([0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0] + [0] * (256 * 3 - 12))
    pal.save(os.path.join(ann_dir, "00000.png"))                                                # only frame 0 is annotated
    index = os.path.join(root, "DAVIS.json")
    json.dump({"test": [{"name": name, "n_frames": n_frames}], "val": [], "train": []}, open(index, "w"))
    return index


def test_literal_inference_net_runs_the_installed_path_and_writes_the_same_overlays(tmp_path):
    if not baseline.available():
        pytest.skip("baseline/_ref not populated")
    try:
        ref = baseline.import_reference(need_cuda_extension=True)
    except RuntimeError as e:
        pytest.skip(str(e))
    try:
        import flow_affine_transformation  # noqa: F401  (utils/data_transforms.py:18 imports it: the reference's from oracle/_ref ...)
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, "rmnet_b200", "dropin"))                          # ... else the drop-in of that name
    from PIL import Image
    import core.inference as ref_inference
    from config import __C as ref_cfg
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False           # runner.py:73-74

    H, W, n, F_ = 240, 432, 2, 7
    data = str(tmp_path / "davis")
    os.makedirs(data)
    index = _make_davis_tree(data, "clipA", F_, H, W, n, seed=3)
    cfg = copy.deepcopy(ref_cfg)
    cfg.DATASETS.DAVIS.INDEXING_FILE_PATH = index
    cfg.DATASETS.DAVIS.IMG_FILE_PATH = os.path.join(data, "JPEGImages", "480p", "%s", "%05d.jpg")
    cfg.DATASETS.DAVIS.ANNOTATION_FILE_PATH = os.path.join(data, "Annotations", "480p", "%s", "%05d.png")
    cfg.DATASETS.DAVIS.OPTICAL_FLOW_FILE_PATH = os.path.join(data, "OpticalFlows", "480p", "%s", "%05d.flo")
    cfg.DATASET.TEST_DATASET = "DAVIS"
    cfg.CONST.N_WORKERS = 0
    cfg.DIR.OUTPUT_DIR = str(tmp_path / "out")
    cfg.TEST.MEMORIZE_EVERY = 3
    # synthetic checkpoint {'tflownet', 'rmnet'} with the DataParallel key prefix inference_net expects on a CUDA box (:35-43)
    tfn, net = baseline.build_nets(0, "cpu", conditioned=True, cpu_generator=False)
    ckpt = str(tmp_path / "ckpt.pth")
    torch.save({"tflownet": torch.nn.DataParallel(tfn).state_dict(), "rmnet": torch.nn.DataParallel(net).state_dict()}, ckpt)
    cfg.CONST.WEIGHTS = ckpt

    def run(exp):
        cfg.CONST.EXP_NAME = exp
        ref_inference.inference_net(cfg)
        d = os.path.join(cfg.DIR.OUTPUT_DIR, "benchmark", exp, "DAVIS", "clipA")
        return [np.array(Image.open(os.path.join(d, "%05d.png" % i))) for i in range(F_)]

    rmnet_b200.uninstall(ref)
    want = run("reference")
    from rmnet_b200 import frame_loop
    L = rmnet_b200.lib()
    rmnet_b200.install(ref)
    try:
        L.rmnet_launch_count_reset()
        g0 = frame_loop.graph_launch_total[0]
        got = run("installed")
        launches = int(L.rmnet_launch_count()) + frame_loop.graph_launch_total[0] - g0
    finally:
        rmnet_b200.uninstall(ref)
    assert launches >= 5 * (F_ - 1), f"the installed forward launched only {launches} kernels of this library"
    same = [float((a == b).all(axis=-1).mean()) for a, b in zip(got, want)]
    print(f"\nliteral inference_net: {launches} kernels of this library; identical overlay pixels per frame {same}")
    assert got[0].shape == (H, W, 3) and min(same) >= 0.999
