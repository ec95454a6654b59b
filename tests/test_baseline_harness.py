"""CPU test of the harness that runs the UNMODIFIED reference (baseline/): BASELINE configs[0] -- a synthetic 240x432 clip,
1 object, through the reference's literal `core.inference.inference_net(cfg)` on the CPU (easydict stand-in,
random-init ResNet-50s, synthetic DAVIS tree, the CUDA-only generator replaced by the C oracle).  No product code runs here;
the test pins the harness the GPU tests and the bench legs rely on, and the reference's own output contract (est_masks of
RMNet.forward: a softmax over the K channels, frame 0 = the given masks; one overlay PNG per frame)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

import baseline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_reference():
    if not baseline.available():
        pytest.skip("reference tree not available (baseline/_ref is populated by __graft_entry__.build())")
    return baseline.import_reference()


def test_reference_forward_runs_on_the_cpu_through_the_harness():
    _need_reference()
    H, W, n, F_ = 240, 432, 1, 4
    tfn, net = baseline.build_nets(0, "cpu")
    frames, masks, n_objects = baseline.synthetic_clip(1, n, F_, H, W)
    import utils.helpers as ref_helpers
    with torch.no_grad():
        flows, probs = ref_helpers.multi_scale_inference(baseline.test_cfg(5), tfn, net, frames, masks, n_objects)
    assert tuple(probs.shape) == (1, F_, baseline.K_TEST, H, W) and tuple(flows.shape) == (1, F_, 2, H, W)
    assert torch.equal(probs[0, 0], masks[0, 0].float())                      # models/rmnet.py:396
    assert float((probs[0, 1:].sum(1) - 1).abs().max()) <= 1e-5             # :450 softmax over the channels
    assert float(probs[0, 1:, n + 1:].max()) <= 1e-6                         # channels of absent objects: -16.1181 logits (:448)


def test_literal_inference_net_on_the_cpu(tmp_path):
    ref = _need_reference()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_inference_net import _make_davis_tree
    try:
        import flow_affine_transformation  # noqa: F401  (utils/data_transforms.py:18)
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, "rmnet_b200", "dropin"))
    from PIL import Image
    import core.inference as ref_inference
    from config import __C as ref_cfg
    H, W, n, F_ = 240, 432, 1, 3
    data = str(tmp_path / "davis")
    os.makedirs(data)
    cfg = copy.deepcopy(ref_cfg)
    cfg.DATASETS.DAVIS.INDEXING_FILE_PATH = _make_davis_tree(data, "clipA", F_, H, W, n, seed=3)
    cfg.DATASETS.DAVIS.IMG_FILE_PATH = os.path.join(data, "JPEGImages", "480p", "%s", "%05d.jpg")
    cfg.DATASETS.DAVIS.ANNOTATION_FILE_PATH = os.path.join(data, "Annotations", "480p", "%s", "%05d.png")
    cfg.DATASETS.DAVIS.OPTICAL_FLOW_FILE_PATH = os.path.join(data, "OpticalFlows", "480p", "%s", "%05d.flo")
    cfg.DATASET.TEST_DATASET, cfg.CONST.N_WORKERS, cfg.DIR.OUTPUT_DIR, cfg.CONST.EXP_NAME = "DAVIS", 0, str(tmp_path / "out"), "cpu"
    tfn, net = baseline.build_nets(0, "cpu")
    ckpt = str(tmp_path / "ckpt.pth")
    torch.save({"tflownet": tfn.state_dict(), "rmnet": net.state_dict()}, ckpt)
    cfg.CONST.WEIGHTS = ckpt
    cuda_was, gen_was = torch.cuda.is_available, ref.RegionalAttentionMapGenerator
    torch.cuda.is_available = lambda: False                                   # core/inference.py:35: stay on the CPU on any box
    ref.RegionalAttentionMapGenerator = baseline.cpu_generator_class()        # the reference's generator refuses CPU tensors
    try:
        ref_inference.inference_net(cfg)
    finally:
        torch.cuda.is_available, ref.RegionalAttentionMapGenerator = cuda_was, gen_was
    d = os.path.join(cfg.DIR.OUTPUT_DIR, "benchmark", "cpu", "DAVIS", "clipA")
    assert sorted(os.listdir(d)) == ["%05d.png" % i for i in range(F_)]
    assert np.array(Image.open(os.path.join(d, "00001.png"))).shape == (H, W, 3)
