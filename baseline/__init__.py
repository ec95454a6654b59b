"""baseline -- TEST / BENCH INFRASTRUCTURE ONLY: runs the UNMODIFIED reference (hzxie/RMNet) next to rmnet_b200.

`baseline/_ref/` is a verbatim, git-ignored copy of the reference tree (made by `populate()` from /root/reference in
the build container; it is NOT gpurun-ignored, so it travels to the GPU box where /root/reference does not exist).
Nothing of it is ever committed, and the product package (`rmnet_b200/`) never imports this module.

What lives here (SURVEY 8c's harness-side shims -- never edits to the reference):
  * `import_reference()`   sys.path + the shims the reference needs to import offline: an `easydict` stand-in
                           (config.py:9), `torchvision.models.resnet50(pretrained=True)` -> random init (no network,
                           models/rmnet.py:57,86), the compiled reference extensions from `oracle/_ref`
                           (extensions/reg_att_map_generator/__init__.py:11 imports `reg_att_map_generator`).
  * `CpuGenerator`         CPU stand-in for the CUDA-only generator (reg_att_map_generator_cuda.cpp:14-19 refuses CPU
                           tensors) through the C oracle -- only for running the reference on the host cores.
  * `build_nets()`         TinyFlowNet + RMNet with default torch init under a fixed seed (never
                           `init_weights`), optionally "conditioned" (key convs rescaled so that the scores are O(10)).
  * `synthetic_clip()`     SURVEY 8d's seeded clip: N(0,1) frames, drifting rectangles, one-hot int32 masks.
  * `test_cfg()`           the few cfg.TEST keys utils/helpers.py:44-62 reads.
"""
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_SRC = "/root/reference"
K_TEST = 11   # N_MAX_OBJECTS + 1 mask channels at test time (config.py:137, utils/data_loaders.py:214-219)


def populate(force=False):
    """Copy /root/reference -> baseline/_ref (git-ignored) when the source tree exists (build container only)."""
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(REF_COPY)
    if force or not os.path.isfile(os.path.join(REF_COPY, "models", "rmnet.py")):
        if os.path.isdir(REF_COPY):
            shutil.rmtree(REF_COPY)
        shutil.copytree(REF_SRC, REF_COPY, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    return True


def ref_root():
    if os.path.isfile(os.path.join(REF_COPY, "models", "rmnet.py")):
        return REF_COPY
    if os.path.isfile(os.path.join(REF_SRC, "models", "rmnet.py")):
        return REF_SRC
    return None


def available():
    return ref_root() is not None


_imported = None


def import_reference(need_cuda_extension=False):
    """-> the reference's `models.rmnet` module (imported once, with the offline shims).  Raises RuntimeError when the
    reference tree is absent."""
    global _imported
    if _imported is not None:
        return _imported
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (baseline/_ref is populated by __graft_entry__.build() in the build container)")
    # (i) easydict stand-in (config.py:9)
    if "easydict" not in sys.modules:
        try:
            import easydict  # noqa: F401
        except ImportError:
            mod = types.ModuleType("easydict")

            class EasyDict(dict):
                def __getattr__(self, k):
                    try:
                        return self[k]
                    except KeyError:
                        raise AttributeError(k)

                def __setattr__(self, k, v):
                    self[k] = v

            mod.EasyDict = EasyDict
            sys.modules["easydict"] = mod
    # (ii) resnet50(pretrained=True) -> random init: there is no network for the ImageNet weights
    import torchvision.models
    if not getattr(torchvision.models.resnet50, "_rmnet_offline", False):
        orig = torchvision.models.resnet50

        def resnet50(pretrained=False, **kw):
            return orig(weights=None, **kw)

        resnet50._rmnet_offline = True
        torchvision.models.resnet50 = resnet50
    # (iii) the compiled reference extensions (oracle/_ref), else a stub module so that the import succeeds on a box
    #       without them (the CUDA generator is then unusable; CpuGenerator / rmnet_b200 replace it)
    ext = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(ext) and ext not in sys.path:
        sys.path.insert(0, ext)
    try:
        import reg_att_map_generator  # noqa: F401
    except Exception:
        if need_cuda_extension:
            raise RuntimeError("the reference CUDA extension (oracle/_ref/reg_att_map_generator*.so) is not built / loadable")
        stub = types.ModuleType("reg_att_map_generator")

        def forward(*a, **k):
            raise RuntimeError("reference extension reg_att_map_generator not built")

        stub.forward = forward
        sys.modules["reg_att_map_generator"] = stub
    if root not in sys.path:
        sys.path.insert(0, root)
    import models.rmnet as ref_rmnet
    _imported = ref_rmnet
    return ref_rmnet


def cpu_generator_class():
    """torch.nn.Module with RegionalAttentionMapGenerator's signature running the C oracle (CPU tensors only)."""
    import numpy as np
    import torch

    import oracle

    class CpuGenerator(torch.nn.Module):
        def forward(self, mask, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64):
            att, bb = oracle.reg_att_map(np.ascontiguousarray(mask.detach().cpu().numpy(), dtype=np.float32), prob_threshold,
                                         n_pts_threshold, n_bbox_loose_pixels)
            return torch.from_numpy(att), torch.from_numpy(bb)

    return CpuGenerator


def test_cfg(memorize_every=5):
    """The cfg keys utils/helpers.py:44-62 (multi_scale_inference) and the model constructors read."""
    import_reference()
    from easydict import EasyDict
    cfg = EasyDict()
    cfg.TEST = EasyDict(FRAME_SCALES=[1.0], FLIP_LR=False, MEMORIZE_EVERY=memorize_every, N_MAX_OBJECTS=K_TEST - 1)
    cfg.CONST = EasyDict(DATASET_MEAN=[0.485, 0.456, 0.406], DATASET_STD=[0.229, 0.224, 0.225])
    return cfg


def build_nets(seed=0, device="cpu", conditioned=False, cpu_generator=None, with_flownet=True):
    """(tflownet, rmnet) in eval mode, default torch init under torch.manual_seed(seed).  conditioned=True
    rescales the key convs of kv_memory / kv_query so that the scaled scores are O(10) instead of O(100s)."""
    import torch
    ref = import_reference()
    from models.tiny_flownet import TinyFlowNet
    cfg = test_cfg()
    torch.manual_seed(seed)
    tfn = TinyFlowNet(cfg).eval() if with_flownet else None
    net = ref.RMNet(cfg).eval()
    if conditioned:
        with torch.no_grad():
            for kv in (net.kv_memory, net.kv_query):
                kv.key_conv.weight.mul_(0.15)
                kv.key_conv.bias.mul_(0.15)
    if cpu_generator is None:
        cpu_generator = str(device) == "cpu"
    if cpu_generator:
        net.att_map_generator = cpu_generator_class()()
    if tfn is not None:
        tfn = tfn.to(device)
    return tfn, net.to(device)


def synthetic_clip(seed, n_objects, n_frames, H, W, K=K_TEST, new_object_at=None):
    """frames [1,F,3,H,W] ~ N(0,1) f32; n rectangles of size U(0.15,0.45)*(H,W) drifting by an integer velocity
    in [-3,3] px/frame, painted in order 1..n; masks [1,F,K,H,W] int32 one-hot; n_objects [1,F] int64.
    new_object_at = frame index from which object n appears (n_objects = n-1 before it)."""
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    frames = torch.randn((1, n_frames, 3, H, W), generator=g, dtype=torch.float32)
    boxes = []
    for _ in range(n_objects):
        bh, bw = int(rng.uniform(0.15, 0.45) * H), int(rng.uniform(0.15, 0.45) * W)
        boxes.append((int(rng.integers(0, H - bh)), int(rng.integers(0, W - bw)), bh, bw, int(rng.integers(-3, 4)), int(rng.integers(-3, 4))))
    masks = np.zeros((1, n_frames, K, H, W), np.int32)
    n_obj = np.zeros((1, n_frames), np.int64)
    for t in range(n_frames):
        lab = np.zeros((H, W), np.int64)
        n_t = n_objects if (new_object_at is None or t >= new_object_at) else n_objects - 1
        for o, (y0, x0, bh, bw, vy, vx) in enumerate(boxes[:n_t], 1):
            y = int(np.clip(y0 + vy * t, 0, H - bh))
            x = int(np.clip(x0 + vx * t, 0, W - bw))
            lab[y:y + bh, x:x + bw] = o
        for k in range(K):
            masks[0, t, k] = lab == k
        n_obj[0, t] = n_t
    return frames, torch.from_numpy(masks), torch.from_numpy(n_obj)
