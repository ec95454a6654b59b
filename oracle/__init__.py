"""oracle -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's hot path).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``rmnet_b200``) never
does: it fails loudly when its CUDA library is missing instead of falling back to this.

Parity pin status (the reference's own tests hold NO golden vector for this path --
``extensions/flow_affine_transformation/test.py:14-22`` only prints):
  * ``update_optical_flow``  -- pinned against the UNMODIFIED reference extension compiled
    into ``oracle/_ref`` (bit-exact, tests/test_oracle.py) and against golden vectors it produced.
  * ``reg_att_map``          -- pinned on the GPU box against the UNMODIFIED reference CUDA
    extension compiled into ``oracle/_ref`` (bit-exact, tests/test_gpu_parity.py).
  * ``warp`` / ``memory_read`` / ``downsample16`` / ``pad`` -- pinned against golden vectors
    produced by importing the reference's Python (models/rmnet.py) in the build container
    (tests/golden/make_golden.py, committed next to the vectors).

  * ``mask_epilogue``        -- pinned against golden vectors produced by the reference's own
    ``RMNet.soft_aggregation`` + the torch ops of models/rmnet.py:368-370, :436-450.

Files: ``rmnet_oracle.c`` (plain C: integer / byte-exact parts), ``memory_read.py`` (numpy:
the BLAS-backed floating-point reader), ``mask_epilogue.py`` (numpy: the post-decoder tail).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile oracle/rmnet_oracle.c (gcc) into oracle/_build/liboracle.so."""
    src = os.path.join(_HERE, "rmnet_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "_build/liboracle.so"])
    return _LIB_PATH


def build_ref(reference_root="/root/reference"):
    """Compile the UNMODIFIED reference extensions into oracle/_ref (only where the reference
    tree exists, i.e. the build container).  Returns the list of built files."""
    if not os.path.isdir(reference_root):
        return []
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REF=" + reference_root])
    d = os.path.join(_HERE, "_ref")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".so"))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_reg_att_map.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_float, ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_void_p]
        L.oracle_pad_amounts.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p]
        L.oracle_pad2d.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, _f32p]
        L.oracle_downsample16_nearest.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        L.oracle_warp.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  _f32p, _f32p]
        L.oracle_update_optical_flow.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int, _f32p]
        L.oracle_memory_read_f64.argtypes = [_f32p, _f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_void_p]
        _lib = L
    return _lib


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def reg_att_map(mask, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64, want_att=True):
    """reg_att_map_generator.cu:31-92.  mask [B,K,H,W] f32 -> (att_map [B,K,H,W] f32 | None, bboxes [B,K,4] i32)."""
    mask = _c(mask)
    B, K, H, W = mask.shape
    bboxes = np.zeros((B, K, 4), np.int32)
    att = np.zeros((B, K, H, W), np.float32) if want_att else None
    lib().oracle_reg_att_map(mask, B, K, H, W, prob_threshold, n_pts_threshold, n_bbox_loose_pixels, bboxes,
                             att.ctypes.data if want_att else None)
    return att, bboxes


def pad_amounts(h, w, d=16):
    """utils/helpers.py:105-119 -> (lw, uw, lh, uh)."""
    pad = np.zeros(4, np.int32)
    lib().oracle_pad_amounts(h, w, d, pad)
    return tuple(int(v) for v in pad)


def pad_divide_by(x, d=16):
    """utils/helpers.py:105-124 for one [..., H, W] array (leading dims flattened)."""
    x = _c(x)
    H, W = x.shape[-2:]
    pad = np.array(pad_amounts(H, W, d), np.int32)
    C = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    out = np.empty(x.shape[:-2] + (H + pad[2] + pad[3], W + pad[0] + pad[1]), np.float32)
    lib().oracle_pad2d(x.reshape(C, H, W), C, H, W, pad, out.reshape(C, out.shape[-2], out.shape[-1]))
    return out, tuple(int(v) for v in pad)


def downsample16(att):
    """F.interpolate(att, scale_factor=1/16) nearest, models/rmnet.py:245,356.  [..., Hp, Wp] -> [..., Hp//16, Wp//16]."""
    att = _c(att)
    Hp, Wp = att.shape[-2:]
    C = int(np.prod(att.shape[:-2])) if att.ndim > 2 else 1
    out = np.empty(att.shape[:-2] + (Hp // 16, Wp // 16), np.float32)
    lib().oracle_downsample16_nearest(att.reshape(C, Hp, Wp), C, Hp, Wp, out.reshape(C, Hp // 16, Wp // 16))
    return out


def warp(img0, flow, arith="cuda"):
    """RMNet.warp, models/rmnet.py:252-278.  img0 [B,C,H,W], flow [B,2,H,W] -> (img1 [B,C,H,W], mask [B,C,H,W]).
    ``arith`` selects which torch backend's floating-point evaluation order is mirrored:
      'cpu'         true division, ATen tap order      (what the reference computes on a CPU-only box)
      'cuda_native' reciprocal multiply, ATen tap order (CUDA with torch.backends.cudnn.enabled = False)
      'cuda' / 'cuda_cudnn'  reciprocal multiply, cuDNN tap order (the reference's default GPU path)"""
    img0, flow = _c(img0), _c(flow)
    B, C, H, W = img0.shape
    img1 = np.empty_like(img0)
    valid = np.empty((B, 1, H, W), np.float32)
    a, order = {"cpu": (0, 0), "cuda_native": (1, 0), "cuda": (1, 1), "cuda_cudnn": (1, 1)}[arith]
    for b in range(B):
        lib().oracle_warp(img0[b], flow[b], C, H, W, a, order, img1[b], valid[b, 0])
    return img1, np.broadcast_to(valid, img0.shape).copy()


def get_att_map(prev_mask, flow=None, arith="cuda", **kw):
    """RMNet.get_att_map, models/rmnet.py:280-287."""
    expt = prev_mask if flow is None else warp(prev_mask, flow, arith)[0]
    return reg_att_map(expt, **kw)


def update_optical_flow(of, m1, m2):
    """flow_affine_transformation.cpp:63-83.  of [H,W,2] f32, m1/m2 [2,3] f32 -> [H,W,2] f32."""
    of, m1, m2 = _c(of), _c(m1), _c(m2)
    H, W = of.shape[:2]
    out = np.empty_like(of)
    lib().oracle_update_optical_flow(of, m1.reshape(-1), m2.reshape(-1), H, W, out)
    return out


def memory_read_f64(m_key, m_val, q_key, q_val, want_p=False):
    """models/rmnet.py:147-165 with scalar loops + double accumulation (small cases only)."""
    m_key, m_val, q_key, q_val = _c(m_key), _c(m_val), _c(q_key), _c(q_val)
    n, Ck, T, h, w = m_key.shape
    Cv = m_val.shape[1]
    M, N = T * h * w, h * w
    out = np.empty((n, 2 * Cv, h, w), np.float32)
    p = np.empty((n, M, N), np.float32) if want_p else None
    for o in range(n):
        lib().oracle_memory_read_f64(m_key[o].reshape(Ck, M), m_val[o].reshape(Cv, M), q_key[o].reshape(Ck, N),
                                     q_val[o].reshape(Cv, N), Ck, Cv, M, N, out[o].reshape(2 * Cv, N),
                                     p[o].ctypes.data if want_p else None)
    return out, p


from .memory_read import (memory_read, regional_mask_memory, regional_mask_query,  # noqa: E402,F401
                          regional_memory_read)
from .mask_epilogue import CH_ABSENT, CH_KEEP, CH_NEW, mask_epilogue  # noqa: E402,F401
