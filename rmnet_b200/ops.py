"""Torch-tensor-facing wrappers around the C ABI: device memory, streams and shape checks only.

Every function takes CUDA float32 tensors, allocates outputs with torch's caching allocator, passes raw
pointers + the current stream to librmnet_b200.so, and never synchronises.  Error behaviour mirrors the
reference's CHECK_INPUT (reg_att_map_generator_cuda.cpp:14-19): non-CUDA / non-contiguous inputs raise RuntimeError.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import (ELEM_BF16, ELEM_FP16, RMNET_IMPL_AUTO, RMNET_IMPL_SIMT, RMNET_IMPL_UMMA, RMNET_PREC_SINGLE,
                   RMNET_PREC_SPLIT3, check, lib)

CK, CV = 128, 512
_ws_cache = {}


def umma_available():
    return bool(lib().rmnet_has_umma())


def _require(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}")


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _check_ws(rc, what, ws):
    """check() for the launches that use a self-cleaning (zero-identity) workspace: a failed call may have left it dirty,
    so it is re-zeroed before the error is raised."""
    if rc != 0:
        try:
            ws.zero_()
        except Exception:
            pass
    check(rc, what)


def _zero_ws(dev, nbytes):
    """Self-cleaning generator workspace, one per (device, stream): zero-filled once."""
    key = (dev.index, _stream(dev))
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def reg_att_map_forward(mask, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64, want_att=True):
    """reg_att_map_generator.forward (reg_att_map_generator_cuda.cpp:26-38) -> [att_map, bboxes]."""
    _require(mask, "mask")
    B, K, H, W = mask.shape
    dev = mask.device
    with torch.cuda.device(dev):
        bboxes = torch.empty((B, K, 4), dtype=torch.int32, device=dev)
        att = torch.empty((B, K, H, W), dtype=torch.float32, device=dev) if want_att else None
        nws = lib().rmnet_reg_att_map_workspace_bytes(B, K)
        ws = _zero_ws(dev, nws)
        _check_ws(lib().rmnet_reg_att_map_forward(mask.data_ptr(), B, K, H, W, float(prob_threshold), int(n_pts_threshold),
                                              int(n_bbox_loose_pixels), bboxes.data_ptr(),
                                              att.data_ptr() if want_att else None, ws.data_ptr(), ws.numel(),
                                              _stream(dev)), "reg_att_map_forward", ws)
    return [att, bboxes]


def default_sampler():
    """Which bilinear sampler the reference's F.grid_sample(bilinear, zeros, align_corners=True) resolves to on CUDA:
    cuDNN's spatial-transformer sampler when cuDNN is enabled (torch's default), ATen's own kernel otherwise."""
    return _lib.SAMPLER_CUDNN if torch.backends.cudnn.enabled else _lib.SAMPLER_ATEN


def warp(img0, flow, want_mask=True, sampler=None):
    """RMNet.warp (models/rmnet.py:252-278) -> (img1, mask)."""
    _require(img0, "img0")
    _require(flow, "flow")
    B, C, H, W = img0.shape
    if tuple(flow.shape) != (B, 2, H, W):
        raise RuntimeError(f"flow must be [{B},2,{H},{W}]")
    dev = img0.device
    with torch.cuda.device(dev):
        img1 = torch.empty_like(img0)
        valid = torch.empty_like(img0) if want_mask else None
        check(lib().rmnet_warp_forward(img0.data_ptr(), flow.data_ptr(), B, C, H, W,
                                       default_sampler() if sampler is None else sampler, img1.data_ptr(),
                                       valid.data_ptr() if want_mask else None, _stream(dev)), "warp_forward")
    return img1, valid


def warp_att_map_forward(prev_mask, flow, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64,
                         want_att=True, sampler=None):
    """RMNet.get_att_map(prev_mask, flow) (models/rmnet.py:280-287), warp fused into the bbox scan."""
    _require(prev_mask, "prev_mask")
    _require(flow, "flow")
    B, K, H, W = prev_mask.shape
    if tuple(flow.shape) != (B, 2, H, W):
        raise RuntimeError(f"flow must be [{B},2,{H},{W}]")
    dev = prev_mask.device
    with torch.cuda.device(dev):
        bboxes = torch.empty((B, K, 4), dtype=torch.int32, device=dev)
        att = torch.empty((B, K, H, W), dtype=torch.float32, device=dev) if want_att else None
        ws = _zero_ws(dev, lib().rmnet_reg_att_map_workspace_bytes(B, K))
        _check_ws(lib().rmnet_warp_att_map_forward(prev_mask.data_ptr(), flow.data_ptr(), B, K, H, W,
                                               default_sampler() if sampler is None else sampler, float(prob_threshold),
                                               int(n_pts_threshold), int(n_bbox_loose_pixels), bboxes.data_ptr(),
                                               att.data_ptr() if want_att else None, ws.data_ptr(), ws.numel(),
                                               _stream(dev)), "warp_att_map_forward", ws)
    return att, bboxes


def regional_boxes(mask, flow=None, padded_frame=True, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64,
                   sampler=None, k_scan=0):
    """One launch: bounding boxes + /16 cell rectangles from the UNPADDED soft masks [B,K,H,W].
    flow=None, padded_frame=True  <->  pad_divide_by + get_att_map(masks) + interpolate(1/16)   (models/rmnet.py:212, :244-245)
    flow given, padded_frame=False <->  get_att_map(prev_mask, flow) + pad + interpolate(1/16)  (:431, :307, :356)
    -> (bboxes [B,K,4] int32, cell_rects [B,K,4] int32)"""
    _require(mask, "mask")
    B, K, H, W = mask.shape
    if flow is not None:
        _require(flow, "flow")
        if tuple(flow.shape) != (B, 2, H, W):
            raise RuntimeError(f"flow must be [{B},2,{H},{W}]")
    lw, uw, lh, uh = pad_amounts(H, W)
    dev = mask.device
    with torch.cuda.device(dev):
        bboxes = torch.empty((B, K, 4), dtype=torch.int32, device=dev)
        rects = torch.empty((B, K, 4), dtype=torch.int32, device=dev)
        ws = _zero_ws(dev, lib().rmnet_reg_att_map_workspace_bytes(B, K))
        _check_ws(lib().rmnet_regional_boxes_forward(mask.data_ptr(), flow.data_ptr() if flow is not None else None, B, K, H, W,
                                                 default_sampler() if sampler is None else sampler, float(prob_threshold),
                                                 int(n_pts_threshold), int(n_bbox_loose_pixels), lw, uw, lh, uh,
                                                 1 if padded_frame else 0, int(k_scan), bboxes.data_ptr(), rects.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), _stream(dev)), "regional_boxes_forward", ws)
    return bboxes, rects


def frame_regions(prev_mask, flow, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64, sampler=None, k_scan=0):
    """Both region descriptors of a frame in one pass over prev_mask [B,K,H,W] (unpadded) and flow [B,2,H,W]:
    -> (mem_bboxes, mem_rects, cur_bboxes, cur_rects), each [B,K,4] int32  (see regional_boxes for the two halves)."""
    _require(prev_mask, "prev_mask")
    _require(flow, "flow")
    B, K, H, W = prev_mask.shape
    if tuple(flow.shape) != (B, 2, H, W):
        raise RuntimeError(f"flow must be [{B},2,{H},{W}]")
    lw, uw, lh, uh = pad_amounts(H, W)
    dev = prev_mask.device
    with torch.cuda.device(dev):
        out = torch.empty((4, B, K, 4), dtype=torch.int32, device=dev)
        ws = _zero_ws(dev, lib().rmnet_reg_att_map_workspace_bytes(B, K))
        _check_ws(lib().rmnet_frame_regions_forward(prev_mask.data_ptr(), flow.data_ptr(), B, K, H, W,
                                                default_sampler() if sampler is None else sampler, float(prob_threshold),
                                                int(n_pts_threshold), int(n_bbox_loose_pixels), lw, uw, lh, uh, int(k_scan),
                                                out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                                ws.data_ptr(), ws.numel(), _stream(dev)), "frame_regions_forward", ws)
    return out[0], out[1], out[2], out[3]


def cell_rects(bboxes, pad_l, pad_t, h, w, skip_channel0_every=0):
    """Closed form of pad + F.interpolate(att_map, 1/16) for box-shaped att maps -> [..., 4] (cx0,cx1,cy0,cy1)."""
    _require(bboxes, "bboxes", torch.int32)
    dev = bboxes.device
    count = bboxes.numel() // 4
    with torch.cuda.device(dev):
        rects = torch.empty_like(bboxes)
        check(lib().rmnet_cell_rects_from_bboxes(bboxes.data_ptr(), count, int(pad_l), int(pad_t), h, w,
                                                 int(skip_channel0_every), rects.data_ptr(), _stream(dev)), "cell_rects")
    return rects


def mask_epilogue(dec_logits, K, frame_hw, channel_modes=None, new_mask=None, want_logit=True):
    """The tail of RMNet.segment / RMNet.forward after the decoder (models/rmnet.py:368-380, :289-302, :436-450) in one
    launch.  dec_logits [n,2,Hp,Wp] (decoder output on the padded frame), frame_hw = (H, W) of the unpadded frame,
    channel_modes: sequence of K ints (CH_KEEP / CH_ABSENT / CH_NEW) or None, new_mask [K,H,W] int32 (masks[i,t]) when a
    channel is CH_NEW.  -> (logit [1,K,H,W] or None, est_mask [1,K,H,W])"""
    import ctypes
    _require(dec_logits, "dec_logits")
    n, two, Hp, Wp = dec_logits.shape
    H, W = frame_hw
    lw, uw, lh, uh = pad_amounts(H, W)
    if two != 2 or (Hp, Wp) != (H + lh + uh, W + lw + uw):
        raise RuntimeError(f"dec_logits must be [n,2,{H + lh + uh},{W + lw + uw}]")
    if new_mask is not None:
        _require(new_mask, "new_mask", torch.int32)
        if tuple(new_mask.shape[-3:]) != (K, H, W):
            raise RuntimeError(f"new_mask must be [{K},{H},{W}]")
    modes = None
    if channel_modes is not None:
        if len(channel_modes) != K:
            raise RuntimeError("channel_modes must have K entries")
        modes = (ctypes.c_int * K)(*[int(m) for m in channel_modes])
    dev = dec_logits.device
    with torch.cuda.device(dev):
        logit = torch.empty((1, K, H, W), dtype=torch.float32, device=dev) if want_logit else None
        est = torch.empty((1, K, H, W), dtype=torch.float32, device=dev)
        check(lib().rmnet_mask_epilogue_forward(dec_logits.data_ptr(), n, K, H, W, lw, uw, lh, uh,
                                                ctypes.cast(modes, ctypes.c_void_p) if modes is not None else None,
                                                new_mask.data_ptr() if new_mask is not None else None,
                                                logit.data_ptr() if want_logit else None, est.data_ptr(), _stream(dev)),
              "mask_epilogue_forward")
    return logit, est


def pad_amounts(h, w, d=16):
    """utils/helpers.py:105-119 pad_divide_by -> (lw, uw, lh, uh)."""
    new_h = h + d - h % d if h % d > 0 else h
    new_w = w + d - w % d if w % d > 0 else w
    lh, uh = (new_h - h) // 2, (new_h - h) - (new_h - h) // 2
    lw, uw = (new_w - w) // 2, (new_w - w) - (new_w - w) // 2
    return lw, uw, lh, uh


class MemoryBank:
    """Preallocated region-compacted memory bank of one clip (replaces the reference's `keys`/`values` tensors
    and their per-frame torch.cat, models/rmnet.py:416-426).  Slot s holds object s+1."""

    def __init__(self, n_slots, h, w, max_frames, device, elem_format=ELEM_FP16):
        self.n_slots, self.h, self.w = int(n_slots), int(h), int(w)
        self.max_frames = int(max_frames)
        self.cap = ((self.max_frames * h * w + 63) // 64) * 64
        self.device = torch.device(device)
        self.elem_format = elem_format
        self.nbytes = lib().rmnet_bank_bytes(self.n_slots, self.cap)
        with torch.cuda.device(self.device):
            self.blob = torch.empty(self.nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self.blob.data_ptr()) % 1024
            self.ptr = self.blob.data_ptr() + off
            self._ws = torch.empty(lib().rmnet_memory_read_workspace_bytes(self.n_slots, h, w, self.cap) + 1024,
                                   dtype=torch.uint8, device=self.device)
            self._ws_ptr = self._ws.data_ptr() + ((-self._ws.data_ptr()) % 1024)
        self.frames_committed = 0
        self.has_temp = False
        self.reset()

    def reset(self):
        with torch.cuda.device(self.device):
            check(lib().rmnet_bank_reset(self.ptr, self.nbytes, self.n_slots, self.cap, _stream(self.device)), "bank_reset")
        self.frames_committed = 0
        self.has_temp = False

    def memorize(self, k4, v4, rects, commit):
        """k4 [n,128,h,w], v4 [n,512,h,w] (unmasked, per object), rects [n,4] int32 cell rectangles of this frame."""
        _require(k4, "k4")
        _require(v4, "v4")
        _require(rects, "rects", torch.int32)
        n = k4.shape[0]
        N = self.h * self.w
        if tuple(k4.shape) != (n, CK, self.h, self.w) or tuple(v4.shape) != (n, CV, self.h, self.w):
            raise RuntimeError("k4 / v4 shape mismatch")
        if rects.numel() != n * 4:
            raise RuntimeError("rects must be [n,4]")
        if self.frames_committed + 1 > self.max_frames:
            raise RuntimeError(f"memory bank full: {self.frames_committed} committed frames, capacity {self.max_frames}")
        with torch.cuda.device(self.device):
            check(lib().rmnet_bank_memorize(self.ptr, self.nbytes, self.n_slots, self.cap, k4.data_ptr(), CK * N, N,
                                            v4.data_ptr(), CV * N, N, rects.data_ptr(), n, self.h, self.w,
                                            self.elem_format, 1 if commit else 0, _stream(self.device)), "bank_memorize")
        if commit:
            self.frames_committed += 1
            self.has_temp = False
        else:
            self.has_temp = True

    def read(self, q_key, q_val, q_rects, n_obj, precision=RMNET_PREC_SPLIT3, impl=RMNET_IMPL_AUTO, stages=7, out=None):
        """q_key [128,h,w] / q_val [512,h,w] (one frame shared by all objects) or [n,128,h,w] / [n,512,h,w];
        q_rects [n,4] int32 or None (dense) -> mem_val [n,1024,h,w]."""
        _require(q_key, "q_key")
        _require(q_val, "q_val")
        N = self.h * self.w
        shared = q_key.dim() == 3
        if q_rects is not None:
            _require(q_rects, "q_rects", torch.int32)
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((n_obj, 2 * CV, self.h, self.w), dtype=torch.float32, device=self.device)
            check(lib().rmnet_bank_memory_read(self.ptr, self.nbytes, self.n_slots, self.cap, q_key.data_ptr(),
                                               q_val.data_ptr(), 0 if shared else CK * N,
                                               q_rects.data_ptr() if q_rects is not None else None, n_obj, self.h,
                                               self.w, self.elem_format, precision, impl, stages, out.data_ptr(), self._ws_ptr,
                                               self._ws.numel() - 1024, _stream(self.device)), "bank_memory_read")
        return out

    def read_plan(self, n_obj):
        """The work plan of the last read / frame step on this bank (sched.cuh): (ns [n_obj], per-CTA piece lists) with
        pieces as tuples (object, query_tile, half, slot, first_tile, tiles, stored_cells).  Introspection / tests."""
        import numpy as np
        ns = np.zeros(n_obj, np.int32)
        hdr = np.zeros((256, 2), np.int32)
        cap = 1 << 16
        pcs = np.zeros((cap, 4), np.int32)
        n_ctas = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            check(lib().rmnet_memory_read_plan_host(self._ws_ptr, n_obj, self.h, self.w, ns.ctypes.data, hdr.ctypes.data,
                                                    pcs.ctypes.data, cap, ctypes.byref(n_ctas), _stream(self.device)),
                  "memory_read_plan_host")
        lists = []
        for c in range(n_ctas.value):
            n, off = int(hdr[c, 0]), int(hdr[c, 1])
            lists.append([(int(v[0]) & 255, (int(v[0]) >> 8) & 255, (int(v[0]) >> 16) & 15, (int(v[0]) >> 20) & 0xf,
                           int(v[1]), int(v[2]), int(v[3])) for v in pcs[off:off + n]])
        return ns, lists

    def stats(self):
        import numpy as np
        out = np.zeros((self.n_slots, 8), np.int32)
        with torch.cuda.device(self.device):
            check(lib().rmnet_bank_stats_host(self.ptr, self.n_slots, self.cap, out.ctypes.data, _stream(self.device)),
                  "bank_stats")
        return out


_reader_ws = {}


def memory_reader_forward(m_key, m_val, q_key, q_val, precision=RMNET_PREC_SPLIT3, impl=RMNET_IMPL_AUTO,
                          elem_format=ELEM_FP16, want_p=False):
    """Literal MemoryReader.forward (models/rmnet.py:147-165) -> mem_val [n,1024,h,w] (dense, region-agnostic);
    with want_p also the reference's second output p [n,T*h*w,h*w] -> (mem_val, p)."""
    for t, nm in ((m_key, "m_key"), (m_val, "m_val"), (q_key, "q_key"), (q_val, "q_val")):
        _require(t, nm)
    n, ck, T, h, w = m_key.shape
    if ck != CK or m_val.shape[1] != CV or tuple(q_key.shape) != (n, CK, h, w) or tuple(q_val.shape) != (n, CV, h, w):
        raise RuntimeError("MemoryReader shapes must be m_key [n,128,T,h,w], m_val [n,512,T,h,w], q_key [n,128,h,w], q_val [n,512,h,w]")
    dev = m_key.device
    with torch.cuda.device(dev):
        need = lib().rmnet_memory_reader_workspace_bytes(n, T, h, w) + 1024
        key = (dev.index, _stream(dev))
        ws = _reader_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _reader_ws[key] = ws
        ptr = ws.data_ptr() + ((-ws.data_ptr()) % 1024)
        out = torch.empty((n, 2 * CV, h, w), dtype=torch.float32, device=dev)
        p = torch.empty((n, T * h * w, h * w), dtype=torch.float32, device=dev) if want_p else None
        check(lib().rmnet_memory_reader_forward(m_key.data_ptr(), m_val.data_ptr(), q_key.data_ptr(), q_val.data_ptr(),
                                                n, T, h, w, elem_format, precision, impl, out.data_ptr(),
                                                p.data_ptr() if want_p else None, ptr, ws.numel() - 1024, _stream(dev)),
              "memory_reader_forward")
    return (out, p) if want_p else out


def update_optical_flow_cuda(of, m1, m2):
    """Device-tensor variant of update_optical_flow: of [H,W,2] CUDA f32; m1, m2 anything convertible to 6 floats."""
    import numpy as np
    _require(of, "of")
    H, W = of.shape[:2]
    a1 = np.ascontiguousarray(np.asarray(m1, dtype=np.float32).reshape(6))
    a2 = np.ascontiguousarray(np.asarray(m2, dtype=np.float32).reshape(6))
    with torch.cuda.device(of.device):
        out = torch.empty_like(of)
        check(lib().rmnet_update_optical_flow(of.data_ptr(), a1.ctypes.data, a2.ctypes.data, H, W, out.data_ptr(),
                                              _stream(of.device)), "update_optical_flow")
    return out


_flow_tls = __import__("threading").local()


def _cuda_usable_here():
    """True when this process may launch on a CUDA device right now: a device exists, torch's CUDA context is already
    initialised in THIS process (never initialise one just for a 6 MB elementwise op) and we are not a forked child of a
    process that had initialised CUDA (a DataLoader worker: utils/data_transforms.py:293-302 runs there)."""
    if not torch.cuda.is_available() or not torch.cuda.is_initialized():
        return False
    bad_fork = getattr(torch.cuda, "_is_in_bad_fork", None)
    return not (bad_fork() if callable(bad_fork) else False)


def update_optical_flow(of, m1, m2, device=None):
    """flow_affine_transformation.update_optical_flow(of, M1, M2) (flow_affine_transformation.cpp:39-85):
    NumPy in, NumPy out, bit-exact with the reference extension either way it runs:

      * on the GPU (host buffers staged through a per-thread device scratch) when this process already has a usable CUDA
        context, or when device="cuda" is forced;
      * by the library's plain-C entry point rmnet_update_optical_flow_cpu otherwise -- in particular inside forked
        DataLoader workers, which is where the reference calls this op -- or when device="cpu" is forced.

    Unlike the reference (no validation, .cpp:45-55) dtype / shape / contiguity are checked and converted (the
    reference reads a float64 zeros array as float32 when a flow file is missing, utils/data_loaders.py:54-55)."""
    import numpy as np
    of = np.ascontiguousarray(of, dtype=np.float32)
    if of.ndim != 3 or of.shape[2] != 2:
        raise ValueError("optical flow must be [H,W,2]")
    a1 = np.ascontiguousarray(np.asarray(m1, dtype=np.float32).reshape(-1)[:6])
    a2 = np.ascontiguousarray(np.asarray(m2, dtype=np.float32).reshape(-1)[:6])
    if a1.size != 6 or a2.size != 6:
        raise ValueError("affine matrices must be 2x3")
    H, W = of.shape[:2]
    out = np.empty_like(of)
    use_cuda = _cuda_usable_here() if device is None else (str(device) != "cpu")
    if not use_cuda:
        check(lib().rmnet_update_optical_flow_cpu(of.ctypes.data, a1.ctypes.data, a2.ctypes.data, H, W, out.ctypes.data),
              "update_optical_flow_cpu")
        return out
    if not torch.cuda.is_available():
        raise RuntimeError("rmnet_b200.update_optical_flow(device='cuda') needs a CUDA device")
    dev = torch.device("cuda", torch.cuda.current_device())
    need = 2 * of.nbytes
    cache = _flow_tls.__dict__.setdefault("scratch", {})      # per thread: concurrent callers never share a staging buffer
    sc = cache.get(dev.index)
    if sc is None or sc.numel() < need:
        sc = torch.empty(need, dtype=torch.uint8, device=dev)
        cache[dev.index] = sc
    check(lib().rmnet_update_optical_flow_host(of.ctypes.data, a1.ctypes.data, a2.ctypes.data, H, W, out.ctypes.data,
                                               sc.data_ptr(), sc.numel(), _stream(dev)), "update_optical_flow_host")
    return out
