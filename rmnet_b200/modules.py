"""Host-side mirror of the reference's operator interface for the hot path: same class names, forward()
signatures, argument meaning and error behaviour as

  extensions/reg_att_map_generator/__init__.py:14-33   RegionalAttentionMapGenerator(.Function)
  models/rmnet.py:143-165                               MemoryReader
  models/rmnet.py:252-287                               RMNet.warp / RMNet.get_att_map
  models/rmnet.py:239-248, :355-361                     the regional parts of RMNet.memorize / RMNet.segment

so they drop in under core/inference.py.  Compute is exclusively librmnet_b200.so.
"""
import torch

from . import ops
from ._lib import ELEM_FP16, RMNET_IMPL_AUTO, RMNET_PREC_SPLIT3


class RegionalAttentionMapGeneratorFunction(torch.autograd.Function):
    """extensions/reg_att_map_generator/__init__.py:14-24 (the backward returns ones, like the reference)."""

    @staticmethod
    def forward(ctx, mask, prob_threshold, n_pts_threshold, n_bbox_loose_pixels):
        att_map, bbox = ops.reg_att_map_forward(mask, prob_threshold, n_pts_threshold, n_bbox_loose_pixels)
        ctx.mark_non_differentiable(bbox)
        return att_map, bbox

    @staticmethod
    def backward(ctx, grad_att_map, grad_bbox):
        return torch.ones_like(grad_att_map), None, None, None


class RegionalAttentionMapGenerator(torch.nn.Module):
    """extensions/reg_att_map_generator/__init__.py:27-33."""

    def forward(self, mask, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64):
        return RegionalAttentionMapGeneratorFunction.apply(mask, prob_threshold, n_pts_threshold, n_bbox_loose_pixels)


class MemoryReader(torch.nn.Module):
    """models/rmnet.py:143-165.  forward(m_key, m_val, q_key, q_val) -> (mem_val, p).

    `p` ([n, T*h*w, h*w], 210 MB per object at 480p / T=20) is never materialised by the fused kernel; the only
    caller ignores it (`m4, viz = self.memory(...)`, models/rmnet.py:361), so by default None is returned in its place.
    `MemoryReader(return_p=True)` restores the reference's tuple (a separate fp32 kernel writes p)."""

    def __init__(self, precision=RMNET_PREC_SPLIT3, impl=RMNET_IMPL_AUTO, elem_format=ELEM_FP16, return_p=False):
        super().__init__()
        self.precision, self.impl, self.elem_format, self.return_p = precision, impl, elem_format, return_p

    def forward(self, m_key, m_val, q_key, q_val):
        if any(t.requires_grad for t in (m_key, m_val, q_key, q_val)) and torch.is_grad_enabled():
            raise RuntimeError("rmnet_b200.MemoryReader is inference-only (run under torch.no_grad())")
        res = ops.memory_reader_forward(m_key.contiguous(), m_val.contiguous(), q_key.contiguous(), q_val.contiguous(),
                                        self.precision, self.impl, self.elem_format, want_p=self.return_p)
        return res if self.return_p else (res, None)


def warp(img0, flow):
    """RMNet.warp(self, img0, flow) -> (img1, mask), models/rmnet.py:252-278."""
    return ops.warp(img0.contiguous(), flow.contiguous())


def get_att_map(prev_mask, flow=None, prob_threshold=0.5, n_pts_threshold=10, n_bbox_loose_pixels=64):
    """RMNet.get_att_map(self, prev_mask, flow=None) -> (att_map, bbox), models/rmnet.py:280-287."""
    if flow is None:
        att, bbox = ops.reg_att_map_forward(prev_mask.contiguous(), prob_threshold, n_pts_threshold, n_bbox_loose_pixels)
    else:
        att, bbox = ops.warp_att_map_forward(prev_mask.contiguous(), flow.contiguous(), prob_threshold, n_pts_threshold,
                                             n_bbox_loose_pixels)
    return att, bbox


class RegionalMemory:
    """The fused regional path of one clip (batch 1, as core/inference.py:26 runs it): owns the preallocated bank
    and exposes the two per-frame steps with the tensors the reference has at hand at those points.

      memorize(k4, v4, masks_padded, commit)  <->  models/rmnet.py:239-248 (+ the cat at :416-426)
      read(k4q, v4q, prev_mask, flow, n_objects)  <->  :431 get_att_map + :307 pad + :355-361
    """

    def __init__(self, n_objects, frame_hw, max_frames, device, precision=RMNET_PREC_SPLIT3, impl=RMNET_IMPL_AUTO,
                 elem_format=ELEM_FP16, scan_all_channels=False):
        H, W = frame_hw
        self.lw, self.uw, self.lh, self.uh = ops.pad_amounts(H, W)
        self.Hp, self.Wp = H + self.lh + self.uh, W + self.lw + self.uw
        self.h, self.w = self.Hp // 16, self.Wp // 16
        self.n = int(n_objects)
        self.precision, self.impl = precision, impl
        # The reference scans all K-1 mask channels but only ever uses the boxes of objects 1..n (models/rmnet.py:326-331);
        # channels above n_max_objects are forced to probability ~1e-7 by its frame loop (:444-448), i.e. absent objects.
        # By default only channels 1..n are read and the rest are reported as absent (identical boxes on such inputs).
        self.k_scan = 0 if scan_all_channels else self.n + 1
        self.bank = ops.MemoryBank(self.n, self.h, self.w, max_frames, device, elem_format)
        self._boxes = None
        # region-kernel workspace of step(): zero-filled once, left zeroed by every launch (self-cleaning); owned by the
        # clip (not by the stream) so that a captured step and an eager step share it
        with torch.cuda.device(self.bank.device):
            self._box_ws = torch.zeros(max(4096, ops.lib().rmnet_reg_att_map_workspace_bytes(1, 64)), dtype=torch.uint8,
                                       device=self.bank.device)

    def memorize(self, k4, v4, masks, commit):
        """k4 [n,128,h,w], v4 [n,512,h,w]: kv_memory outputs (models/rmnet.py:236); masks [1,K,H,W]: the UNPADDED soft
        masks RMNet.memorize receives (the zero padding of :212 is applied analytically).  Returns bboxes [1,K,4] in
        padded coordinates, exactly what the reference's memorize returns (:244, :250)."""
        bboxes, rects = ops.regional_boxes(masks, None, padded_frame=True, k_scan=self.k_scan)
        self.bank.memorize(k4, v4, rects[0, 1:self.n + 1], commit)
        return bboxes

    def read(self, k4q, v4q, prev_mask, flow, out=None):
        """k4q [128,h,w], v4q [512,h,w]: kv_query outputs of the current frame (:315); prev_mask [1,K,H,W] and
        flow [1,2,H,W] in UNPADDED coordinates (:431).  Returns (m4 [n,1024,h,w], curr_bbox [1,K,4])."""
        bbox, rects = ops.regional_boxes(prev_mask, flow, padded_frame=False, k_scan=self.k_scan)
        m4 = self.bank.read(k4q, v4q, rects[0, 1:self.n + 1], self.n, self.precision, self.impl, out=out)
        return m4, bbox


    def step(self, k4, v4, prev_mask, flow, k4q, v4q, commit, out=None):
        """One frame of the reference's loop body (models/rmnet.py:414-432 minus the convs) in ONE library call
        (rmnet_frame_step: regions of both sides from one pass over prev_mask, pack k4/v4, regional read).
        Returns (m4 [n,1024,h,w], prev_bbox [1,K,4] padded coords, curr_bbox [1,K,4] raw coords)."""
        bank = self.bank
        for t, nm in ((k4, "k4"), (v4, "v4"), (prev_mask, "prev_mask"), (flow, "flow"), (k4q, "k4q"), (v4q, "v4q")):
            ops._require(t, nm)
        _, K, H, W = prev_mask.shape
        if (H + self.lh + self.uh, W + self.lw + self.uw) != (self.Hp, self.Wp) or tuple(flow.shape) != (1, 2, H, W):
            raise RuntimeError("prev_mask / flow shape mismatch")
        if k4.shape[0] != self.n or k4q.numel() != 128 * self.h * self.w or v4q.numel() != 512 * self.h * self.w:
            raise RuntimeError("k4 / k4q shape mismatch")
        if bank.frames_committed + 1 > bank.max_frames:
            raise RuntimeError(f"memory bank full: {bank.frames_committed} committed frames, capacity {bank.max_frames}")
        dev = bank.device
        if self._boxes is None or self._boxes.shape[1] != K:
            self._boxes = torch.empty((4, K, 4), dtype=torch.int32, device=dev)
        if out is None:
            out = torch.empty((self.n, 1024, self.h, self.w), dtype=torch.float32, device=dev)
        ws = self._box_ws
        with torch.cuda.device(dev):   # the caller's current device is left as it was
            rc = ops.lib().rmnet_frame_step(
                bank.ptr, bank.nbytes, bank.n_slots, bank.cap, prev_mask.data_ptr(), flow.data_ptr(), K, H, W,
                ops.default_sampler(), 0.5, 10, 64, self.lw, self.uw, self.lh, self.uh, self.k_scan, k4.data_ptr(), v4.data_ptr(),
                k4q.data_ptr(), v4q.data_ptr(), self.n, bank.elem_format, self.precision, self.impl, 1 if commit else 0,
                self._boxes.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), bank._ws_ptr, bank._ws.numel() - 1024,
                torch.cuda.current_stream(dev).cuda_stream)
            if rc != 0:
                ws.zero_()     # the self-cleaning region workspace must be zero before the next launch, whatever happened
            ops.check(rc, "frame_step")
        if commit:
            bank.frames_committed += 1
            bank.has_temp = False
        else:
            bank.has_temp = True
        return out, self._boxes[0:1], self._boxes[2:3]

    def capture_step(self, k4, v4, prev_mask, flow, k4q, v4q, commit, out=None):
        """step() as a replayable CUDA graph over static input tensors: see CapturedStep."""
        return CapturedStep(self, k4, v4, prev_mask, flow, k4q, v4q, commit, out)


class CapturedStep:
    """RegionalMemory.step captured once into a CUDA graph (the four chained kernels with their programmatic-dependency
    edges) and replayed with ONE launch per frame: the per-frame host cost drops from four kernel launches plus the
    Python / ctypes plumbing to a cudaGraphLaunch.  The input tensors given to capture() are static: write the next
    frame's data into them (copy_ or let the producing conv write there), then call replay().

        cs = rm.capture_step(k4, v4, prev_mask, flow, k4q, v4q, commit=False)
        ...fill the static inputs...; m4, prev_bbox, curr_bbox = cs.replay()
    """

    def __init__(self, rm, k4, v4, prev_mask, flow, k4q, v4q, commit, out=None):
        self.rm, self.commit = rm, bool(commit)
        self.inputs = (k4, v4, prev_mask, flow, k4q, v4q)
        dev = rm.bank.device
        if out is None:
            out = torch.empty((rm.n, 1024, rm.h, rm.w), dtype=torch.float32, device=dev)
        self.out = out
        # warm-up on a side stream (allocates the box buffer, sets kernel attributes); always without commit: it only
        # rewrites the temporary frame, exactly what the first replay will do again
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            rm.step(k4, v4, prev_mask, flow, k4q, v4q, commit=False, out=out)
        torch.cuda.current_stream(dev).wait_stream(side)
        state = (rm.bank.frames_committed, rm.bank.has_temp)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            _, self.prev_bbox, self.curr_bbox = rm.step(k4, v4, prev_mask, flow, k4q, v4q, commit=self.commit, out=out)
        rm.bank.frames_committed, rm.bank.has_temp = state   # capture enqueued nothing: undo its host bookkeeping

    def replay(self):
        bank = self.rm.bank
        if bank.frames_committed + 1 > bank.max_frames:
            raise RuntimeError(f"memory bank full: {bank.frames_committed} committed frames, capacity {bank.max_frames}")
        self.graph.replay()
        if self.commit:
            bank.frames_committed += 1
            bank.has_temp = False
        else:
            bank.has_temp = True
        return self.out, self.prev_bbox, self.curr_bbox


_ORIG = "_rmnet_b200_originals"
_replica_loops = {}   # see fused_forward


def fused_forward(self, frames, masks, optical_flows, n_objects, memorize_every, device=None):
    """RMNet.forward (models/rmnet.py:385) bound by install(): the GPU-resident RegionalFrameLoop built once per model
    instance from the instance's OWN encoder_memory / kv_memory / encoder_query / kv_query / decoder.  Inference, batch 1
    (core/inference.py:26, core/test.py) only; a batched or differentiable call is handed to the reference's own forward
    untouched (training is outside this library's path)."""
    differentiable = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
    if frames.size(0) != 1 or differentiable:
        return getattr(type(self), _ORIG)["forward"](self, frames, masks, optical_flows, n_objects, memorize_every, device)
    from .frame_loop import RegionalFrameLoop
    opts = getattr(type(self), "_rmnet_b200_options")
    if getattr(self, "_is_replica", False):
        # nn.DataParallel over several GPUs (core/inference.py:36-37 on a multi-GPU box) calls a fresh REPLICA of the model on
        # every forward; its device-0 tensors alias the original parameters, so the loop (bank, captured graphs) is kept in
        # a small module-level cache keyed by a parameter's storage instead of on the short-lived replica
        key = (self.kv_memory.key_conv.weight.data_ptr(), id(opts))
        loop = _replica_loops.pop(key, None)
        if loop is None:
            loop = RegionalFrameLoop.from_rmnet(self, **opts)
            loop._options = opts
        _replica_loops[key] = loop
        while len(_replica_loops) > 2:
            _replica_loops.pop(next(iter(_replica_loops)))
    else:
        loop = self.__dict__.get("_rmnet_b200_loop")
        if loop is None or loop._options is not opts:
            loop = RegionalFrameLoop.from_rmnet(self, **opts)
            loop._options = opts
            object.__setattr__(self, "_rmnet_b200_loop", loop)
    return loop.forward(frames, masks, optical_flows, n_objects, memorize_every, device)


def install(models_rmnet_module, fused=True, use_graph=None, output="reference", precision=RMNET_PREC_SPLIT3,
            impl=RMNET_IMPL_AUTO, elem_format=ELEM_FP16):
    """Rebind the reference's names so that an unmodified core/inference.py builds an RMNet that runs on this library
    (RMNet.__init__ looks the classes up at construction time, models/rmnet.py:187-189; core/inference.py:17 imports the
    RMNet class itself, so its methods are patched in place):

      MemoryReader, RegionalAttentionMapGenerator, RMNet.warp, RMNet.get_att_map  -> the literal operator drop-ins;
      RMNet.forward (fused=True)  -> fused_forward: the whole frame loop on the fused regional path.  RMNet.memorize /
      RMNet.segment keep the reference's bodies (they are only reached through the reference's own forward, i.e. batched /
      training calls, where every op of theirs that is on the path already resolves to the literal drop-ins above).

    use_graph: replay each frame as one CUDA graph (None = on unless RMNET_B200_GRAPH=0); output: see RegionalFrameLoop.
    uninstall() restores the reference's own attributes."""
    import os
    cls = models_rmnet_module.RMNet
    if not hasattr(cls, _ORIG):
        setattr(cls, _ORIG, {"forward": cls.forward, "warp": cls.warp, "get_att_map": cls.get_att_map,
                             "MemoryReader": models_rmnet_module.MemoryReader,
                             "RegionalAttentionMapGenerator": models_rmnet_module.RegionalAttentionMapGenerator})
    models_rmnet_module.MemoryReader = MemoryReader
    models_rmnet_module.RegionalAttentionMapGenerator = RegionalAttentionMapGenerator
    cls.warp = lambda self, img0, flow: warp(img0, flow)
    cls.get_att_map = lambda self, prev_mask, flow=None: get_att_map(prev_mask, flow)
    if use_graph is None:
        use_graph = os.environ.get("RMNET_B200_GRAPH", "1") != "0"
    cls._rmnet_b200_options = dict(use_graph=bool(use_graph), output=output, precision=precision, impl=impl, elem_format=elem_format)
    if fused:
        cls.forward = fused_forward
    return models_rmnet_module


def uninstall(models_rmnet_module):
    """Undo install(): the reference's own forward / warp / get_att_map / MemoryReader / generator are back."""
    cls = models_rmnet_module.RMNet
    orig = getattr(cls, _ORIG, None)
    if orig is None:
        return models_rmnet_module
    cls.forward, cls.warp, cls.get_att_map = orig["forward"], orig["warp"], orig["get_att_map"]
    models_rmnet_module.MemoryReader = orig["MemoryReader"]
    models_rmnet_module.RegionalAttentionMapGenerator = orig["RegionalAttentionMapGenerator"]
    delattr(cls, _ORIG)
    return models_rmnet_module
