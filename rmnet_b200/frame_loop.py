"""GPU-resident mirror of RMNet.forward's frame loop (models/rmnet.py:385-452) for one clip (batch 1, as
core/inference.py:26 runs it), built on the fused per-frame calls of this library.

The reference keeps `est_masks` on the host when one GPU is visible and moves 18 MB each way per frame (:388-392, :412,
:450), rebuilds the warp grid on the CPU (:257-262), re-copies the whole memory bank every frame (:416-426) and runs
about fifty small ATen kernels around its three conv nets.  Here a frame is

    memorize_net  ->  query_net  ->  RegionalMemory.step (ONE call: 4 chained kernels)  ->  decoder_net  ->  mask_epilogue (1 kernel)

with everything resident on the device, and (use_graph=True) the whole frame body -- conv nets included -- replayed as
ONE CUDA graph per (object count, frame shape, commit flag).  The three conv nets stay the reference's (out of scope,
SURVEY 2) and are passed in as callables:

    memorize_net(frame_p [1,3,Hp,Wp], obj_masks [n,Hp,Wp], other_masks [n,Hp,Wp]) -> (k4 [n,128,h,w], v4 [n,512,h,w])
        = kv_memory(encoder_memory(f, m, o))                                  models/rmnet.py:219-236
    query_net(frame_p [1,3,Hp,Wp]) -> (k4q [1,128,h,w], v4q [1,512,h,w], ctx)
        = kv_query(encoder_query(frame)), ctx = (r3, r2)                      :311-315
    decoder_net(m4 [n,1024,h,w], ctx) -> logits [n,2,Hp,Wp]
        = decoder(m4, r3e, r2e)                                               :366

`RegionalFrameLoop.from_rmnet(net)` builds them from an RMNet instance's own sub-modules; `rmnet_b200.install()` binds
`RMNet.forward` to that loop, so an unmodified core/inference.py runs it.
"""
import torch
import torch.nn.functional as F

from . import ops
from ._lib import CH_ABSENT, CH_KEEP, CH_NEW, ELEM_FP16, RMNET_IMPL_AUTO, RMNET_PREC_SPLIT3, lib
from .modules import RegionalMemory


# kernels of this library launched through CUDA-graph replays by every loop of this process (rmnet_launch_count() counts
# only eager launches); [0] so that importers see updates
graph_launch_total = [0]


def object_batches(masks_p, n):
    """The per-object mask batches of RMNet.memorize (models/rmnet.py:219-229): m = the object's soft mask, o = the other
    objects' masks summed (before + after, in the reference's order) and clamped.  masks_p [1,K,Hp,Wp] -> ([n,Hp,Wp], [n,Hp,Wp])"""
    m, o = [], []
    for k in range(1, n + 1):
        m.append(masks_p[0, k].unsqueeze(0))
        o.append((torch.sum(masks_p[0, 1:k].unsqueeze(0), dim=1) + torch.sum(masks_p[0, k + 1:n + 1].unsqueeze(0), dim=1)).clamp(0, 1))
    return torch.cat(m, dim=0), torch.cat(o, dim=0)


def memorize_schedule(n_frames, memorize_every, n_objects_per_frame):
    """Host logic of models/rmnet.py:405-408, :424: -> (to_memorize, contains_new_objects, commit[t] for t = 1..F-1, #commits).
    Frame t-1 is committed to the permanent memory iff it is a multiple of memorize_every or a frame whose object count
    differs from its predecessor's."""
    to_memorize = set(range(0, n_frames, memorize_every))
    new_at = {j for j in range(1, n_frames) if n_objects_per_frame[j] != n_objects_per_frame[j - 1]}
    commit = {t: ((t - 1) in to_memorize or (t - 1) in new_at) for t in range(1, n_frames)}
    return to_memorize, new_at, commit, sum(commit.values())


def channel_modes(K, n_max, existing, labels_in_gt):
    """Host logic of models/rmnet.py:436-448 for one frame: `existing` (list, updated in place) are the objects seen so
    far; `labels_in_gt` = the labels present in masks[i, t] when t introduces objects, else None.
    -> per-channel modes for rmnet_mask_epilogue_forward (CH_KEEP / CH_NEW / CH_ABSENT)."""
    modes = [CH_KEEP] * K
    if labels_in_gt is not None:
        for j in labels_in_gt:
            if j not in existing:
                existing.append(j)
                modes[j] = CH_NEW                      # logit := masks[i,t,j] * 32.0605 - 16.1181  (:442)
    for j in range(n_max + 1):
        if j not in existing:
            modes[j] = CH_ABSENT                       # logit := -16.1181                          (:448)
    return modes


class _ClipState:
    """Everything of the loop that survives from clip to clip for one (n, K, H, W): the preallocated bank, the static
    frame buffers and the captured frame graphs (their kernels hold the bank's and the buffers' addresses)."""

    def __init__(self, loop, n, K, H, W, max_frames, dev):
        self.n, self.K, self.H, self.W, self.dev = n, K, H, W, dev
        self.max_frames = max_frames
        self.rm = RegionalMemory(n, (H, W), max_frames=max_frames, device=dev, precision=loop.precision, impl=loop.impl,
                                 elem_format=loop.elem_format)
        f32 = dict(dtype=torch.float32, device=dev)
        self.prev_frame = torch.zeros((1, 3, H, W), **f32)
        self.cur_frame = torch.zeros((1, 3, H, W), **f32)
        self.flow = torch.zeros((1, 2, H, W), **f32)
        self.prev_mask = torch.zeros((1, K, H, W), **f32)
        self.graphs = {}          # (commit, modes, want_logit) -> (CUDAGraph, outputs)
        self.pool = None
        self.side = torch.cuda.Stream(dev)   # the query encoder's branch of a frame


class RegionalFrameLoop:
    """forward(frames, masks, optical_flows, n_objects, memorize_every) -> est_masks [1,F,K,H,W], the signature and the
    semantics of RMNet.forward (models/rmnet.py:385) for batch 1.

    output: "device" keeps est_masks on the GPU; "host" returns a (pinned) CPU tensor, each frame copied back
    asynchronously while the next one is computed; "reference" follows the reference's rule (:388-392: a CUDA tensor iff
    more than one GPU is visible and no device is given, else a CPU tensor)."""

    def __init__(self, memorize_net, query_net, decoder_net, precision=RMNET_PREC_SPLIT3, impl=RMNET_IMPL_AUTO,
                 elem_format=ELEM_FP16, use_graph=False, output="device", min_bank_frames=24, overlap_query=True):
        self.memorize_net, self.query_net, self.decoder_net = memorize_net, query_net, decoder_net
        self.precision, self.impl, self.elem_format = precision, impl, elem_format
        self.use_graph, self.output = bool(use_graph), output
        self.overlap_query = bool(overlap_query)
        self.max_cached_states = 8                    # (object count, frame shape) combinations kept warm between clips
        self.min_bank_frames = int(min_bank_frames)   # bank capacity floor: 24 frames hold a 115-frame clip at memorize_every = 5
        self.last_bboxes = None   # [(prev_bbox, curr_bbox)] of the last clip (keep_bboxes=True), for inspection / tests
        self.last_logits = None   # [logit [1,K,H,W]] of the last clip (keep_logits=True): the return values of RMNet.segment + overrides
        self.last_frame_ms = None  # per-frame device time of the last clip (time_frames=True)
        self.record_frame_times = False   # forward() records per-frame CUDA events into last_frame_ms (bench)
        self.graph_launches = 0    # kernels of this library launched through graph replays (rmnet_launch_count() only sees eager calls)
        self._states = {}

    @classmethod
    def from_rmnet(cls, net, **kw):
        """The three conv callables from an RMNet instance's own sub-modules (the reference's cuDNN code, untouched)."""
        def memorize_net(frame_p, m, o):
            f = frame_p.expand(m.shape[0], -1, -1, -1).contiguous()                # :222, :232 (cat of n copies)
            r4 = net.encoder_memory(f, m, o)[0]                                   # :234
            return net.kv_memory(r4)                                              # :236

        def query_net(frame_p):
            r4, r3, r2, _, _ = net.encoder_query(frame_p)                         # :311
            k4, v4 = net.kv_query(r4)                                             # :315
            return k4, v4, (r3, r2)

        def decoder_net(m4, ctx):
            n = m4.shape[0]
            r3e = ctx[0].expand(n, -1, -1, -1).contiguous()                       # :334-335, :347-349
            r2e = ctx[1].expand(n, -1, -1, -1).contiguous()
            return net.decoder(m4, r3e, r2e)                                      # :366

        return cls(memorize_net, query_net, decoder_net, **kw)

    # ------------------------------------------------------------------------------------------------------------
    def _state(self, n, K, H, W, n_commits, dev):
        key = (n, K, H, W, dev.index)
        st = self._states.pop(key, None)                 # (re-inserted below: the dict is kept in least-recently-used order)
        if st is None or st.max_frames < n_commits + 1:
            # capacity in whole multiples of 8 frames so that clips of similar length reuse the bank and its graphs
            cap = max(self.min_bank_frames, ((n_commits + 1 + 7) // 8) * 8)
            st = _ClipState(self, n, K, H, W, cap, dev)
        else:
            st.rm.bank.reset()
        self._states[key] = st
        while len(self._states) > self.max_cached_states:   # each state holds a bank and the graphs' private memory pool
            self._states.pop(next(iter(self._states)))
        return st

    def _frame_body(self, st, commit, modes, new_mask, want_logit):
        """One frame on the current stream, reading the state's static buffers: -> (logit or None, est, prev_bbox, curr_bbox)."""
        n, K, H, W = st.n, st.K, st.H, st.W
        rm = st.rm
        pad = (rm.lw, rm.uw, rm.lh, rm.uh)
        # The query branch (encoder_query + kv_query of the CURRENT frame, batch 1: latency-bound) does not depend on the
        # memorise branch (encoder_memory + kv_memory of the previous frame and mask, batch n): they run on two streams
        # (two branches of the captured graph) and join before the fused step.  Same kernels, same bits.
        cur = torch.cuda.current_stream(st.dev)
        if self.overlap_query:
            st.side.wait_stream(cur)
        with torch.cuda.stream(st.side if self.overlap_query else cur):
            k4q, v4q, ctx = self.query_net(F.pad(st.cur_frame, pad))                # :307-315
            k4q0, v4q0 = k4q[0].contiguous(), v4q[0].contiguous()
        masks_p = F.pad(st.prev_mask, pad)                                          # :212
        frame_p = F.pad(st.prev_frame, pad)
        m, o = object_batches(masks_p, n)                                           # :219-229
        k4, v4 = self.memorize_net(frame_p, m, o)                                   # :234-236
        if self.overlap_query:
            cur.wait_stream(st.side)
            for x in (k4q, v4q, k4q0, v4q0) + tuple(ctx or ()):
                if isinstance(x, torch.Tensor):
                    x.record_stream(cur)
        m4, prev_bbox, curr_bbox = rm.step(k4.contiguous(), v4.contiguous(), st.prev_mask, st.flow, k4q0, v4q0,
                                           commit=commit)                           # :239-248, :416-426, :431, :355-361
        logits = self.decoder_net(m4, ctx)                                          # :366
        logit, est = ops.mask_epilogue(logits.contiguous(), K, (H, W), modes, new_mask, want_logit=want_logit)   # :368-380, :289-302, :436-450
        return logit, est, prev_bbox, curr_bbox

    def _graph_frame(self, st, commit, modes, want_logit):
        key = (bool(commit), tuple(modes), bool(want_logit))
        ent = st.graphs.get(key)
        rm = st.rm
        if ent is None:
            dev = st.dev
            state = (rm.bank.frames_committed, rm.bank.has_temp)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):   # warm-up: cuDNN plans / workspaces, kernel attributes; never commits (it only rewrites the temporary frame)
                self._frame_body(st, False, modes, None, want_logit)
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            before = int(lib().rmnet_launch_count())
            with torch.cuda.graph(g, pool=st.pool):
                outs = self._frame_body(st, commit, modes, None, want_logit)
            captured = int(lib().rmnet_launch_count()) - before     # this library's kernels inside the graph
            if st.pool is None:
                st.pool = g.pool()
            rm.bank.frames_committed, rm.bank.has_temp = state      # warm-up and capture did no lasting bank bookkeeping
            ent = (g, outs, captured)
            st.graphs[key] = ent
        g, outs, captured = ent
        if rm.bank.frames_committed + 1 > rm.bank.max_frames:
            raise RuntimeError(f"memory bank full: {rm.bank.frames_committed} committed frames, capacity {rm.bank.max_frames}")
        g.replay()
        self.graph_launches += captured
        graph_launch_total[0] += captured
        if commit:
            rm.bank.frames_committed += 1
            rm.bank.has_temp = False
        else:
            rm.bank.has_temp = True
        return outs

    @torch.no_grad()
    def forward(self, frames, masks, optical_flows, n_objects, memorize_every, device=None, teacher_masks=None,
                keep_logits=False, keep_bboxes=False, time_frames=False):
        """teacher_masks [1,F,K,H,W] (tests): frame t is segmented from teacher_masks[:, t-1] instead of the loop's own
        est_masks[:, t-1] (teacher forcing against another implementation's masks)."""
        if frames.shape[0] != 1:
            raise RuntimeError("RegionalFrameLoop handles one clip at a time (batch 1, core/inference.py:26)")
        if device is not None and torch.device(device).type == "cuda":
            dev = torch.device(device)
        elif frames.is_cuda:
            dev = frames.device
        else:
            dev = torch.device("cuda", torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        _, n_frames, _, H, W = frames.shape
        K = masks.shape[2]
        out_mode = self.output
        if out_mode == "reference":                                                 # :388-392
            out_mode = "device" if (torch.cuda.device_count() > 1 and device is None) else "host"
        with torch.cuda.device(dev):
            return self._run(frames, masks, optical_flows, n_objects, memorize_every, dev, n_frames, K, H, W, out_mode,
                             teacher_masks, keep_logits, keep_bboxes, time_frames or self.record_frame_times)

    def _run(self, frames, masks, optical_flows, n_objects, memorize_every, dev, n_frames, K, H, W, out_mode, teacher_masks,
             keep_logits, keep_bboxes, time_frames):
        n_obj_host = n_objects.cpu()
        n = int(n_obj_host.max().item())                                            # n_max_objects (:398)
        if n < 1:
            raise RuntimeError("RMNet.forward needs at least one object (the reference's torch.cat of an empty batch fails too)")
        _, new_at, commit_at, n_commits = memorize_schedule(n_frames, memorize_every, n_obj_host[0].tolist())   # :405-408
        st = self._state(n, K, H, W, n_commits, dev)
        main = torch.cuda.current_stream(dev)

        # inputs: whole-clip frames / flows are moved once (the reference moves a frame at a time, :413, :428-429); of the
        # int32 `masks` only frame 0 and the frames that introduce objects are ever read (:396, :399-402, :436-442)
        frames_d = frames if frames.is_cuda else frames.to(dev, non_blocking=True)
        flows_d = optical_flows if optical_flows.is_cuda else optical_flows.to(dev, non_blocking=True)
        mask0 = masks[:, 0].to(dev)
        existing = torch.unique(torch.argmax(mask0[0], dim=0)).cpu().tolist()       # :399-402
        first = mask0.float()                                                       # :396

        if out_mode == "host":
            est_host = torch.empty((1, n_frames, K, H, W), dtype=torch.float32, pin_memory=True)
            est_host[:, 0].copy_(first, non_blocking=True)
            copy_stream = torch.cuda.Stream(dev)
            stage = [torch.empty((1, K, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
            stage_free = [None, None]
            est_masks = None
        else:
            est_masks = torch.zeros((1, n_frames, K, H, W), dtype=torch.float32, device=dev)   # :387 (kept on the device)
            est_masks[:, 0] = first
        self.last_bboxes = [] if keep_bboxes else None
        self.last_logits = [] if keep_logits else None
        events = [] if time_frames else None

        st.prev_mask.copy_(first)
        for t in range(1, n_frames):
            if time_frames:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record(main)
            if teacher_masks is not None:
                st.prev_mask.copy_(teacher_masks[:, t - 1], non_blocking=True)
            st.prev_frame.copy_(frames_d[:, t - 1], non_blocking=True)               # :413
            st.cur_frame.copy_(frames_d[:, t], non_blocking=True)                    # :428
            st.flow.copy_(flows_d[:, t], non_blocking=True)                          # :429
            commit = commit_at[t]                                                   # :424
            labels = new_mask = None
            if t in new_at:                                                         # :436-438
                mt = masks[0, t].to(dev)
                labels = torch.unique(torch.argmax(mt, dim=0)).cpu().tolist()
                new_mask = mt.to(torch.int32).contiguous()
            modes = channel_modes(K, n, existing, labels)                           # :439-448
            if self.use_graph and CH_NEW not in modes:
                logit, est, prev_bbox, curr_bbox = self._graph_frame(st, commit, modes, keep_logits)
            else:
                logit, est, prev_bbox, curr_bbox = self._frame_body(st, commit, modes, new_mask, keep_logits)
            if keep_bboxes:
                self.last_bboxes.append((prev_bbox.clone(), curr_bbox.clone()))
            if keep_logits:
                self.last_logits.append(logit.clone())
            st.prev_mask.copy_(est)                                                 # :450 -> :412 of the next frame
            if out_mode == "host":
                b = t & 1
                if stage_free[b] is not None:
                    main.wait_event(stage_free[b])
                stage[b].copy_(est)
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    est_host[:, t].copy_(stage[b], non_blocking=True)
                    stage_free[b] = torch.cuda.Event()
                    stage_free[b].record(copy_stream)
            else:
                est_masks[:, t] = est
            if time_frames:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record(main)
                events.append((e0, e1))
        if out_mode == "host":
            copy_stream.synchronize()
            main.synchronize()
            est_masks = est_host
        if time_frames:
            torch.cuda.synchronize(dev)
            self.last_frame_ms = [a.elapsed_time(b) for a, b in events]
        return est_masks

    __call__ = forward
