"""GPU-resident mirror of RMNet.forward's frame loop (models/rmnet.py:385-452) for one clip (batch 1, as
core/inference.py:26 runs it), built on the fused per-frame calls of this library.

The reference keeps `est_masks` on the host when one GPU is visible and moves 18 MB each way per frame (:388-392, :412,
:450), rebuilds the warp grid on the CPU (:257-262), re-copies the whole memory bank every frame (:416-426) and runs
about fifty small ATen kernels around its three conv nets.  Here a frame is

    memorize_net  ->  query_net  ->  RegionalMemory.step (ONE call: 4 chained kernels)  ->  decoder_net  ->  mask_epilogue (1 kernel)

with everything resident on the device.  The three conv nets stay the reference's (out of scope) and are
passed in as callables:

    memorize_net(frame_p [1,3,Hp,Wp], obj_masks [n,Hp,Wp], other_masks [n,Hp,Wp]) -> (k4 [n,128,h,w], v4 [n,512,h,w])
        = kv_memory(encoder_memory(f, m, o))                                  models/rmnet.py:219-236
    query_net(frame_p [1,3,Hp,Wp]) -> (k4q [1,128,h,w], v4q [1,512,h,w], ctx)
        = kv_query(encoder_query(frame)), ctx = (r3, r2)                      :311-315
    decoder_net(m4 [n,1024,h,w], ctx) -> logits [n,2,Hp,Wp]
        = decoder(m4, r3e, r2e)                                               :366

With the reference model `net` (an RMNet instance) these are, e.g.,
    memorize_net = lambda f, m, o: net.kv_memory(net.encoder_memory(f.expand(m.shape[0], -1, -1, -1), m, o)[0])
    query_net    = lambda f: (lambda r4, r3, r2, *_: (*net.kv_query(r4), (r3, r2)))(*net.encoder_query(f))
    decoder_net  = lambda m4, ctx: net.decoder(m4, ctx[0].expand(m4.shape[0], -1, -1, -1), ctx[1].expand(m4.shape[0], -1, -1, -1))
"""
import torch
import torch.nn.functional as F

from . import ops
from ._lib import CH_ABSENT, CH_KEEP, CH_NEW, ELEM_BF16, RMNET_IMPL_AUTO, RMNET_PREC_SPLIT3
from .modules import RegionalMemory


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


class RegionalFrameLoop:
    """forward(frames, masks, optical_flows, n_objects, memorize_every) -> est_masks [1,F,K,H,W], the signature and the
    semantics of RMNet.forward (models/rmnet.py:385) for batch 1; est_masks stays on the device."""

    def __init__(self, memorize_net, query_net, decoder_net, precision=RMNET_PREC_SPLIT3, impl=RMNET_IMPL_AUTO,
                 elem_format=ELEM_BF16):
        self.memorize_net, self.query_net, self.decoder_net = memorize_net, query_net, decoder_net
        self.precision, self.impl, self.elem_format = precision, impl, elem_format
        self.last_bboxes = None   # [(prev_bbox, curr_bbox)] of the last clip, for inspection / tests

    @torch.no_grad()
    def forward(self, frames, masks, optical_flows, n_objects, memorize_every, device=None):
        if frames.shape[0] != 1:
            raise RuntimeError("RegionalFrameLoop handles one clip at a time (batch 1, core/inference.py:26)")
        dev = torch.device(device) if device is not None else (frames.device if frames.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        _, n_frames, _, H, W = frames.shape
        K = masks.shape[2]
        frames, masks, optical_flows = frames.to(dev), masks.to(dev), optical_flows.to(dev)
        n_obj_host = n_objects.cpu()
        n = int(n_obj_host.max().item())                                            # n_max_objects (:398)
        lw, uw, lh, uh = ops.pad_amounts(H, W)
        pad = (lw, uw, lh, uh)
        est_masks = torch.zeros((1, n_frames, K, H, W), dtype=torch.float32, device=dev)   # :387 (kept on the device)
        est_masks[:, 0] = masks[:, 0]                                               # :396
        existing = torch.unique(torch.argmax(masks[0, 0], dim=0)).cpu().tolist()    # :399-402
        _, new_at, commit_at, n_commits = memorize_schedule(n_frames, memorize_every, n_obj_host[0].tolist())   # :405-408
        rm = RegionalMemory(n, (H, W), max_frames=n_commits + 1, device=dev, precision=self.precision, impl=self.impl,
                            elem_format=self.elem_format)
        self.last_bboxes = []
        for t in range(1, n_frames):
            prev_mask = est_masks[:, t - 1]                                         # :412 (already on the device)
            masks_p = F.pad(prev_mask, pad)                                         # :212
            frame_p = F.pad(frames[:, t - 1], pad)
            m, o = object_batches(masks_p, n)                                       # :219-229
            k4, v4 = self.memorize_net(frame_p, m, o)                               # :234-236
            k4q, v4q, ctx = self.query_net(F.pad(frames[:, t], pad))                # :307-315
            commit = commit_at[t]                                                   # :424
            m4, prev_bbox, curr_bbox = rm.step(k4.contiguous(), v4.contiguous(), prev_mask.contiguous(),
                                               optical_flows[:, t].contiguous(), k4q[0].contiguous(), v4q[0].contiguous(),
                                               commit=commit)                       # :239-248, :416-426, :431, :355-361
            self.last_bboxes.append((prev_bbox.clone(), curr_bbox.clone()))
            logits = self.decoder_net(m4, ctx)                                      # :366
            labels = torch.unique(torch.argmax(masks[0, t], dim=0)).cpu().tolist() if t in new_at else None   # :436-438
            modes = channel_modes(K, n, existing, labels)                           # :439-448
            new_mask = masks[0, t].to(torch.int32).contiguous() if t in new_at else None
            _, est = ops.mask_epilogue(logits.contiguous(), K, (H, W), modes, new_mask, want_logit=False)   # :368-380, :289-302, :450
            est_masks[:, t] = est
        return est_masks

    __call__ = forward
