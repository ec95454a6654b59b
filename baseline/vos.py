"""baseline.vos -- BENCH / TEST INFRASTRUCTURE: the VOS frames/sec leg (BASELINE.json `metric`).

Runs exactly the two calls `utils.helpers.multi_scale_inference` makes per clip at scale 1.0 (utils/helpers.py:55-56)

    _est_flows = tflownet(_frames)
    _est_probs = rmnet(_frames, _masks, _est_flows, n_objects, cfg.TEST.MEMORIZE_EVERY)

on host tensors, with the models wrapped like core/inference.py:35-37 wraps them (DataParallel(...).cuda() when CUDA is
there), either on the UNMODIFIED reference (`impl="reference"`, GPU: + its unmodified CUDA extension from oracle/_ref;
CPU: + the C-oracle generator stand-in, since that extension refuses CPU tensors) or after `rmnet_b200.install()`
(`impl="ours"`).  FPS = segmented frames (F - 1) / wall time of the two calls, bracketed by device synchronisation; data
loading, the final bilinear resize of est_probs (utils/helpers.py:60-62) and PNG writing are excluded.
"""
import os
import time

import baseline

WORKLOADS = {
    # name: (H, W, objects).  Memory length T is a function of the clip length at memorize_every = 5 (config.py:139):
    # T(t) = 1 + #{j in {0,5,10,...} : j < t-1}  ->  T = 20 from frame 92 on
    "c1": (240, 432, 1),
    "c2": (480, 854, 3),
    "c3": (480, 854, 5),
    "c4": (720, 1280, 10),
}


def reference_flags():
    """runner.py:73-74 (deterministic cuDNN, no autotuning); TF32 convs stay at torch's default (allowed), as the reference runs."""
    import torch
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def wrap(tfn, net, device_ids=None):
    """core/inference.py:35-37."""
    import torch
    if torch.cuda.is_available() and next(net.parameters()).is_cuda:
        ids = device_ids if device_ids is not None else [torch.cuda.current_device()]
        return torch.nn.DataParallel(tfn, device_ids=ids), torch.nn.DataParallel(net, device_ids=ids)
    return tfn, net


def run_clip(tfn, net, frames, masks, n_objects, every=5):
    """utils/helpers.py:55-56 on host tensors -> (est_probs, seconds, seconds of the tflownet call alone)."""
    import torch
    cuda = torch.cuda.is_available() and next(net.parameters()).is_cuda
    with torch.no_grad():
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        flows = tfn(frames)
        if cuda:
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        probs = net(frames, masks, flows, n_objects, every)
        if cuda:
            torch.cuda.synchronize()
        t2 = time.perf_counter()
    return probs, t2 - t0, t1 - t0


def module_split(net, tfn, H, W, n, T, dev, reps=5):
    """Per-module device time of ONE frame at (H, W, n objects, T memory frames): the reference's cuDNN nets (called as
    the frame loop calls them) next to this library's fused step and mask epilogue -> dict of ms."""
    import torch
    import torch.nn.functional as F

    import rmnet_b200
    from rmnet_b200 import ops
    from rmnet_b200.frame_loop import RegionalFrameLoop, object_batches
    loop = RegionalFrameLoop.from_rmnet(net)
    lw, uw, lh, uh = ops.pad_amounts(H, W)
    pad = (lw, uw, lh, uh)
    K = baseline.K_TEST
    g = torch.Generator(device="cpu").manual_seed(3)
    frame = torch.randn((1, 3, H, W), generator=g).to(dev)
    frames, masks, _ = baseline.synthetic_clip(11, n, 2, H, W)
    prev_mask = masks[:, 0].float().to(dev)
    flow = (torch.randn((1, 2, H, W), generator=g) * 2).to(dev)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev)
    out = {}

    def timed(name, fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out[name] = sorted(ts)[len(ts) // 2]
        return r

    with torch.no_grad():
        masks_p = F.pad(prev_mask, pad)
        frame_p = F.pad(frame, pad)
        m, o = timed("object_batches (torch glue, :219-229)", lambda: object_batches(masks_p, n))
        k4, v4 = timed("encoder_memory + kv_memory (cuDNN, n objects)", lambda: loop.memorize_net(frame_p, m, o))
        k4q, v4q, ctx = timed("encoder_query + kv_query (cuDNN)", lambda: loop.query_net(frame_p))
        for t in range(T - 1):
            rm.memorize(k4.contiguous(), v4.contiguous(), prev_mask, commit=True)
        m4, _, _ = timed("rmnet_b200 frame step (regions + pack + read + merge)",
                         lambda: rm.step(k4.contiguous(), v4.contiguous(), prev_mask, flow, k4q[0].contiguous(), v4q[0].contiguous(), commit=False))
        logits = timed("decoder (cuDNN, n objects)", lambda: loop.decoder_net(m4, ctx))
        timed("rmnet_b200 mask epilogue", lambda: ops.mask_epilogue(logits.contiguous(), K, (H, W), None, None, want_logit=False))
        if tfn is not None:
            timed("TinyFlowNet._forward (cuDNN)", lambda: tfn._forward(frame, frame))
    return out


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def set_host_threads():
    """All host cores for the CPU arm (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    import torch
    n = os.cpu_count() or 1
    try:
        torch.set_num_threads(n)
    except RuntimeError:
        pass
    return n
