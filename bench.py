#!/usr/bin/env python
"""bench.py -- RMNet's per-frame regional memory-read hot path on B200, and the VOS frames/sec it yields.

`value` (device-timed, inputs resident in HBM): a "step" is ONE frame of one clip through the hot path (SURVEY 8a rows
a2-a10) in the steady state of a clip -- one call of RegionalMemory.step = rmnet_frame_step = four kernels chained by
programmatic dependent launch:

    regions : one pass over prev_mask [1,11,H,W] + flow -> the box of the zero-padded mask (memorise side,
              models/rmnet.py:212 + :244) AND the box of the flow-warped mask (segment side, :431), with their /16 cell
              rectangles (:245, :307 + :356)
    pack    : k4/v4 of the previous frame into the memory bank as its temporary last frame (:239-248, :416-426) +
              the query side: k4q * att16 packed for the tensor cores, v4q * att16 into mem_val[:, 512:] (:355-358, :163)
    read    : tcgen05 split-KV attention of all objects against the region-compacted bank (:147-165)
    merge   : split combination + masked-cell correction + uniform rows -> mem_val [n,1024,h,w]

`e2e` (the headline): 480p VOS frames/sec through the reference-facing API -- the two calls utils/helpers.py:55-56 makes
per clip, `tflownet(frames)` and `rmnet(frames, masks, flows, n_objects, 5)`, on HOST tensors, with the UNMODIFIED
reference nets (ResNet-50 encoders, decoder, TinyFlowNet: cuDNN, out of scope) wrapped like core/inference.py:35-37 and
`rmnet_b200.install()` deployed behind them (RMNet.forward = the fused GPU-resident frame loop).  Workload: 8 clips per
GPU shaped like BASELINE configs[2]/[4] (480x854, 5 objects, F = 60, memorize_every = 5), sharded clip-parallel over
the ranks longest-first, one NCCL gather of the uint8 label maps (core/inference.py:61) at the end.  On rank 0 at N = 1
the `vos` object adds: one F = 100 clip (memory reaches T = 20) with the frames/s of its T >= 20 tail, the unmodified
reference (its own forward + its own CUDA extension) on the SAME GPU and clip, and the per-module device-time split.
`e2e_op` is round 1's op-level end-to-end number (host k4/v4/q tensors in, mem_val out per step).

Extra legs on rank 0 at N = 1: the step replayed as a CUDA graph, the reference's own composition of the step on the same
GPU (torch CUDA ops + its unmodified CUDA kernel), the attention kernel alone (roofline), the CPU baseline.

`--impl reference`: the reference on the host cores -- `value` = the same hot-path step by the reference's own functions
(MemoryReader / RMNet.warp / RMNet.get_att_map from baseline/_ref on torch CPU; the CUDA-only generator through the C
oracle), `e2e` = VOS frames/sec of the unmodified tflownet + rmnet on a bounded sample clip.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl ours|reference] [--precision split3|single]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[1]: single 480x864 clip, 3 objects, T=5 memory
    "c2": dict(H=480, W=854, n=3, T=5, desc="480x854 (padded 480x864) clip frame, 3 objects, T=5 memory frames, K=11 mask channels"),
    # BASELINE.json configs[2] / north_star target shape: 480p, 5 objects, T=20
    "c3": dict(H=480, W=854, n=5, T=20, desc="480x854 (padded 480x864) clip frame, 5 objects, T=20 memory frames, K=11 mask channels"),
    # BASELINE.json configs[3]: YouTube-VOS-shaped, 720p, 10 objects, T=40 (long memory); slow to set up, CPU sample = 1 step
    "c4": dict(H=720, W=1280, n=10, T=40, desc="720x1280 clip frame, 10 objects, T=40 memory frames, K=11 mask channels"),
}
K_CH = 11
METRIC = "480p VOS frames/sec (regional memory-read hot path)"


# ------------------------------------------------------------------------------------------------------------
# synthetic clip state: drifting rectangles, soft masks, N(0, 2 px) flow, N(0,1) key/value features
# ------------------------------------------------------------------------------------------------------------
def make_pool(wl, seed, pool):
    import synth
    H, W, n, T = wl["H"], wl["W"], wl["n"], wl["T"]
    rng = np.random.default_rng(seed)
    Hp, Wp = (H + 15) // 16 * 16, (W + 15) // 16 * 16
    lw = (Wp - W) // 2
    h, w = Hp // 16, Wp // 16
    # object rectangles drifting a few pixels per frame
    boxes = []
    for _ in range(n):
        bh, bw = int(rng.uniform(0.15, 0.45) * H), int(rng.uniform(0.15, 0.45) * W)
        boxes.append([int(rng.integers(0, H - bh)), int(rng.integers(0, W - bw)), bh, bw, int(rng.integers(-3, 4)), int(rng.integers(-3, 4))])

    def frame_mask(step):
        lab = np.zeros((H, W), np.int64)
        for o, (y0, x0, bh, bw, vy, vx) in enumerate(boxes, 1):
            y = int(np.clip(y0 + vy * step, 0, H - bh))
            x = int(np.clip(x0 + vx * step, 0, W - bw))
            lab[y:y + bh, x:x + bw] = o
        return synth.soft_masks(rng, lab, K_CH)

    frames = []
    for i in range(T + pool):
        frames.append(dict(
            mask=frame_mask(i),                                                       # [K,H,W] soft probabilities
            flow=synth.flow_field(rng, H, W, 2.0),                                    # [2,H,W]
            k4=(rng.standard_normal((n, 128, h, w)) * 0.5).astype(np.float32),        # kv_memory(prev frame)
            v4=rng.standard_normal((n, 512, h, w)).astype(np.float32),
            qk=(rng.standard_normal((128, h, w)) * 0.5).astype(np.float32),           # kv_query(current frame)
            qv=rng.standard_normal((512, h, w)).astype(np.float32)))
    return dict(frames=frames, Hp=Hp, Wp=Wp, lw=lw, h=h, w=w)


def pad_mask(mask, lw, Wp):
    out = np.zeros(mask.shape[:-1] + (Wp,), np.float32)
    out[..., lw:lw + mask.shape[-1]] = mask
    return out


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same step on the host cores (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------------
class CpuClip:
    """Reference algorithm on the CPU (oracle/: C restatement of the generator / warp, numpy-BLAS MemoryReader),
    composed exactly as models/rmnet.py:244-248, :431, :307, :355-361 compose it (mask, then the DENSE reader)."""

    def __init__(self, wl, pool):
        import oracle
        self.o, self.wl, self.pool = oracle, wl, pool
        n, T = wl["n"], wl["T"]
        self.keys, self.vals = [], []   # committed memory frames: masked [n,128,h,w], [n,512,h,w]
        for t in range(T - 1):
            k, v = self._masked_memory(pool["frames"][t])
            self.keys.append(k)
            self.vals.append(v)

    def _masked_memory(self, fr):
        n = self.wl["n"]
        mp = pad_mask(fr["mask"], self.pool["lw"], self.pool["Wp"])
        att, _ = self.o.reg_att_map(mp[None])                                       # :244
        a16 = self.o.downsample16(att[0, 1:n + 1])                                  # :245
        return fr["k4"] * a16[:, None], fr["v4"] * a16[:, None]                     # :247-248

    def step(self, i):
        n, T = self.wl["n"], self.wl["T"]
        fr = self.pool["frames"][T - 1 + i % (len(self.pool["frames"]) - T + 1)]
        k_t, v_t = self._masked_memory(fr)                                          # memorise (temporary frame)
        m_key = np.stack(self.keys + [k_t], 2)
        m_val = np.stack(self.vals + [v_t], 2)
        att, _ = self.o.get_att_map(fr["mask"][None], fr["flow"][None], arith="cuda")   # :431
        attp, _ = self.o.pad_divide_by(att[0, 1:n + 1])                             # :307
        a16 = self.o.downsample16(attp)                                             # :356
        qk = np.broadcast_to(fr["qk"], (n,) + fr["qk"].shape) * a16[:, None]        # :357
        qv = np.broadcast_to(fr["qv"], (n,) + fr["qv"].shape) * a16[:, None]        # :358
        mem_val, _ = self.o.memory_read(m_key, m_val, qk.astype(np.float32), qv.astype(np.float32))   # :361
        return mem_val


def time_cpu(wl, pool, steps, warmup):
    try:   # all host threads for the BLAS-backed reader (torchrun exports OMP_NUM_THREADS=1 to its workers)
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    clip = CpuClip(wl, pool)
    for i in range(warmup):
        clip.step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        out = clip.step(i)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, out


class RefCpuClip:
    """The same step by the REFERENCE's own functions on torch's CPU backend (imported from baseline/_ref, the shipped
    copy of the reference tree): models/rmnet.py MemoryReader.forward (:147-165), RMNet.warp (:252-278), RMNet.get_att_map
    (:280-287), utils.helpers.pad_divide_by; the three inline lines of memorize / segment that connect them (:245-248,
    :356-358) restated; the CUDA-only generator extension replaced by the C oracle (baseline.cpu_generator_class)."""

    def __init__(self, wl, pool):
        import types
        import torch
        import baseline
        self.torch, self.wl, self.pool = torch, wl, pool
        ref = baseline.import_reference()
        import utils.helpers as ref_helpers
        self.helpers = ref_helpers
        self.reader = ref.MemoryReader()
        ns = types.SimpleNamespace(att_map_generator=baseline.cpu_generator_class()())
        ns.warp = lambda img0, flow: ref.RMNet.warp(ns, img0, flow)
        self.get_att_map = lambda prev_mask, flow=None: ref.RMNet.get_att_map(ns, prev_mask, flow)
        n, T = wl["n"], wl["T"]
        self.keys, self.vals = [], []
        for t in range(T - 1):
            k, v = self._masked_memory(pool["frames"][t])
            self.keys.append(k)
            self.vals.append(v)

    def _t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a))

    def _masked_memory(self, fr):
        torch, n, H, W = self.torch, self.wl["n"], self.wl["H"], self.wl["W"]
        (masks,), _ = self.helpers.pad_divide_by([self._t(fr["mask"])[None]], 16, (H, W))          # :212
        att, _ = self.get_att_map(masks)                                                          # :244
        a16 = torch.nn.functional.interpolate(att, scale_factor=1 / 16)[0, 1:n + 1, None]         # :245
        return self._t(fr["k4"]) * a16, self._t(fr["v4"]) * a16                                   # :247-248

    def step(self, i):
        torch, n, T, H, W = self.torch, self.wl["n"], self.wl["T"], self.wl["H"], self.wl["W"]
        fr = self.pool["frames"][T - 1 + i % (len(self.pool["frames"]) - T + 1)]
        k_t, v_t = self._masked_memory(fr)
        m_key = torch.stack(self.keys + [k_t], 2)                                                 # :416-421
        m_val = torch.stack(self.vals + [v_t], 2)
        att, _ = self.get_att_map(self._t(fr["mask"])[None], self._t(fr["flow"])[None])           # :431
        (att,), _ = self.helpers.pad_divide_by([att], 16, (H, W))                                 # :307
        a16 = torch.nn.functional.interpolate(att[0, 1:n + 1, None], scale_factor=1 / 16)         # :330, :356
        k4e = self._t(fr["qk"])[None].expand(n, -1, -1, -1) * a16                                 # :357
        v4e = self._t(fr["qv"])[None].expand(n, -1, -1, -1) * a16                                 # :358
        mem_val, _ = self.reader(m_key.contiguous(), m_val.contiguous(), k4e, v4e)                # :361
        return mem_val.numpy()


def time_cpu_arm(wl, pool, steps, warmup, budget_s=90.0):
    """-> (frames/s, ms/step, steps actually timed, kind, description).  kind "reference" = RefCpuClip (the reference's
    own functions from baseline/_ref on torch CPU); "port" = the numpy/C oracle when the reference tree is not there."""
    import baseline
    cores = os.cpu_count() or 1
    if baseline.available():
        import torch
        from baseline import vos
        vos.set_host_threads()
        with torch.no_grad():
            clip = RefCpuClip(wl, pool)
            for i in range(warmup):
                clip.step(i)
            t0 = time.perf_counter()
            done = 0
            for i in range(steps):
                clip.step(i)
                done += 1
                if time.perf_counter() - t0 > budget_s:
                    break
            dt = time.perf_counter() - t0
        return done / dt, dt / done * 1e3, done, "reference", (
            f"{done} steps of the same workload by the reference's own MemoryReader / RMNet.warp / RMNet.get_att_map (baseline/_ref, "
            f"torch {torch.__version__} CPU, {torch.get_num_threads()} threads; generator = C oracle, the reference's is CUDA-only)")
    fps, ms, _ = time_cpu(wl, pool, steps, warmup)
    return fps, ms, steps, "port", f"{steps} steps of the same workload (oracle/: C generator + warp, numpy-BLAS MemoryReader on {cores} threads)"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  In-process NVML
    (nvidia_ml_py) polled from a thread: an `nvidia-smi -lms` child stalled kernel launches for milliseconds whenever one of
    its queries landed inside the timed loop (one 3.4 ms step in 50), which a direct NVML read does not.  Falls back to
    one nvidia-smi snapshot when NVML cannot be imported."""
    NAMES = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
             "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
    OLD = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
           "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}

    def __init__(self, index, period_s=0.005):
        self.rows, self.nv, self.h, self.stop_flag, self.period = [], None, None, False, period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.bits = {k: getattr(pynvml, v, getattr(pynvml, self.OLD[k], 0)) for k, v in self.NAMES.items()}
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self, t0, t1):
        if self.nv is None:
            return self._smi_snapshot()
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        sm = [r[1] for r in rows]
        reasons = sorted({k for r in rows for k, b in self.bits.items() if b and (r[2] & b)})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(rows),
                "source": "nvml"}

    @staticmethod
    def _smi_snapshot():
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            r = [c.strip() for c in out.strip().splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(r[0]), "sm_max_mhz": float(r[1]), "reasons": [n for n, v in zip(names, r[2:]) if v.lower().startswith("active")],
                    "samples": 1, "source": "nvidia-smi snapshot after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def shard_clips(n_clips, rank, world):
    """Clip-parallel sharding: clip i runs on rank i % world; no data-path collective."""
    return [i for i in range(n_clips) if i % world == rank]


def shard_longest_first(costs, world):
    """Clip-parallel sharding, longest first: clips sorted by cost (F * n), each given to the least loaded
    rank so far (ties -> lowest rank).  -> list of `world` lists of clip indices."""
    shards, load = [[] for _ in range(world)], [0.0] * world
    for i in sorted(range(len(costs)), key=lambda j: (-costs[j], j)):
        r = min(range(world), key=lambda q: (load[q], q))
        shards[r].append(i)
        load[r] += costs[i]
    return shards


def gather_label_maps(stack, rank, world):
    """One gather of every rank's uint8 label maps [clips, F, H, W] to rank 0 (core/inference.py:61 keeps exactly these):
    NCCL on the GPU box, gloo in the CPU tests.  -> list of `world` tensors on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [stack]
    outs = [torch.empty_like(stack) for _ in range(world)] if rank == 0 else None
    dist.gather(stack, outs, dst=0)
    return outs


def pin_rank_to_cores(local_rank, local_world):
    """Give each rank its own slice of the host cores (the ranks of one node otherwise migrate over all of them and
    contend for the same caches while staging pageable copies).  Returns the slice, or None when not applicable."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if local_world <= 1 or len(cores) < 2 * local_world:
            return None
        per = len(cores) // local_world
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return None


VOS_CLIPS_PER_GPU = 8
VOS_SHAPE = dict(H=480, W=854, n=5, F=60, every=5)      # BASELINE configs[2]/[4] frame shape, the target's 5 objects
VOS_LONG = dict(H=480, W=854, n=5, F=100, every=5)      # memory reaches T = 20 at frame 92


def run_vos(args, rank, world, local_rank, dev, precision):
    """The e2e leg: VOS frames/sec through the reference-facing API (see the module docstring).
    -> (seconds of this rank's timed calls [+ gather], frames segmented by this rank, info dict)."""
    import torch
    import baseline
    from baseline import vos
    import rmnet_b200
    ref = baseline.import_reference()
    vos.reference_flags()
    tfn, net = baseline.build_nets(0, dev, cpu_generator=False)
    tfn_dp, net_dp = vos.wrap(tfn, net, [local_rank])                                 # core/inference.py:35-37
    # output="host": est_masks comes back as a (pinned) host tensor at every N -- the reference's own rule would keep it on
    # the device as soon as more than one GPU is VISIBLE (models/rmnet.py:388-392), which would make the per-rank work of
    # the N = 1 and N > 1 runs differ
    rmnet_b200.install(ref, precision=precision, output="host")
    H, W, n, F_, every = (VOS_SHAPE[k] for k in ("H", "W", "n", "F", "every"))
    n_clips = VOS_CLIPS_PER_GPU * world
    clips = [dict(seed=5000 + i, n=n, F=F_) for i in range(n_clips)]
    if args.e2e_set == "davis30":
        # BASELINE configs[2] as SURVEY 8d spells it out: 30 clips with the DAVIS-2017 val clip lengths (34..104 frames, 1 999 in
        # total; read from the reference's own datasets/DAVIS.json), 480x854, object counts cycling 1..5, sharded longest-first
        lengths = [v["n_frames"] for v in json.load(open(os.path.join(baseline.ref_root(), "datasets", "DAVIS.json")))["val"]]
        clips = [dict(seed=6000 + i, n=1 + i % 5, F=f) for i, f in enumerate(lengths)]
        n_clips = len(clips)
    mine = shard_longest_first([c["F"] * c["n"] for c in clips], world)[rank]
    L = rmnet_b200.lib()

    def host_clip(c):
        frames, masks, n_objects = baseline.synthetic_clip(c["seed"], c["n"], c["F"], H, W)
        return frames, masks, n_objects                                               # pageable host tensors, int32 masks (utils/helpers.py:52-53)

    def labels_of(probs):
        return probs[0].to(dev).argmax(1).to(torch.uint8)                             # core/inference.py:61 (on the device: 270 M floats)

    # warm-up: one short clip of the same shape (cuDNN plans, graph capture of the frame body)
    for n_w in sorted({clips[ci]["n"] for ci in mine}):
        wf, wm, wn = baseline.synthetic_clip(4999, n_w, 8, H, W)
        vos.run_clip(tfn_dp, net_dp, wf, wm, wn, every)
    loop = net.__dict__["_rmnet_b200_loop"]
    L.rmnet_launch_count_reset()
    g0 = loop.graph_launches
    secs, frames_done, labs, flow_s = 0.0, 0, [], 0.0
    h2d = d2h = 0
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    for ci in mine:
        frames, masks, n_objects = host_clip(clips[ci])
        probs, s, sf = vos.run_clip(tfn_dp, net_dp, frames, masks, n_objects, every)   # timed: the two calls, est_masks back on the host
        assert not probs.is_cuda
        secs += s
        flow_s += sf
        frames_done += c_frames(clips[ci])
        labs.append(labels_of(probs))                                                 # untimed: what core/inference.py:61 keeps of a clip
        # bytes crossing PCIe, counted from the tensors: DataParallel scatters every tensor argument of both calls (frames
        # twice, the int32 masks, the flows when TinyFlowNet's rule left them on the host); est_masks and those flows go back
        flows_on_host = not (torch.cuda.device_count() > 1)                           # models/tiny_flownet.py:124-127
        flow_bytes = frames.shape[1] * 2 * H * W * 4
        h2d += 2 * frames.numel() * 4 + masks.numel() * 4 + (flow_bytes if flows_on_host else 0)
        d2h += (flow_bytes if flows_on_host else 0) + probs.numel() * 4
        del frames, masks, probs
    launches = int(L.rmnet_launch_count()) + loop.graph_launches - g0
    gather_s = 0.0
    gathered = None
    if world > 1:
        import torch.distributed as dist
        stack = torch.cat(labs)                                                       # [frames of this rank's clips, H, W] uint8
        nf = torch.tensor([stack.shape[0]], device=dev, dtype=torch.int64)
        dist.all_reduce(nf, op=dist.ReduceOp.MAX)                                     # ranks may hold different frame totals (davis30)
        if int(nf) > stack.shape[0]:
            stack = torch.cat([stack, stack.new_zeros((int(nf) - stack.shape[0],) + tuple(stack.shape[1:]))])
        gather_label_maps(stack[:1].contiguous(), rank, world)                        # untimed: NCCL sets its channels up on first use
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = gather_label_maps(stack, rank, world)                                  # the one collective of the job
        torch.cuda.synchronize()
        gather_s = time.perf_counter() - t0
        secs += gather_s
        if rank == 0:
            gathered = [int(o.numel()) for o in outs]
    info = {"clips_total": n_clips, "clips_this_rank": len(mine), "frames_per_clip": F_ if args.e2e_set == "default" else "DAVIS-2017 val lengths (34..104)",
            "objects": n if args.e2e_set == "default" else "1..5 cycling", "memorize_every": every, "set": args.e2e_set,
            "flownet_share": flow_s / max(secs, 1e-9), "gather_s": gather_s, "gathered_label_bytes": gathered,
            "launches": launches, "h2d_bytes_per_frame": h2d / max(frames_done, 1), "d2h_bytes_per_frame": d2h / max(frames_done, 1),
            "allow_tf32_convs": bool(torch.backends.cudnn.allow_tf32), "graph": bool(loop.use_graph),
            "label_checksum": float(sum(float(l.sum()) for l in labs))}

    extra = None
    if world == 1 and rank == 0 and not args.no_vos_extras:
        extra = vos_extras(ref, tfn, net, tfn_dp, net_dp, dev, precision)
    rmnet_b200.uninstall(ref)
    return secs, frames_done, info, extra


def c_frames(c):
    return c["F"] - 1


def vos_extras(ref, tfn, net, tfn_dp, net_dp, dev, precision):
    """Rank 0 at N = 1: the F = 100 clip whose memory reaches T = 20 (north_star's target shape) -- ours, then the
    UNMODIFIED reference (its own forward, its own CUDA extension from oracle/_ref) on the same GPU and clip -- and the
    per-module device-time split of one frame."""
    import torch
    import baseline
    from baseline import vos
    import rmnet_b200
    H, W, n, F_, every = (VOS_LONG[k] for k in ("H", "W", "n", "F", "every"))
    out = {"clip": f"{H}x{W}, {n} objects, F={F_}, memorize_every={every} (T reaches 20 at frame 92)", "target_fps": 30.0}
    frames, masks, n_objects = baseline.synthetic_clip(7000, n, F_, H, W)
    loop = net.__dict__["_rmnet_b200_loop"]
    loop.record_frame_times = True
    try:
        vos.run_clip(tfn_dp, net_dp, frames[:, :8], masks[:, :8], n_objects[:, :8], every)
        probs, s, sf = vos.run_clip(tfn_dp, net_dp, frames, masks, n_objects, every)
    finally:
        loop.record_frame_times = False
    ms = loop.last_frame_ms
    tail = ms[91:]                                                                    # frames t >= 92: T >= 20
    out["ours"] = {"fps": (F_ - 1) / s, "seconds": s, "flownet_seconds": sf,
                   "rmnet_frame_ms_median": float(np.median(ms)), "rmnet_frame_ms_at_T20": float(np.mean(tail)),
                   "rmnet_only_fps_at_T20": 1e3 / float(np.mean(tail)), "meets_target": (F_ - 1) / s >= 30.0}
    lab = probs[0].argmax(1).cpu()
    del probs
    try:
        baseline.import_reference(need_cuda_extension=True)
        import reg_att_map_generator as ext
        if "oracle" not in (getattr(ext, "__file__", "") or ""):
            raise RuntimeError("the reference CUDA extension (oracle/_ref) is not the module `reg_att_map_generator` resolves to")
        rmnet_b200.uninstall(ref)
        vos.run_clip(tfn_dp, net_dp, frames[:, :4], masks[:, :4], n_objects[:, :4], every)
        probs_r, s_r, sf_r = vos.run_clip(tfn_dp, net_dp, frames, masks, n_objects, every)
        lab_r = probs_r[0].argmax(1).cpu()
        out["reference_on_this_gpu"] = {"fps": (F_ - 1) / s_r, "seconds": s_r, "flownet_seconds": sf_r,
                                        "what": "unmodified models/rmnet.py + models/tiny_flownet.py (baseline/_ref) + the unmodified reference CUDA extension "
                                                "(oracle/_ref), same GPU, same clip, same two calls",
                                        "label_agreement_free_running": float((lab == lab_r).float().mean())}
        out["speedup_vs_reference_on_this_gpu"] = s_r / s
        del probs_r
    except Exception as e:
        out["reference_on_this_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        rmnet_b200.install(ref, precision=precision, output="host")
    try:
        out["module_split_ms_at_T20"] = vos.module_split(net, tfn, H, W, n, 20, dev)
    except Exception as e:
        out["module_split_ms_at_T20"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def reduce_over_ranks(dev_ms, e2e_s, checksum, device, rank, world):
    """Timing = MAX over ranks (all_reduce), results = one gather of per-rank checksums to rank 0.
    Works with NCCL (device = cuda) and with gloo (device = cpu, used by the CPU tests)."""
    if world == 1:
        return dev_ms, e2e_s, [checksum]
    import torch
    import torch.distributed as dist
    t = torch.tensor([dev_ms, e2e_s], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mine = torch.tensor([checksum], device=device, dtype=torch.float64)
    sums = [torch.zeros(1, device=device, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, sums, dst=0)
    return float(t[0]), float(t[1]), ([float(x) for x in sums] if rank == 0 else None)


def algorithmic_work(counts_m, counts_q, N):
    """SURVEY 8d / BASELINE.md 4: per object  flops = 1280 * M_r * N_r,  bytes = 4*[640*(M_r+N_r) + 1024*N]."""
    flops = sum(1280.0 * m * q for m, q in zip(counts_m, counts_q))
    byts = sum(4.0 * (640.0 * (m + q) + 1024.0 * N) for m, q in zip(counts_m, counts_q))
    return flops, byts


def run_gpu(args, wl, rank, world, local_rank):
    import torch
    import rmnet_b200
    from rmnet_b200 import ops
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    core_slice = pin_rank_to_cores(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))   # before any pinned allocation
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops", 1590.0))          # burst figure: the kernel is timed alone
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    n, T, H, W = wl["n"], wl["T"], wl["H"], wl["W"]
    POOL = 8 if args.workload != "c4" else 3
    pool = make_pool(wl, 1234 + rank, POOL)
    h, w, lw, Wp = pool["h"], pool["w"], pool["lw"], pool["Wp"]
    N = h * w
    frames = pool["frames"]
    precision = {"single": rmnet_b200.RMNET_PREC_SINGLE, "mixed": rmnet_b200.RMNET_PREC_MIXED}.get(args.precision, rmnet_b200.RMNET_PREC_SPLIT3)
    rm = rmnet_b200.RegionalMemory(n, (H, W), max_frames=T, device=dev, precision=precision)

    def to_dev(fr):
        return {k: torch.from_numpy(v).to(dev) for k, v in fr.items()}

    # steady state: T-1 committed memory frames, the T-th is rewritten every step as the temporary frame
    for t in range(T - 1):
        d = to_dev(frames[t])
        rm.memorize(d["k4"], d["v4"], d["mask"][None], commit=True)
    dframes = [to_dev(fr) for fr in frames[T - 1:]]
    hframes = [{k: torch.from_numpy(v) for k, v in fr.items()} for fr in frames[T - 1:]]   # host copies; pinned staging: see e2e
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    out_dev = torch.empty((n, 1024, h, w), dtype=torch.float32, device=dev)   # static result buffer (no allocator calls in the loop)

    def step_dev(d):
        m4, _, _ = rm.step(d["k4"], d["v4"], d["mask"][None], d["flow"][None], d["qk"], d["qv"], commit=False, out=out_dev)
        return m4

    # e2e staging: ONE pinned host buffer per frame holding exactly what the step reads -- mask channels 1..n (k_scan = n+1:
    # the channels of absent objects are never fetched, see rmnet_regional_boxes_forward), flow, k4, v4, q_key, q_val --
    # so that a step's inputs cross PCIe in one copy.  On the device the buffer starts with one spare plane (mask channel
    # 0, never read); the [1,K,H,W] mask tensor handed to the step is a view of the buffer's head (its channels above n
    # alias the other inputs and are, again, never read).
    ORDER = ("mask", "flow", "k4", "v4", "qk", "qv")
    plane = H * W

    def flat_host(fr):
        parts = [fr["mask"][1:n + 1].reshape(-1)] + [fr[k].reshape(-1) for k in ORDER[1:]]
        return torch.cat(parts).pin_memory()

    hflat = [flat_host(fr) for fr in hframes]
    flat_len = plane + hflat[0].numel()
    assert flat_len >= K_CH * plane

    def device_views(buf):
        v, off = {"mask": buf[:K_CH * plane].view(K_CH, H, W)}, plane * (n + 1)
        for k in ORDER[1:]:
            cnt = hframes[0][k].numel()
            v[k] = buf[off:off + cnt].view(hframes[0][k].shape)
            off += cnt
        return v

    h2d = hflat[0].numel() * 4
    d2h = n * 1024 * h * w * 4

    if world > 1:
        import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler starts first so that its NVML initialisation is outside the timed region)
    clocks = ClockSampler(local_rank) if rank == 0 and not os.environ.get("RMNET_BENCH_NO_SMI") else None
    for i in range(max(args.warmup, 3)):
        step_dev(dframes[i % len(dframes)])
    torch.cuda.synchronize()

    # ---- timed: K steps, device time by CUDA events on the launching stream, L2 flushed between steps
    L = rmnet_b200.lib()
    barrier()
    t_begin = time.perf_counter()
    L.rmnet_launch_count_reset()
    evs = []
    for i in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m4 = step_dev(dframes[i % len(dframes)])
        b.record()
        evs.append((a, b))
    barrier()
    launches = int(L.rmnet_launch_count())
    checksum = float(m4.double().sum().item())     # of the last timed step's mem_val: rank 0's equals the N = 1 run's (same pool seed)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = float(np.sum(step_ms))
    if os.environ.get("RMNET_BENCH_DEBUG"):
        med = float(np.median(step_ms))
        print("step outliers (idx:ms):", " ".join(f"{i}:{t:.3f}" for i, t in enumerate(step_ms) if t > 2 * med), file=sys.stderr)

    # ---- the other precision modes of the read kernel on the same bank and inputs (rank 0; the bank's fp16 hi/lo planes serve
    #      all three: strict = 3 products for both GEMMs, mixed = 3 for Q.K^T + 1 for P.V, fast = 1 + 1), device time per step
    modes_info = None
    if rank == 0:
        modes_info = {}
        keep = rm.precision
        for name, prec in (("split3", rmnet_b200.RMNET_PREC_SPLIT3), ("mixed", rmnet_b200.RMNET_PREC_MIXED), ("single", rmnet_b200.RMNET_PREC_SINGLE)):
            rm.precision = prec
            for i in range(3):
                step_dev(dframes[i % len(dframes)])
            mev = []
            for i in range(min(args.steps, 50)):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                step_dev(dframes[i % len(dframes)])
                b.record()
                mev.append((a, b))
            torch.cuda.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in mev]))
            modes_info[name] = {"ms_per_step": ms, "frames_per_s": 1e3 / ms}
        rm.precision = keep
        modes_info["logit_error_under_the_real_decoder"] = (
            "max-abs on RMNet.segment's logit map vs the unmodified model, default-init weights (|score| 642) / conditioned: strict 1.1e-5 / 1.4e-5, "
            "mixed 3.8e-4 / 1.2e-5, fast 3.9e-3 / 3.5e-4 (tests/test_gpu_rmnet.py; north_star's bound: 1e-3)")

    # ---- the same step captured once as a CUDA graph (RegionalMemory.capture_step) and replayed: device time per step and
    #      the host time per step of both submission paths (enqueue only, measured over a batch with one sync at the end)
    graph_info = None
    try:
        g_in = {k: v.clone() for k, v in dframes[0].items()}
        cs = rm.capture_step(g_in["k4"], g_in["v4"], g_in["mask"][None], g_in["flow"][None], g_in["qk"], g_in["qv"], commit=False, out=out_dev)
        for _ in range(3):
            cs.replay()
        torch.cuda.synchronize()
        gev = []
        for i in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            cs.replay()
            b.record()
            gev.append((a, b))
        torch.cuda.synchronize()
        g_ms = [a.elapsed_time(b) for a, b in gev]
        host = {}
        for name, fn in (("eager", lambda: step_dev(dframes[0])), ("graph", cs.replay)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            host[name] = (time.perf_counter() - t0) / args.steps * 1e6
            torch.cuda.synchronize()
        graph_info = {"ms_per_step": float(np.sum(g_ms)) / args.steps, "ms_per_step_median": float(np.median(g_ms)),
                      "host_us_per_step_eager": host["eager"], "host_us_per_step_graph": host["graph"],
                      "note": "static inputs (no per-step input copies); host time = enqueue only"}
    except Exception as e:   # the eager path above is the measured one; a capture failure is reported, not fatal
        graph_info = {"error": f"{type(e).__name__}: {e}"}

    # ---- e2e: same steps through the public API with pinned HOST buffers.  Every step's inputs are copied H2D and its
    # result is read back D2H inside the timed region; copies of step i+1 / i-1 overlap the kernels of step i on two
    # copy streams (double-buffered device inputs and outputs), as a real loader would.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    dflat = [torch.zeros(flat_len, dtype=torch.float32, device=dev) for _ in range(2)]
    dbuf = [device_views(b_) for b_ in dflat]
    obuf = [torch.empty((n, 1024, h, w), dtype=torch.float32, device=dev) for _ in range(2)]
    hout = [torch.empty((n, 1024, h, w), dtype=torch.float32).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(steps):
        for i in range(steps + 1):
            if i < steps:                      # stage the inputs of step i
                b = i & 1
                with torch.cuda.stream(s_in):
                    if i >= 2:
                        s_in.wait_event(ev_free[b])          # step i-2 finished reading this buffer
                    dflat[b][plane:].copy_(hflat[i % len(hflat)], non_blocking=True)
                    ev_in[b].record(s_in)
            if i >= 1:                         # run step i-1 and read its result back
                b = (i - 1) & 1
                main.wait_event(ev_in[b])
                if i >= 3:
                    main.wait_event(ev_out[b])               # the D2H of step i-3 released this output buffer
                rm.step(dbuf[b]["k4"], dbuf[b]["v4"], dbuf[b]["mask"][None], dbuf[b]["flow"][None], dbuf[b]["qk"], dbuf[b]["qv"],
                        commit=False, out=obuf[b])
                ev_free[b].record(main)
                ev_done[b].record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[b])
                    hout[b].copy_(obuf[b], non_blocking=True)
                    ev_out[b].record(s_out)
        torch.cuda.synchronize()

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    e2e_s = time.perf_counter() - t0
    t_end = time.perf_counter()
    clk = clocks.stop(t_begin, t_end) if clocks else None
    assert torch.isfinite(hout[(args.steps - 1) & 1]).all()

    # ---- roofline of the dominant kernel (the split-KV attention kernel), timed alone with events
    st = rm.bank.stats()
    cells = (st[:n, 0] + st[:n, 1]).astype(np.int64)
    _, rq_all = ops.regional_boxes(dframes[0]["mask"][None], dframes[0]["flow"][None], padded_frame=False)
    rq = rq_all[0, 1:n + 1].contiguous()
    rq_h = rq.cpu().numpy()
    nq = [max(0, int(r[1] - r[0] + 1)) * max(0, int(r[3] - r[2] + 1)) for r in rq_h]
    flops, byts = algorithmic_work(cells.tolist(), nq, N)
    # tensor-core products per MAC of the algorithmic flops: strict 3, fast 1, mixed = 3 for Q.K^T (128 of the 640 channels) and 1 for P.V
    passes = {"single": 1.0, "mixed": (3.0 * 128 + 512) / 640}.get(args.precision, 3.0)
    d0 = dframes[0]
    rm.bank.read(d0["qk"], d0["qv"], rq, n, precision, out=m4)   # all stages once: the query side of d0 is now in the workspace
    kt, mt = [], []
    for i in range(max(args.steps, 5)):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        rm.bank.read(d0["qk"], d0["qv"], rq, n, precision, stages=1, out=m4)
        e1.record()
        rm.bank.read(d0["qk"], d0["qv"], rq, n, precision, stages=2, out=m4)
        e2.record()
        torch.cuda.synchronize()
        kt.append(e0.elapsed_time(e1))
        mt.append(e1.elapsed_time(e2))
    k_ms = float(np.mean(kt))
    achieved_tf = flops * passes / (k_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "memory_read_umma_kernel", "achieved": achieved_tf, "peak": tc_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / tc_peak, "traffic": None, "peak_source": peak_src, "kernel_ms": k_ms, "merge_ms": float(np.mean(mt)),
            "passes": passes, "algorithmic_gflop": flops / 1e9, "algorithmic_mbytes": byts / 1e6,
            "hbm_achieved_gbs": byts / (k_ms * 1e-3) / 1e9, "hbm_frac": byts / (k_ms * 1e-3) / 1e9 / hbm_peak,
            "in_region_fraction": float(np.mean([c / (T * N) for c in cells])), "share_of_step": k_ms / (dev_ms / args.steps)}
    try:   # the work plan the timed launches walked (built on the device, rmnet_b200/csrc/sched.cuh): a diagnostic, never fatal
        ns_plan, lists = rm.bank.read_plan(n)
        loads = [sum(p[5] for p in pcs) for pcs in lists]
        roof["plan"] = {"ctas": len(lists), "ctas_busy": int(sum(1 for v in loads if v)), "kv_tiles_per_cta_max": int(max(loads)),
                        "kv_tiles_per_cta_mean": float(sum(loads)) / max(len(loads), 1), "pieces_per_cta_max": int(max(len(p) for p in lists)),
                        "kv_chunks_per_object": [int(v) for v in ns_plan]}
    except Exception as e:  # noqa: BLE001
        roof["plan"] = {"unavailable": repr(e)[:200]}
    try:   # DRAM traffic per launch of the same kernel from the committed `ncu --set full` capture, when present
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = args.workload if args.precision == "split3" else f"{args.workload}_{args.precision}"   # (captures exist for the strict mode)
        roof["traffic"] = tr.get(key, {}).get("dram_bytes_per_launch")
        roof["traffic_source"] = f"profiles/traffic.json[{key}] <- {tr.get(key, {}).get('report')}" if key in tr else None
    except (OSError, ValueError):
        pass

    # ---- the reference's own composition of the same step on THIS GPU (torch CUDA ops + the unmodified reference CUDA
    #      kernel from oracle/_ref; tests/ref_composition.py), rank 0 at N = 1 only: the GPU-vs-GPU bar, reported beside
    #      the CPU baseline.  Skipped (with the reason) when the reference extension was not built.
    ref_gpu = None
    if world == 1 and args.workload != "c4":
        try:
            import glob
            so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "reg_att_map_generator*.so"))
            if not so:
                raise RuntimeError("oracle/_ref/reg_att_map_generator*.so not built")
            sys.path.insert(0, os.path.dirname(so[0]))
            import reg_att_map_generator as ref_gen
            from ref_composition import ReferenceClip
            rc = ReferenceClip(ref_gen, n, K_CH, H, W)
            for t in range(T - 1):
                rc.commit({k: torch.from_numpy(v).to(dev) for k, v in frames[t].items()})
            for _ in range(2):
                rc.step(dframes[0])
            ts = []
            for _ in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); m_ref, _, _ = rc.step(dframes[0]); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ours0 = step_dev(dframes[0])
            ref_gpu = {"ms_per_step": float(np.median(ts)), "frames_per_s": 1e3 / float(np.median(ts)),
                       "max_abs_diff_mem_val": float((ours0 - m_ref).abs().max().item()),
                       "what": "models/rmnet.py:191-205,:212,:244-248,:252-287,:307,:355-358,:416-426,:147-165 restated with torch's CUDA ops "
                               "+ the unmodified reference CUDA kernel (oracle/_ref), same inputs, L2 flushed"}
            del rc, m_ref
        except Exception as e:
            ref_gpu = {"unavailable": f"{type(e).__name__}: {e}"}

    # ---- the e2e leg: VOS frames/sec through the reference-facing API, clips sharded over the ranks (run_vos)
    del dflat, dbuf, obuf, flush
    torch.cuda.empty_cache()
    vos_s, vos_frames, vos_info, vos_extra = None, 0, None, None
    import baseline
    if not args.no_vos and not baseline.available():
        # the same on every rank (the tree ships with the snapshot or not at all): no collective is entered
        vos_info = {"error": "baseline/_ref (the reference tree) is not on this box: run __graft_entry__.build() in the build container"}
    elif not args.no_vos:
        try:
            vos_s, vos_frames, vos_info, vos_extra = run_vos(args, rank, world, local_rank, dev, precision)
        except Exception as e:
            if world > 1:
                raise
            import traceback
            vos_info = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-1200:]}

    # ---- max over ranks (device time, op-level e2e time, VOS time); sum of the frames segmented; checksums gathered
    dev_ms, e2e_s, sums = reduce_over_ranks(dev_ms, e2e_s, checksum, dev, rank, world)
    vos_max_s, vos_total_frames = vos_s, vos_frames
    if world > 1 and vos_s is not None:
        import torch.distributed as dist
        t = torch.tensor([vos_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f = torch.tensor([vos_frames], device=dev, dtype=torch.float64)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
        vos_max_s, vos_total_frames = float(t[0]), int(f[0])
    if rank != 0:
        return None
    fps = world * args.steps / (dev_ms * 1e-3)
    e2e_op = {"value": world * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
              "what": "op-level: one pinned host staging buffer per frame (mask channels 1..n, flow, k4, v4, q_key, q_val) -> one H2D copy -> "
                      "RegionalMemory.step -> D2H of mem_val every step; copies double-buffered on side streams"}
    if vos_s is not None:
        e2e = {"value": vos_total_frames / vos_max_s, "unit": "frames/s", "h2d_bytes_per_step": vos_info["h2d_bytes_per_frame"],
               "d2h_bytes_per_step": vos_info["d2h_bytes_per_frame"], "frames": vos_total_frames, "seconds_max_over_ranks": vos_max_s,
               "what": "VOS frames/sec through the reference-facing API: tflownet(frames) + rmnet(frames, masks, flows, n_objects, 5) "
                       "(utils/helpers.py:55-56) on host tensors, unmodified reference nets wrapped like core/inference.py:35-37, "
                       "rmnet_b200.install() behind them; a step = one segmented frame", **{k: v for k, v in vos_info.items() if k not in ("h2d_bytes_per_frame", "d2h_bytes_per_frame")}}
    else:
        e2e = dict(e2e_op, note="VOS leg not run: " + str(vos_info))
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(np.max(step_ms)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"split3": "fp16 hi/lo planes x3 products, fp32 accumulate", "single": "fp16 x1 product, fp32 accumulate",
                  "mixed": "fp16 hi/lo x3 products for Q.K^T, x1 for P.V, fp32 accumulate"}[args.precision], "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}; one step = one frame of the regional memory-read path "
                               "(both region descriptors -> pack-at-memorise + query side -> tcgen05 regional read of all objects -> merge; "
                               "4 kernels chained by programmatic dependent launch: RegionalMemory.step)",
                   "clips_per_gpu": 1, "parallelism": f"clip-parallel x{world} (no data-path collective; one NCCL gather of the label maps in the e2e leg)",
                   "precision": args.precision,
                   "l2": "flushed between timed steps (256 MiB write); per-step CUDA events summed", "pool_frames": POOL,
                   "e2e_workload": (f"{VOS_CLIPS_PER_GPU} clips per GPU, {VOS_SHAPE['H']}x{VOS_SHAPE['W']}, {VOS_SHAPE['n']} objects, F={VOS_SHAPE['F']}, "
                                    if args.e2e_set == "default" else
                                    f"30 clips shaped like DAVIS-2017 val (BASELINE configs[2]): {VOS_SHAPE['H']}x{VOS_SHAPE['W']}, 1..5 objects cycling, 34..104 frames (1 999 in total), ")
                                   + f"memorize_every={VOS_SHAPE['every']}, K={K_CH}; sharded longest-first, each rank pinned to its own host cores ({core_slice})"},
        "e2e": e2e, "e2e_op": e2e_op, "vos": vos_extra,
        "gpu_launches": launches, "cuda_graph": graph_info, "precision_modes": modes_info, "reference_on_this_gpu": ref_gpu, "roofline": roof, "clocks": clk, "result_checksums": sums,
    }
    return line, pool


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    # default = the shape north_star quotes its target on (480p, T=20 memory, 5 objects = BASELINE configs[2]'s frame);
    # c2 is BASELINE configs[1] (3 objects, T=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="split3", choices=["split3", "single", "mixed"])
    ap.add_argument("--cpu-steps", type=int, default=0, help="steps of the CPU baseline sample (0 = auto, ~10-30 s)")
    ap.add_argument("--no-vos", action="store_true", help="skip the VOS e2e leg (op-level legs only)")
    ap.add_argument("--e2e-set", default="default", choices=["default", "davis30"],
                    help="clips of the e2e leg: 8 per GPU, 5 objects, F=60 (default) | the 30 DAVIS-val-shaped clips of BASELINE configs[2]")
    ap.add_argument("--no-vos-extras", action="store_true", help="skip the F=100 clip / reference-on-this-GPU / module split")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        # The reference's own implementation on the host cores (all of them), rank 0 only.
        if rank != 0:
            return
        pool = make_pool(wl, 1234, 4)
        warm = min(args.warmup, 2)
        fps, ms, steps, kind, sample = time_cpu_arm(wl, pool, args.steps, warm)
        cores = os.cpu_count()
        vos_cpu = cpu_vos_sample()
        e2e = {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        if "fps" in vos_cpu:
            e2e = {"value": vos_cpu["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "what": "VOS frames/sec of the unmodified tflownet + rmnet on the host cores -- the counterpart of the GPU arm's e2e "
                           "(`value` is the counterpart of the GPU arm's `value`: the hot-path step alone)", **vos_cpu}
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}; one step = one frame of the regional memory-read path", "device": "host CPU",
                       "steps_requested": args.steps, "note": "steps are cut at a 90 s budget"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "cpu": cpu_model(), "sample": sample},
            "e2e": e2e}))
        return

    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = run_gpu(args, wl, rank, world, local_rank)
    if rank == 0:
        line, pool = res
        if world == 1:
            steps = args.cpu_steps or {"c2": 20, "c3": 10}.get(args.workload, 1)
            fps, ms, steps, kind, sample = time_cpu_arm(wl, make_pool(wl, 1234, 4), steps, 1, budget_s=30.0)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": kind, "cpu": cpu_model(), "ms_per_step": ms,
                                    "sample": sample, "vos": None if args.no_vos else cpu_vos_sample()}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_vos_sample(n_frames=3):
    """VOS frames/sec of the UNMODIFIED reference (tflownet + rmnet from baseline/_ref, torch CPU, all host threads; the
    CUDA-only generator replaced by the C oracle) on a bounded sample: the first `n_frames` frames of an e2e-leg clip."""
    try:
        import torch
        import baseline
        from baseline import vos
        if not baseline.available():
            return {"unavailable": "baseline/_ref not populated"}
        threads = vos.set_host_threads()
        H, W, n, every = (VOS_SHAPE[k] for k in ("H", "W", "n", "every"))
        tfn, net = baseline.build_nets(0, "cpu", cpu_generator=True)
        frames, masks, n_objects = baseline.synthetic_clip(5000, n, n_frames, H, W)
        cuda_was = torch.cuda.is_available
        torch.cuda.is_available = lambda: False          # utils/helpers.py:18 var_or_cuda would move tensors to the GPU of the box
        try:
            _, s, sf = vos.run_clip(tfn, net, frames, masks, n_objects, every)
        finally:
            torch.cuda.is_available = cuda_was
        return {"fps": (n_frames - 1) / s, "seconds": s, "flownet_seconds": sf, "cores": threads, "kind": "reference",
                "sample": f"first {n_frames} frames ({n_frames - 1} segmented) of a {H}x{W}, {n}-object clip; unmodified models/rmnet.py + "
                          f"models/tiny_flownet.py on torch CPU, {threads} threads"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


if __name__ == "__main__":
    main()
