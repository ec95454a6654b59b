"""GPU probe: VOS frames/s of the unmodified reference vs the installed path (eager / graph), per-module split.
usage: python tools/vos_probe.py [workload] [frames] [every] [lean]"""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import baseline
from baseline import vos
import rmnet_b200

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
F_ = int(sys.argv[2]) if len(sys.argv) > 2 else 16
every = int(sys.argv[3]) if len(sys.argv) > 3 else 5
H, W, n = vos.WORKLOADS[wl]
dev = torch.device("cuda", 0)
ref = baseline.import_reference(need_cuda_extension=True)
vos.reference_flags()
tfn, net = baseline.build_nets(0, dev, cpu_generator=False)
frames, masks, n_objects = baseline.synthetic_clip(3, n, F_, H, W)
res = {"workload": wl, "frames": F_, "every": every, "allow_tf32": torch.backends.cudnn.allow_tf32}
tfn_dp, net_dp = vos.wrap(tfn, net)

def leg(name, fn):
    try:
        fn()
    except Exception:
        res[name] = "FAILED: " + traceback.format_exc()[-1500:]

def ref_leg():
    rmnet_b200.uninstall(ref)
    vos.run_clip(tfn_dp, net_dp, frames[:, :3], masks[:, :3], n_objects[:, :3], every)
    probs, s, sf = vos.run_clip(tfn_dp, net_dp, frames, masks, n_objects, every)
    res["reference"] = {"fps": (F_ - 1) / s, "s": s, "flownet_s": sf}
    res["_ref_lab"] = probs[0].argmax(1).cpu()

def our_leg(graph, output="reference"):
    def f():
        rmnet_b200.install(ref, use_graph=graph, output=output)
        try:
            vos.run_clip(tfn_dp, net_dp, frames, masks, n_objects, every)   # warm (captures graphs)
            probs, s, sf = vos.run_clip(tfn_dp, net_dp, frames, masks, n_objects, every)
            lab = probs[0].argmax(1).cpu()
            agree = float((lab == res["_ref_lab"]).float().mean()) if "_ref_lab" in res else None
            res[f"ours_graph{int(graph)}_{output}"] = {"fps": (F_ - 1) / s, "s": s, "flownet_s": sf, "label_agreement": agree}
        finally:
            rmnet_b200.uninstall(ref)
    return f

lean = len(sys.argv) > 4 and sys.argv[4] == "lean"     # reference + the default installed path only (large workloads)
leg("reference", ref_leg)
if lean:
    leg("ours_graph", our_leg(True))
    res.pop("_ref_lab", None)
    print(json.dumps(res, indent=1))
    sys.exit(0)
leg("ours_eager", our_leg(False))
leg("ours_graph", our_leg(True))
leg("ours_graph_device", our_leg(True, "device"))
def cl_leg(bench_flag):
    def f():
        torch.backends.cudnn.benchmark = bench_flag
        net.to(memory_format=torch.channels_last); tfn.to(memory_format=torch.channels_last)
        try:
            our_leg(True, "host")()
            res[f"ours_channels_last_benchmark{int(bench_flag)}"] = res.pop("ours_graph1_host")
        finally:
            net.to(memory_format=torch.contiguous_format); tfn.to(memory_format=torch.contiguous_format)
            torch.backends.cudnn.benchmark = False
    return f

def bench_only_leg():
    torch.backends.cudnn.benchmark = True
    try:
        our_leg(True, "host")()
        res["ours_nchw_benchmark1"] = res.pop("ours_graph1_host")
    finally:
        torch.backends.cudnn.benchmark = False

leg("nchw_benchmark", bench_only_leg)
leg("channels_last", cl_leg(False))
leg("channels_last_benchmark", cl_leg(True))
leg("split", lambda: res.__setitem__("module_split_ms", vos.module_split(net, tfn, H, W, n, 5, dev)))
res.pop("_ref_lab", None)
print(json.dumps(res, indent=1))
