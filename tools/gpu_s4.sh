#!/bin/bash
# round-2 session 4: fast-mode test, conv-layout probe, op-level bench lines (strict / fast; c2, c3, c4)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rmnet.py -m gpu -q -s -k "fast_precision" > gpurun_out/r2_fast_tests.log 2>&1; echo "fast tests rc=$?"; grep -E "^\[|passed|failed" gpurun_out/r2_fast_tests.log
timeout 600 python tools/vos_probe.py c3 31 5 > gpurun_out/r2_vos_c3_layouts.json 2> gpurun_out/r2_vos_c3_layouts.err; echo "probe rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_vos_c3_layouts.json"))
for k,v in d.items():
    if isinstance(v,dict) and "fps" in v: print(k, round(v["fps"],1), v.get("label_agreement"))
    elif isinstance(v,str): print(k, v[-300:])
PY
for wl in c2 c3 c4; do for pr in split3 single; do
  timeout 600 python bench.py --workload $wl --precision $pr --steps 50 --warmup 5 --no-vos --cpu-steps 1 > gpurun_out/r2_op_${wl}_${pr}.json 2> gpurun_out/r2_op_${wl}_${pr}.err; echo "bench $wl $pr rc=$?"
done; done
python - <<'PY'
import json
for wl in ("c2","c3","c4"):
    for pr in ("split3","single"):
        try:
            d=json.loads(open(f"gpurun_out/r2_op_{wl}_{pr}.json").read().strip().splitlines()[-1]); r=d["roofline"]
            print(wl, pr, "step_us", round(d["ms_per_step"]*1e3,1), "fps", round(d["value"]), "kernel_us", round(r["kernel_ms"]*1e3,1), "tc_frac", round(r["frac"],3), "hbm_frac", round(r["hbm_frac"],3), "merge_us", round(r["merge_ms"]*1e3,1), "ref_gpu", (d.get("reference_on_this_gpu") or {}).get("max_abs_diff_mem_val"))
        except Exception as e:
            print(wl, pr, "unreadable", e)
PY
