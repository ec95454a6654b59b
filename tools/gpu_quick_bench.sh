#!/bin/bash
# op-level bench lines only (no pytest): WLS / PRS select workloads and precisions
set -u
mkdir -p gpurun_out
for wl in ${WLS:-c3 c2 c4}; do for pr in ${PRS:-split3 single}; do
  timeout 400 python bench.py --workload $wl --precision $pr --steps 50 --warmup 5 --no-vos --cpu-steps 1 > gpurun_out/quick_${wl}_${pr}.json 2> gpurun_out/quick_${wl}_${pr}.err || tail -5 gpurun_out/quick_${wl}_${pr}.err
done; done
python - <<'PY'
import json, os
for wl in os.environ.get("WLS", "c3 c2 c4").split():
    for pr in os.environ.get("PRS", "split3 single").split():
        try:
            d=json.loads(open(f"gpurun_out/quick_{wl}_{pr}.json").read().strip().splitlines()[-1]); r=d["roofline"]
            print(wl, pr, "step_us", round(d["ms_per_step"]*1e3,1), "fps", round(d["value"]), "kernel_us", round(r["kernel_ms"]*1e3,1), "tc_frac", round(r["frac"],3), "hbm_frac", round(r["hbm_frac"],3), "merge_us", round(r["merge_ms"]*1e3,1), "ref_gpu_diff", (d.get("reference_on_this_gpu") or {}).get("max_abs_diff_mem_val"))
        except Exception as e:
            print(wl, pr, "unreadable", e)
PY
