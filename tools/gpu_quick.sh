#!/bin/bash
# quick GPU check: parity tests + both bench workloads + per-op breakdown + attention-kernel timeline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --workload c2 --cpu-steps 2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 400 python bench.py --workload c3 --cpu-steps 1 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python tools/step_breakdown.py > gpurun_out/step_breakdown.log 2>&1; echo "breakdown rc=$?"
timeout 300 python tools/umma_timeline.py > gpurun_out/umma_timeline.log 2>&1; echo "timeline rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2.json", "gpurun_out/bench_c3.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "fps", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "kernel_us", round(r["kernel_ms"] * 1e3, 1), "merge_us", round(r["merge_ms"] * 1e3, 1), "frac", round(r["frac"], 3), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-2000:])
PY
cat gpurun_out/step_breakdown.log gpurun_out/umma_timeline.log
