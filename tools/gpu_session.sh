#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines (default = c3, c2, reference arm).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench (default, c3) rc=$?"
timeout 400 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "bench ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_c3.json", "gpurun_out/bench_c2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d["roofline"]
        print(f, "fps", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "kernel_us", round(r["kernel_ms"] * 1e3, 1),
              "frac", round(r["frac"], 3), "traffic", r["traffic"], "cpu", round(d["cpu_baseline"]["value"], 2), "ref_gpu", d.get("reference_on_this_gpu"), "graph", d.get("cuda_graph"))
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
print(open("gpurun_out/bench_ref.json").read()[-600:])
PY
