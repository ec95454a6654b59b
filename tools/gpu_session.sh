#!/bin/bash
# One GPU-box session: the whole -m gpu suite, smoke(), the default bench line and the reference arm.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench (default, c3) rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_c3.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "kernel_us", round(r["kernel_ms"] * 1e3, 1), "frac", round(r["frac"], 3),
      "modes", {k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["precision_modes"].items() if isinstance(v, dict)}, "checksum", d["result_checksums"],
      "vos", {k: (round(v["fps"], 1) if isinstance(v, dict) and "fps" in v else None) for k, v in (d["vos"] or {}).items() if k in ("ours", "reference_on_this_gpu")},
      "cpu", round(d["cpu_baseline"]["value"], 2), d["cpu_baseline"]["vos"])
ref = json.loads(open("gpurun_out/r2_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm: value", round(ref["value"], 2), "e2e", round(ref["e2e"]["value"], 3), ref["cpu_baseline"]["kind"])
PY
