#!/bin/bash
# round-2 session 3: suite after the fp16-plane default / two-stream frame body, sanitizer evidence, bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^\[key|^\[c[1234]_|passed|failed|rror" gpurun_out/r2_pytest_gpu.log | tail -24
SEL='step_equals_memorize_then_read or captured_step or step_variants or regional_path_vs_oracle or flow_affine_bit_exact or generator_bit_exact or mask_epilogue_vs_golden'
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_c3.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("seconds_max_over_ranks"), "vos", json.dumps(d["vos"])[:1500])
PY
tail -3 gpurun_out/r2_bench_c3.err
