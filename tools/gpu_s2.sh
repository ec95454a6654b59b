#!/bin/bash
# round-2 session 2: whole GPU suite + smoke + the new bench (ours, reference arm)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|passed|failed|error|Error" gpurun_out/r2_pytest_gpu.log | tail -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2_bench_c3.json; tail -5 gpurun_out/r2_bench_c3.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_bench_ref.err
