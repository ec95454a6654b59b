#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list + full capture.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 400 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rmnet -s 40 -c 12 -o gpurun_out/full_c3 -f python bench.py --workload c3 --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench_c2.json gpurun_out/bench_c3.json
