#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/step_breakdown.py > gpurun_out/step_breakdown.log 2>&1; echo "breakdown rc=$?"
timeout 300 python tools/umma_timeline.py > gpurun_out/umma_timeline.log 2>&1; echo "timeline rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"memory_read_umma|merge_kernel|frame_boxes|bank_pack" -s 60 -c 8 -o gpurun_out/full_c3 -f python bench.py --workload c3 --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_f3.log 2>&1; echo "ncu full c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"memory_read_umma|merge_kernel|frame_boxes|bank_pack" -s 12 -c 8 -o gpurun_out/full_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_f2.log 2>&1; echo "ncu full c2 rc=$?"
cat gpurun_out/step_breakdown.log gpurun_out/umma_timeline.log
