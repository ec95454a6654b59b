#!/bin/bash
# ncu evidence for profiles/: launch lists of the bench command (both workloads) + one --set full capture per kernel.
set -u
mkdir -p gpurun_out
K='memory_read_umma|merge_kernel|frame_boxes|bank_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_l2.log 2>&1; echo "launch list c2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_l3.log 2>&1; echo "launch list c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 40 -c 8 -o gpurun_out/full_c3 -f python bench.py --workload c3 --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_f3.log 2>&1; echo "ncu full c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 16 -c 8 -o gpurun_out/full_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --cpu-steps 1 > gpurun_out/ncu_f2.log 2>&1; echo "ncu full c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_epilogue -s 4 -c 2 -o gpurun_out/full_epilogue -f python tools/epilogue_probe.py > gpurun_out/ncu_fe.log 2>&1; echo "ncu full epilogue rc=$?"
ls -la gpurun_out/*.ncu-rep
