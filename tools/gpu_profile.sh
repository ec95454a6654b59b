#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of the default bench command's op legs + one --set full capture of the
# chain's kernels (strict c3, strict c2, fast c3).  A number printed under ncu is never a bench value.
set -u
mkdir -p gpurun_out
K='memory_read_umma|merge_kernel|frame_boxes|bank_pack'
B="--steps 3 --warmup 3 --cpu-steps 1 --no-vos"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --workload c3 $B > gpurun_out/ncu_l3.log 2>&1; echo "launch list c3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --workload c2 $B > gpurun_out/ncu_l2.log 2>&1; echo "launch list c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 40 -c 8 -o gpurun_out/r2_full_c3 -f python bench.py --workload c3 $B > gpurun_out/ncu_f3.log 2>&1; echo "ncu full c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 16 -c 8 -o gpurun_out/r2_full_c2 -f python bench.py --workload c2 $B > gpurun_out/ncu_f2.log 2>&1; echo "ncu full c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"memory_read_umma" -s 10 -c 2 -o gpurun_out/r2_full_c3_fast -f python bench.py --workload c3 --precision single $B > gpurun_out/ncu_f3f.log 2>&1; echo "ncu full c3 fast rc=$?"
ls -la gpurun_out/r2_*.ncu-rep
