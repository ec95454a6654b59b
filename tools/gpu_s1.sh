#!/bin/bash
# round-2 session 1: the installed path under the real RMNet (parity + FPS probes)
set -u
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv >> gpurun_out/host.txt
timeout 900 python -m pytest tests/test_gpu_rmnet.py -m gpu -q -s > gpurun_out/r2_rmnet_tests.log 2>&1; echo "rmnet tests rc=$?"; tail -40 gpurun_out/r2_rmnet_tests.log
timeout 400 python tools/vos_probe.py c2 16 5 > gpurun_out/r2_vos_c2.json 2> gpurun_out/r2_vos_c2.err; echo "probe c2 rc=$?"; cat gpurun_out/r2_vos_c2.json; tail -5 gpurun_out/r2_vos_c2.err
timeout 600 python tools/vos_probe.py c3 31 5 > gpurun_out/r2_vos_c3.json 2> gpurun_out/r2_vos_c3.err; echo "probe c3 rc=$?"; cat gpurun_out/r2_vos_c3.json; tail -5 gpurun_out/r2_vos_c3.err
