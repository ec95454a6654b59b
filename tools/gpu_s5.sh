#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inference_net.py -m gpu -q -s > gpurun_out/r2_inference_net.log 2>&1; echo "inference_net test rc=$?"; tail -5 gpurun_out/r2_inference_net.log
bash tools/gpu_profile.sh
