#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_rmnet.py tests/test_gpu_parity.py -m gpu -q -s -x -k "fast_precision or formats_and_precisions" > gpurun_out/r2_mixed_tests.log 2>&1; echo "mixed tests rc=$?"; grep -E "^\[|^\.\[|passed|failed|rror" gpurun_out/r2_mixed_tests.log | tail -14
WLS="c3 c2 c4" PRS="mixed" bash tools/gpu_quick_bench.sh
