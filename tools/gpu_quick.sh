#!/bin/bash
# inner loop of kernel work: parity tests of the read path + op-level bench lines (no VOS leg)
set -u
mkdir -p gpurun_out
timeout ${PYTEST_TIMEOUT:-400} python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -q -x > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/quick_pytest.log
bash tools/gpu_quick_bench.sh
