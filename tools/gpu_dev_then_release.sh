#!/bin/bash
# dev-build probes first (the tree ships with a `make DEV=1` library), then a release rebuild on the box + parity + bench lines
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_plan.py -m gpu -q -x -s > gpurun_out/plan_pytest.log 2>&1; echo "plan pytest rc=$?"; grep -E "plan:|passed|failed|Error|error" gpurun_out/plan_pytest.log | tail -25
timeout 120 python tools/umma_timeline.py standalone > gpurun_out/tl_standalone.log 2>&1; echo "timeline rc=$?"
timeout 120 python tools/chain_timeline.py > gpurun_out/chain_tl.log 2>&1; cat gpurun_out/chain_tl.log
make -C rmnet_b200/csrc clean > /dev/null; make -C rmnet_b200/csrc -j16 > gpurun_out/make.log 2>&1; echo "release build rc=$?"
bash tools/gpu_quick.sh
