#!/bin/bash
# release library (as shipped): parity + plan tests + op-level bench lines; then a DEV rebuild on the box for the chain timeline
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_plan.py -m gpu -q -x > gpurun_out/plan_pytest.log 2>&1; echo "plan pytest rc=$?"; tail -1 gpurun_out/plan_pytest.log
bash tools/gpu_quick.sh
make -C rmnet_b200/csrc clean > /dev/null; make -C rmnet_b200/csrc -j16 DEV=1 > gpurun_out/make.log 2>&1; echo "dev build rc=$?"
timeout 120 python tools/chain_timeline.py > gpurun_out/chain_tl.log 2>&1; grep -v "np.float64" gpurun_out/chain_tl.log
