#!/bin/bash
# compute-sanitizer evidence for profiles/: memcheck + racecheck over the chained step (PDL), the captured graph, the TMA-fed
# tcgen05 kernel behind the device-built plan, the ticket-finalised scans, the persistent merge
set -u
mkdir -p gpurun_out
SEL='step_equals_memorize_then_read or captured_step or step_variants or regional_path_vs_oracle or flow_affine_bit_exact or generator_bit_exact or mask_epilogue_vs_golden or (plan_covers and seed2) or (plan_covers and seed4) or plan_is_rebuilt'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r2_sanitizer_$tool.log
done
