#!/bin/bash
# compute-sanitizer over the tensor-core kernels on the small parity shapes (run on a GPU box):
#   memcheck  -- out-of-bounds / misaligned accesses in global, shared and tensor memory staging
#   racecheck -- shared-memory hazards between the roles of the warp-specialised kernels (mbarrier hand-offs)
# Logs go to gpurun_out/ (copy the summaries to profiles/).  Usage: bash scripts/sanitize.sh [pytest -k expression]
set -x
K=${1:-"qlane_bit_exact or coarse_tensor_core or tcgen05_tables or ties_overflow"}
mkdir -p gpurun_out
export IVFADC_SANITIZE=1
for tool in memcheck racecheck; do
  timeout 3000 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/r2_sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" gpurun_out/r2_sanitize_$tool.log | tail -5
done
