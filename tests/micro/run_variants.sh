#!/bin/bash
# bring-up helper (GPU box): timeline + scan ms of each variant library
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp ivfadc.jl_b200/libivfadc_cuda.so /tmp/lib_orig.so
for v in "$@"; do
  cp ivfadc.jl_b200/build/variants/$v.so ivfadc.jl_b200/libivfadc_cuda.so
  echo "=== $v"
  timeout 120 python tests/debug_timeline.py 2>&1 | tail -3 | cut -c1-400
  timeout 200 python bench.py --steps 10 > gpurun_out/bench_$v.log 2>&1
  grep -o "\"breakdown_ms\": {[^}]*}" gpurun_out/bench_$v.log; grep -o "\"parity\": {[^}]*}" gpurun_out/bench_$v.log
done
cp /tmp/lib_orig.so ivfadc.jl_b200/libivfadc_cuda.so
