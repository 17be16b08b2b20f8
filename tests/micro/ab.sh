# A/B of library variants on one box: tests/micro/ab.sh <steps> <variant.so> [<variant.so> ...]; two rounds, interleaved
steps=$1; shift
cp ivfadc.jl_b200/libivfadc_cuda.so /tmp/lib_product.so
for round in 1 2; do
  for v in "$@"; do
    cp "$v" ivfadc.jl_b200/libivfadc_cuda.so
    timeout 150 python bench.py --steps $steps --warmup 3 --no-cpu-baseline --check 64 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'scan_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.4f' % d['roofline']['frac'], 'step_ms %.4f' % d['ms_per_step'], 'parity', d['parity']['ok'], d['clocks']['sm_mhz'])
"
  done
done
cp /tmp/lib_product.so ivfadc.jl_b200/libivfadc_cuda.so
