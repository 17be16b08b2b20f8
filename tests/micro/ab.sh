# A/B of library variants on one box: tests/micro/ab.sh <steps> <variant.so> [<variant.so> ...]; two rounds, interleaved.
# WORKLOADS="B C" selects the bench workloads (default B).
steps=$1; shift
cp ivfadc.jl_b200/libivfadc_cuda.so /tmp/lib_product.so
for round in 1 2; do
  for v in "$@"; do
    cp "$v" ivfadc.jl_b200/libivfadc_cuda.so
    for wl in ${WORKLOADS:-B}; do
    timeout -s KILL 200 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --check 64 --extras none 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', '$wl', 'scan_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.4f' % d['roofline']['frac'], 'step_ms %.4f' % d['ms_per_step'], 'parity', (d['parity'] or {}).get('ok'), d['clocks']['sm_mhz'], d['breakdown_ms'])
"
    done
  done
done
cp /tmp/lib_product.so ivfadc.jl_b200/libivfadc_cuda.so
