# the two CTA shapes of the scan kernel and the automatic choice, workloads B and C on one box
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2j_pytest_gpu.log | cut -c 1-200
for shape in 12 16 auto; do for wl in B C; do
if [ $shape = auto ]; then unset IVFADC_SCANW_SHAPE; else export IVFADC_SCANW_SHAPE=$shape; fi
timeout -s KILL 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --check 64 --extras none 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('shape $shape', '$wl', 'scan_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.4f' % d['roofline']['frac'], 'step_ms %.4f' % d['ms_per_step'], 'parity', d['parity']['ok'])
"
done; done 2>&1 | grep -v "^+" | tee gpurun_out/shape_result.txt
