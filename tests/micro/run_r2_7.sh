for n in 192 480 768 1152; do
IVFADC_BENCH_PER_LIST=$n python bench.py --steps 20 --warmup 3 --no-cpu-baseline --check 64 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('per list $n', 'scan_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.4f' % d['roofline']['frac'], 'bytes', d['roofline']['algorithmic_bytes_per_launch'], 'parity', d['parity']['ok'], d['breakdown_ms'])
"
done
