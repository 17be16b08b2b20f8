# scan time against list length (fixed cost per table round vs cost per lookup): workload B with fewer vectors
mkdir -p gpurun_out
cp ivfadc.jl_b200/libivfadc_cuda.so /tmp/lib_product.so
for v in "$@"; do
cp "$v" ivfadc.jl_b200/libivfadc_cuda.so
for n in 65536 196608 393216 589824 786432 1000000; do
IVFADC_BENCH_N=$n timeout -s KILL 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --check 0 --extras none 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'N $n', 'per list %.0f' % ($n / 1024.0), 'scan_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.4f' % d['roofline']['frac'], 'bytes %d' % d['roofline']['algorithmic_bytes_per_launch'])
"
done
done 2>&1 | tee gpurun_out/fill_result.txt
cp /tmp/lib_product.so ivfadc.jl_b200/libivfadc_cuda.so
