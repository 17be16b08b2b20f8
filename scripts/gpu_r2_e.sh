set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2e_pytest_gpu.log
IVFADC_BENCH_N=196608 timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --check 64 --extras none 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('B with 192 vectors per list: scan_ms %.4f' % d['roofline']['kernel_ms'], d['roofline']['kernel'][:20], 'parity', d['parity']['ok'])"
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench_B.json 2> gpurun_out/r2e_bench_B.err; tail -3 gpurun_out/r2e_bench_B.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench_B.json').read().strip().splitlines()[-1])
print('B', d['value'], d['ms_per_step'], d['breakdown_ms'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k,v in d['extra'].items():
    if 'error' in v: print(k, v); continue
    if k=='E': print('E', v['device_resident']['value'], v['host_api']['value'], v['delete']['seconds'], v['parity']); continue
    print(k, v['value'], v['ms_per_step'], v['breakdown_ms'], 'frac', v['roofline']['frac'], v['roofline']['kernel'][:30], v['parity']['ok'])
PY
