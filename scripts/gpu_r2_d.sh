set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2d_pytest_gpu.log
timeout 300 python tests/debug_coarse_redo.py 2>&1 | tail -5 | cut -c 1-200
timeout 900 python bench.py --steps 20 --warmup 3 --extras C,D > gpurun_out/r2d_bench_B.json 2> gpurun_out/r2d_bench_B.err; tail -5 gpurun_out/r2d_bench_B.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench_B.json').read().strip().splitlines()[-1])
print('B', d['value'], d['ms_per_step'], d['breakdown_ms'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k,v in d['extra'].items():
    if 'error' in v: print(k, v); continue
    print(k, v['value'], v['ms_per_step'], v['breakdown_ms'], 'frac', v['roofline']['frac'], v['roofline']['kernel'][:30], v['parity']['ok'])
PY
for kn in scanw_kernel coarse3_kernel; do
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=ivf,kne=$kn --print-limit 6 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "qlane_bit_exact and 128-16-256-64 or coarse_tensor_core_bit_exact and uniform and 1000" > gpurun_out/r2_racecheck_$kn.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" gpurun_out/r2_racecheck_$kn.log | cut -c 1-220 | tail -8
done
