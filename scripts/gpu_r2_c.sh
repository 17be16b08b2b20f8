# round 2, step C on one B200: parity suite after the coarse redo-list change, bench (B + extras C, D, E), sanitizer, H3 capture
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2c_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_B.json 2> gpurun_out/r2c_bench_B.err; tail -5 gpurun_out/r2c_bench_B.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_B.json').read().strip().splitlines()[-1])
print('B', d['value'], d['ms_per_step'], d['breakdown_ms'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k,v in d['extra'].items():
    if 'error' in v: print(k, v); continue
    if k=='E': print('E', v['device_resident'], v['host_api'], v['delete'], v['parity']); continue
    print(k, v['value'], v['ms_per_step'], v['breakdown_ms'], 'frac', v['roofline']['frac'], v['roofline']['kernel'][:30], v['parity']['ok'])
PY
timeout 900 bash scripts/sanitize.sh "qlane_bit_exact or coarse_tensor_core_bit_exact and 1000" 
# H3: the list scan with the code array beyond the L2 (workload D unsharded, 800 MB of codes): DRAM bytes next to the algorithmic bytes
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ivf::scan_kernel|ivf::scanw_kernel" -s 3 -c 1 -f -o gpurun_out/r2c_scan_D_n1 python bench.py --workload D --steps 2 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2c_ncu_D.log 2>&1
tail -3 gpurun_out/r2c_ncu_D.log | cut -c 1-300
ls -la gpurun_out | tail -8
