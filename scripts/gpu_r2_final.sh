# final single-GPU record of the round: the driver's own command lines, then the launch list and one full ncu capture of the scan
set -x
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_B.json 2> gpurun_out/r2_final_bench_B.err; tail -3 gpurun_out/r2_final_bench_B.err
timeout -s KILL 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; tail -3 gpurun_out/r2_final_bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_B.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_final_bench_ref.json').read().strip().splitlines()[-1])
print('B', d['value'], d['ms_per_step'], d['breakdown_ms'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], 'same_config', d['config']==r['config'], 'ref', r['value'])
for k,v in d['extra'].items():
    if 'error' in v: print(k, v); continue
    if k=='E': print('E', v['device_resident']['value'], v['host_api']['value'], v['delete']['seconds'], v['parity']); continue
    print(k, v['value'], v['ms_per_step'], v['breakdown_ms'], 'frac', v['roofline']['frac'], v['roofline']['kernel'][:30], v['parity']['ok'])
PY
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ivf -c 400 --csv --log-file gpurun_out/r2_final_launches_B.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2_final_launches.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:scanw_kernel -s 4 -c 1 -f -o gpurun_out/r2_final_scanw python bench.py --steps 2 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2_final_ncu_scanw.log 2>&1
ls -la gpurun_out | tail -6
