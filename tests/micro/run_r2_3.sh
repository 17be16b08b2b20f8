mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "qlane or tables or kernel_choice or config_b or search" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w3.json 2> gpurun_out/bench_w3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','breakdown_ms')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['parity'], d['e2e']['value'])
PY
