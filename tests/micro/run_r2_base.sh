set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_r2_base.log; tail -5 gpurun_out/pytest_r2_base.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err; tail -c 2500 gpurun_out/bench_r2_base.json
