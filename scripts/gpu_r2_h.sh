set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2h_pytest_gpu.log | cut -c 1-300
