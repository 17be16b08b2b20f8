set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "device_trainer or coarse_tensor or constructor" > gpurun_out/r2g_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2g_pytest_gpu.log | cut -c 1-300
