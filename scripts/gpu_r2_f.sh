set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2f_pytest_gpu.log | cut -c 1-300
timeout -s KILL 200 python tests/debug_coarse_redo.py 2>&1 | tail -4 | cut -c 1-110
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c 1-300
