set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_sanity_pytest.log 2>&1; tail -3 gpurun_out/r2_sanity_pytest.log | cut -c 1-200
timeout -s KILL 300 python tests/debug_f64.py 2>&1 | tail -3 | cut -c 1-300
