# what the driver runs at round end on one GPU: the parity suite and smoke()
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_sanity_pytest.log 2>&1; tail -3 gpurun_out/r2_sanity_pytest.log | cut -c 1-200
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c 1-400
