set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "shard or merge or device_api or two_shards" > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
bash scripts/gpu_r2_n.sh 2 none
