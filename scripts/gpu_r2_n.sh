# round 2: N-GPU check of the in-library sharded step + the bench line at N GPUs.  usage: gpu_r2_n.sh N [extras]
set -x
N=${1:-2}
EX=${2:-auto}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r2_multigpu_check_n$N.log 2>&1; tail -6 gpurun_out/r2_multigpu_check_n$N.log | cut -c 1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --extras $EX > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 5000 gpurun_out/r2_bench_n$N.json; tail -15 gpurun_out/r2_bench_n$N.err
