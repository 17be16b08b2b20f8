set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "qlane or tables or kernel_choice or config_b" 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -c 1500 gpurun_out/bench_w.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --flags 16 > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; tail -c 1500 gpurun_out/bench_v1.json
timeout 300 python tests/debug_timeline_w.py > gpurun_out/timeline_w.txt 2>&1; tail -50 gpurun_out/timeline_w.txt
