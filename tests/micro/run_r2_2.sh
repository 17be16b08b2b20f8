set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "qlane or tables or kernel_choice or config_b or search" 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w2.json 2> gpurun_out/bench_w2.err; tail -c 1200 gpurun_out/bench_w2.json
timeout 300 python tests/debug_timeline_w.py > gpurun_out/timeline_w2.txt 2>&1; tail -45 gpurun_out/timeline_w2.txt
