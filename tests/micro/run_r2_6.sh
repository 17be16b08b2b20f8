mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "qlane or tables or kernel_choice or config_b or search" 2>&1 | tail -12
bash tests/micro/ab.sh 20 variants/f16.so variants/nofused.so
timeout 300 python tests/debug_timeline_w.py > gpurun_out/timeline_w4.txt 2>&1; tail -17 gpurun_out/timeline_w4.txt
