mkdir -p gpurun_out
timeout 300 python tests/debug_timeline_w.py > gpurun_out/timeline_w4.txt 2>&1; tail -20 gpurun_out/timeline_w4.txt
