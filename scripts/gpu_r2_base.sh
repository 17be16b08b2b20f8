# round-2 baseline on one B200: parity suite, bench lines (B, C), launch list, one full ncu capture of the scan
set -x
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_B.json 2> gpurun_out/r2_bench_B.err; tail -c 1500 gpurun_out/r2_bench_B.json
timeout 400 python bench.py --workload C --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_C.json 2> gpurun_out/r2_bench_C.err; tail -c 1500 gpurun_out/r2_bench_C.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_benchB.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scanw_kernel -s 4 -c 1 -f -o gpurun_out/r2_scanw python bench.py --steps 2 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/r2_ncu_scanw.log 2>&1
ls -la gpurun_out
