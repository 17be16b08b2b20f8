# round 2, step B on one B200: launch lists of the library's kernels on B and D, ncu captures of the D-shape scan and the coarse kernels
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ivf -c 400 --csv --log-file gpurun_out/r2b_launches_B.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2b_launches_B.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ivf -s 300 -c 300 --csv --log-file gpurun_out/r2b_launches_D.csv python bench.py --workload D --steps 3 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2b_launches_D.log 2>&1
tail -2 gpurun_out/r2b_launches_D.log | cut -c 1-600
IVFADC_BENCH_N=12500000 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/r2b_scan_legacy_D python bench.py --workload D --steps 2 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2b_ncu_D.log 2>&1
tail -3 gpurun_out/r2b_ncu_D.log | cut -c 1-300
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:coarse -s 40 -c 4 -f -o gpurun_out/r2b_coarse_B python bench.py --steps 2 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2b_ncu_coarse.log 2>&1
tail -3 gpurun_out/r2b_ncu_coarse.log | cut -c 1-300
ls -la gpurun_out | tail -12
