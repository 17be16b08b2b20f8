set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scanw_kernel -s 4 -c 1 -f -o gpurun_out/scanw_r2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_scanw.log 2>&1
tail -5 gpurun_out/ncu_scanw.log
ls -la gpurun_out/*.ncu-rep
