# ncu captures beyond the headline kernel: the 16 x 64 shape of the scan kernel on workload C, the coarse kernels on B
set -x
mkdir -p gpurun_out
tail -2 gpurun_out/r2_ncu_scanw_C.log | cut -c 1-200
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"coarse3_kernel|coarse3_rerank|coarse_redo_small|merge_cands" -s 12 -c 4 -f -o gpurun_out/r2_coarse_B python bench.py --steps 2 --warmup 3 --no-cpu-baseline --check 0 --extras none > gpurun_out/r2_ncu_coarse_B.log 2>&1
tail -2 gpurun_out/r2_ncu_coarse_B.log | cut -c 1-200
ls -la gpurun_out/*.ncu-rep
