# round 2, step A on one B200: new parity tests, smoke, validation bundle, bench with the extra shapes, scan phase profile
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -5 gpurun_out/r2a_smoke.log
timeout 300 python tests/golden/make_validation_bundle.py gpurun_out/validation > gpurun_out/r2a_bundle.log 2>&1; tail -3 gpurun_out/r2a_bundle.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_B.json 2> gpurun_out/r2a_bench_B.err; tail -c 6000 gpurun_out/r2a_bench_B.json; tail -20 gpurun_out/r2a_bench_B.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; tail -c 1500 gpurun_out/r2a_bench_ref.json; tail -5 gpurun_out/r2a_bench_ref.err
timeout 300 python tests/debug_timeline_w.py > gpurun_out/r2a_timeline_w.log 2>&1; tail -60 gpurun_out/r2a_timeline_w.log
