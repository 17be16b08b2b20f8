set -x
mkdir -p gpurun_out
WORKLOADS="${WL:-B}" bash tests/micro/ab.sh ${STEPS:-10} "$@" 2>&1 | grep -v "^+" | tee gpurun_out/ab_result.txt
