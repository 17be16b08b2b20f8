set -x
mkdir -p gpurun_out
WORKLOADS="${WL:-B C}" bash tests/micro/ab.sh ${STEPS:-20} "$@" 2>&1 | grep -v "^+" | tee gpurun_out/ab_result.txt
