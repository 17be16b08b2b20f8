set -x
mkdir -p gpurun_out
cp ivfadc.jl_b200/libivfadc_cuda.so /tmp/lib_product.so
for v in "$@"; do IVFADC_LIB=$v timeout -s KILL 100 python tests/micro/try_variant.py 2>&1 | tail -2 | cut -c 1-200; done
cp /tmp/lib_product.so ivfadc.jl_b200/libivfadc_cuda.so
WORKLOADS="${WL:-B C}" bash tests/micro/ab.sh ${STEPS:-20} "$@" 2>&1 | grep -v "^+" | tee gpurun_out/ab_result.txt
