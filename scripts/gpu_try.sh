set -x
cp ivfadc.jl_b200/libivfadc_cuda.so /tmp/lib_product.so
IVFADC_LIB=$1 timeout -s KILL 120 python tests/micro/try_variant.py 2>&1 | tail -5
rc=$?
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
cp /tmp/lib_product.so ivfadc.jl_b200/libivfadc_cuda.so
