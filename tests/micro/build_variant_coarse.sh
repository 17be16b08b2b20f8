#!/bin/bash
# bring-up helper: build libivfadc_cuda with extra -D flags for coarse.cu into ivfadc.jl_b200/build/variants/<name>.so
set -e
cd "$(dirname "$0")/../../ivfadc.jl_b200"
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../include -Icsrc "$@" -c csrc/coarse.cu -o build/variants/coarse_$name.o
nvcc -shared -o build/variants/$name.so build/api.o build/variants/coarse_$name.o build/scan.o build/encode.o build/lists.o -gencode arch=compute_100a,code=sm_100a
echo built build/variants/$name.so
