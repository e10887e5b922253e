#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..." : builds lib/variants/libh2aes_NAME.so with extra nvcc flags
# (tuning experiments; select at run time with H2AES_LIB=<path>)
set -e
cd "$(dirname "$0")"
name=$1; shift
out=../lib/variants; mkdir -p $out
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" \
     -shared -cudart static -ldl -o $out/libh2aes_$name.so api.cu witness.cu lookup.cu gprod.cu comm.cu ntt.cu quotient.cu msm.cu
echo built $out/libh2aes_$name.so
