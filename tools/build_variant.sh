#!/bin/bash
# builds an alternative libnanogi_gpu.so into build/<tag>.so with extra -D flags (A/B sessions select it with NGI_GPU_LIB)
# usage: tools/build_variant.sh <tag> [-DNAME=VALUE ...]
set -e
TAG=$1; shift
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --shared "$@" \
    -o build/$TAG.so nanogi_b200/csrc/ngi_gpu.cu
echo "built build/$TAG.so $*"
