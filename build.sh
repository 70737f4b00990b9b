#!/bin/sh
# Build libm1cu.so (CUDA kernels + C ABI) for sm_100a in-tree.  Extra args go to nvcc (e.g. -Xptxas -v).
cd "$(dirname "$0")" || exit 1
exec nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
    -o ec504_imageencoder_b200/libm1cu.so ec504_imageencoder_b200/csrc/m1cu_kernels.cu ec504_imageencoder_b200/csrc/m1cu_api.cu
