#!/bin/sh
# Profiling build of the CUDA library: the same sources with -DM1_EXPERIMENTS (tools/experiments/m1x_env.h knobs).
cd "$(dirname "$0")/.." || exit 1
mkdir -p build_variants
exec nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DM1_EXPERIMENTS "$@" \
    -o build_variants/libm1cu_exp.so ec504_imageencoder_b200/csrc/m1cu_kernels.cu ec504_imageencoder_b200/csrc/m1cu_api.cu
