#!/bin/sh
# Non-product builds of the CUDA library (same sources), into build_variants/ (git-ignored, travels with gpurun):
#   libm1cu_exp.so        -DM1_EXPERIMENTS: profiling knobs of tools/experiments/m1x_env.h (M1_DEBUG_SKIP, M1_PAD_SMEM)
#   libm1cu_intcolour.so  -DM1_COLOUR_SPLIT=0: integer colour path + fix-up queue   (tests/test_gpu_variants.py)
#   libm1cu_mixcolour.so  -DM1_COLOUR_SPLIT=1: row 0 integer, row 1 double chain
# usage: tools/build_experiments.sh [exp|intcolour|mixcolour ...]   (default: all three); extra nvcc flags via NVCC_EXTRA
cd "$(dirname "$0")/.." || exit 1
mkdir -p build_variants
[ $# -eq 0 ] && set -- exp intcolour mixcolour
for v in "$@"; do
  case $v in
    exp) D="-DM1_EXPERIMENTS" ;;
    intcolour) D="-DM1_COLOUR_SPLIT=0" ;;
    mixcolour) D="-DM1_COLOUR_SPLIT=1" ;;
    *) echo "unknown variant $v"; exit 2 ;;
  esac
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $D $NVCC_EXTRA \
      -o build_variants/libm1cu_$v.so ec504_imageencoder_b200/csrc/m1cu_kernels.cu ec504_imageencoder_b200/csrc/m1cu_api.cu || exit 1
done
