#!/bin/sh
# A/B of one candidate ($1) against the current library, then the GPU parity suite with the candidate as the product library.
sh tools/gpu_variants.sh new $1 new $1 2>&1 | tee gpurun_out/ab_$1.log
cp build_variants/libm1cu_$1.so ec504_imageencoder_b200/libm1cu.so
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_$1.log
