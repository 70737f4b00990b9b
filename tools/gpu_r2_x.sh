#!/bin/sh
# round 2, call x: A/B of prebuilt library variants (build_variants/libm1cu_<name>.so): k_encode_chunks time on the default content,
# on noise and on scattered content (a quarter of the 8x8 pixel tiles noise)
mkdir -p gpurun_out
{
for v in "$@" "$@"; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  echo "== $v: natural $(timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1 | cut -d' ' -f1-2)   noise: $(timeout 120 python tools/time_kernel.py 300 1 2>&1 | tail -1 | cut -d' ' -f1-2)   scattered: $(timeout 120 python tools/time_kernel.py 300 4 2>&1 | tail -1 | cut -d' ' -f1-2)"
done
} 2>&1 | tee gpurun_out/r2x_variants_ab.txt
