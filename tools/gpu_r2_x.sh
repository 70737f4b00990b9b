#!/bin/sh
# round 2, call x: A/B of prebuilt library variants (build_variants/libm1cu_<name>.so), kernel time on default content / noise / grey
mkdir -p gpurun_out
{
for v in "$@" "$@"; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  echo "== $v: natural $(timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)   noise: $(timeout 120 python tools/time_kernel.py 300 1 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2x_variants_ab.txt
