#!/bin/sh
# round 2, call v: 8 resident CTAs per SM (64 registers, spills in the rare DCT path) against 7 (72 registers), now that most blocks skip the DCT
mkdir -p gpurun_out
{
for v in cta7 cta8 cta7 cta8; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  echo "== $v: natural $(timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)   noise: $(timeout 120 python tools/time_kernel.py 300 1 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2v_cta8_ab.txt
