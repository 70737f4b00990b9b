#!/bin/sh
# round 2, call d: what do the CTA barriers cost (garbage-output timing builds), and how does time depend on
# resident CTAs per SM (profiling build, shared-memory padding)?
mkdir -p gpurun_out
{
for v in split0 split2 nobar0 nobar2; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  echo "== $v: $(timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)"
done
cp build_variants/libm1cu_exp.so ec504_imageencoder_b200/libm1cu.so
for pad in 0 6000 12000 20000 30000 48000 85000; do
  echo "== exp pad $pad: $(M1_PAD_SMEM=$pad timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2d_barrier_occupancy.txt
