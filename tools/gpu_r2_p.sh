#!/bin/sh
# round 2, call p: k_stitch grid size (CTAs per launch; each warp loops over chunks) -- profiling build
mkdir -p gpurun_out
cp build_variants/libm1cu_exp.so ec504_imageencoder_b200/libm1cu.so
{
for n in 0 9472 4736 2368 1184 592; do
  echo "stitch CTAs $n: $(M1_STITCH_CTAS=$n timeout 120 python tools/time_kernel.py 300 0 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2p_stitch_grid.txt
