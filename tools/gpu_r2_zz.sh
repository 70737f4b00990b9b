#!/bin/sh
# round 2, call zz: ncu --set full captures of k_encode_chunks on noise and on scattered content (40 frames each), final kernel
mkdir -p gpurun_out
for k in 1 4; do
  python tools/time_kernel.py 40 $k > gpurun_out/r2zz_plain_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/r2zz_prof_kind$k python tools/time_kernel.py 40 $k > gpurun_out/r2zz_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2zz_ncu_$k.log
done
