#!/bin/sh
# ncu captures of the encode kernel with one half switched off (profiling knob M1_DEBUG_SKIP)
for s in 1 2; do
  M1_DEBUG_SKIP=$s python bench.py --steps 1 --warmup 3 --frames 40 --no-cpu-baseline > gpurun_out/plain_s$s.log 2>&1 && \
  M1_DEBUG_SKIP=$s ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/prof_skip$s \
      python bench.py --steps 1 --warmup 3 --frames 40 --no-cpu-baseline > gpurun_out/ncu_s$s.log 2>&1
  tail -1 gpurun_out/ncu_s$s.log
done
