#!/bin/sh
# round 2, call b: integer-colour micro-benchmark, then ncu --set full of k_encode_chunks (product build), then of the
# colour phase alone (profiling build, M1_DEBUG_SKIP=2).  bench.py at 40 frames keeps the ncu replays short.
mkdir -p gpurun_out
./tools/ubench/int_colour_issue > gpurun_out/r2b_ubench_int_colour.txt 2>&1; cat gpurun_out/r2b_ubench_int_colour.txt
B="python bench.py --steps 1 --warmup 3 --frames 40 --no-cpu-baseline --no-other-configs"
$B > gpurun_out/r2b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/r2b_prof_full $B > gpurun_out/r2b_ncu_full.log 2>&1
tail -2 gpurun_out/r2b_ncu_full.log
cp build_variants/libm1cu_exp.so ec504_imageencoder_b200/libm1cu.so
M1_DEBUG_SKIP=2 $B > gpurun_out/r2b_plain2.log 2>&1 && \
M1_DEBUG_SKIP=2 ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/r2b_prof_colour $B > gpurun_out/r2b_ncu_colour.log 2>&1
tail -2 gpurun_out/r2b_ncu_colour.log
