#!/bin/sh
# Runs on the GPU box: one full ncu capture of k_encode_chunks (tag = $1), after a plain run of the same command.
python bench.py --steps 1 --warmup 3 --frames 40 --no-cpu-baseline > gpurun_out/plain_$1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/prof_$1 \
    python bench.py --steps 1 --warmup 3 --frames 40 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
tail -1 gpurun_out/ncu_$1.log | cut -c1-200
