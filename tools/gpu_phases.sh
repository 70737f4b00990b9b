#!/bin/sh
# how long do the colour phase and the block phases take on their own? (profiling knob M1_DEBUG_SKIP)
for s in 0 2 1; do
  M1_DEBUG_SKIP=$s python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip', $s, 'enc_ms_per_step', round(d['roofline']['kernel_ms_per_step']['k_encode_chunks'],3))"
done
