#!/bin/sh
# resident CTAs per SM vs throughput: pad the dynamic shared memory (27.9 KB per CTA at 15-MB chunks)
# 0 -> 6 CTAs (register limit), 18000 -> 4, 30000 -> 3, 50000 -> 2, 90000 -> 1
for pad in 0 10000 18000 30000 50000 90000; do
  M1_PAD_SMEM=$pad python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pad', $pad, 'fps', round(d['value']), 'enc_ms', round(d['roofline']['kernel_ms_per_step']['k_encode_chunks'],3))"
done
