#!/bin/sh
# chunk-size sweep (tuning knob M1_CHUNK_MBS); prints fps per setting
for c in 32 30 24 20 16 15 12 10 8 6 4; do
  M1_CHUNK_MBS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk_mbs', $c, 'fps', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'enc_ms', round(d['roofline']['kernel_ms_per_step']['k_encode_chunks'],3), 'stitch', round(d['roofline']['kernel_ms_per_step']['k_stitch'],3))"
done
