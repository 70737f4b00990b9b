#!/bin/sh
# A/B: fused kernel (M1_NO_WS=1) vs warp-specialised kernel
for ws in 1 0; do
  M1_NO_WS=$ws python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('M1_NO_WS', $ws, 'fps', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'enc_ms', round(d['roofline']['kernel_ms_per_step']['k_encode_chunks'],3))"
done
