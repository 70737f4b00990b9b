#!/bin/sh
# Runs on the GPU box: bench each prebuilt library variant in build_variants/ (A/B kernel experiments).
for v in "$@"; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  for i in 1 2; do
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/var_$v.json
    python - "$v" <<'PY'
import json, sys
d = json.load(open('gpurun_out/var_%s.json' % sys.argv[1]))
print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), d['roofline']['kernel_ms_per_step'])
PY
  done
done
