#!/bin/sh
# Runs on the GPU box: parity tests, a short bench, and (optional, $1=ncu) a full ncu capture of k_encode_chunks.
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/quick.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/quick.json'))
print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4),
      d['roofline']['kernel_ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks'])
PY
if [ "$1" = "ncu" ]; then
  python bench.py --steps 1 --warmup 3 --frames 40 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/prof_$2 \
      python bench.py --steps 1 --warmup 3 --frames 40 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
  tail -1 gpurun_out/ncu2.log
fi
