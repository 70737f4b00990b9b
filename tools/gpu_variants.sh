#!/bin/sh
# Runs on the GPU box: bench each prebuilt library variant in build_variants/ (A/B kernel experiments).
# Arguments: name[:ENV=val[:ENV=val...]] -- libm1cu_<name>.so, optional environment knobs for the run.
for spec in "$@"; do
  v=${spec%%:*}
  envs=$(echo "$spec" | tr ':' ' ' | cut -s -d' ' -f2-)
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  for i in 1 2; do
    env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/var_$v.json
    python - "$spec" "$v" <<'PY'
import json, sys
d = json.load(open('gpurun_out/var_%s.json' % sys.argv[2]))
print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), {k: round(v, 4) for k, v in d['roofline']['kernel_ms_per_step'].items()})
PY
  done
done
