#!/bin/sh
# Runs on a multi-GPU box: bench.py under torchrun at N GPUs ($1) for each prebuilt library variant ($2...).
N=$1; shift
for v in "$@"; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  for i in 1 2; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale_$v.err | tail -1 > gpurun_out/scale_$v.json
    python - "$v" <<'PY'
import json, sys
d = json.load(open('gpurun_out/scale_%s.json' % sys.argv[1]))
print(sys.argv[1], 'n', d['n_gpus'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), round(d['roofline']['kernel_ms_per_step']['k_encode_chunks'], 4), 'e2e', round(d['e2e']['value']))
PY
  done
done
