#!/bin/sh
# Runs on an N-GPU box ($1 = N): the bench at N GPUs only.
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/scale_${N}gpu.err | tail -1 > gpurun_out/scale_${N}gpu.json
python -c "
import json; d = json.load(open('gpurun_out/scale_${N}gpu.json'))
print('N', d['n_gpus'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']))"
