#!/bin/sh
# round 2, call e (2 GPUs): whole GPU suite incl. the peer-memory tests, then the sharded bench (configs[3] split) at N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2e_pytest_2gpu.txt; cat gpurun_out/r2e_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 \
   2>gpurun_out/r2e_bench2.err | tail -1 > gpurun_out/r2e_bench2.json
tail -5 gpurun_out/r2e_bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2e_bench2.json'))
print('N', d['n_gpus'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'scaling', d['scaling'], 'frac', round(d['roofline']['frac'], 4))
print('gather_verified', d.get('gather_verified'), d.get('gather'))
print('parity', d['parity'])
print('e2e', d['e2e'])
print('weak', d.get('weak'))
for o in d.get('other_configs', []):
    print(' ', o['workload'], round(o['value']), 'frac', round(o['roofline']['frac'], 4), o.get('gather_verified'), o['parity'])
print(d['config'])
PY
