#!/bin/sh
# round 2, call s (N GPUs, N = $1): final sharded bench exactly as the driver launches it (configs[3] split + configs[4] + weak);
# at N = 2 the whole GPU suite first (peer-memory tests included)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = 2 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2s_pytest_2gpu.txt; cat gpurun_out/r2s_pytest_2gpu.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 \
   2>gpurun_out/r2s_bench$N.err | tail -1 > gpurun_out/r2s_bench$N.json
tail -3 gpurun_out/r2s_bench$N.err
python - $N <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2s_bench%s.json' % sys.argv[1]))
print('N', d['n_gpus'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'scaling', d['scaling'], 'frac', round(d['roofline']['frac'], 4), d['clocks'])
print('gather_verified', d.get('gather_verified'), d.get('gather'))
print('parity', d['parity']['identical'], '/', d['parity']['frames_checked'])
e = d['e2e']; print('e2e', round(e['value']), 'plain_h2d', round(e['plain_h2d_copy_gbs'], 1), 'aggregate', round(e['plain_h2d_aggregate_gbs'], 1), 'e2e_input_gbs', round(e['e2e_input_gbs'], 1), 'wc', e['write_combined_input'])
print('weak', d.get('weak', {}).get('value'), d.get('weak', {}).get('gather_verified'))
for o in d.get('other_configs', []):
    print(' ', o['workload'], round(o['value']), 'frac', round(o['roofline']['frac'], 4), o.get('gather_verified'), o['parity']['identical'], '/', o['parity']['frames_checked'])
PY
