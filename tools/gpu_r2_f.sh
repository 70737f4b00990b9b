#!/bin/sh
# round 2, call f (8 GPUs): the sharded bench exactly as the driver launches it (configs[3] split + configs[4] + weak), then N=4
mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 \
   2>gpurun_out/r2f_bench$n.err | tail -1 > gpurun_out/r2f_bench$n.json
tail -3 gpurun_out/r2f_bench$n.err
python - $n <<'PY'
import json, sys
d = json.load(open('gpurun_out/r2f_bench%s.json' % sys.argv[1]))
print('N', d['n_gpus'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'scaling', d['scaling'], 'frac', round(d['roofline']['frac'], 4), d['clocks'])
print('gather_verified', d.get('gather_verified'), d.get('gather'))
print('parity', d['parity']['identical'], '/', d['parity']['frames_checked'])
e = d['e2e']; print('e2e', round(e['value']), 'plain_h2d', round(e['plain_h2d_copy_gbs'], 1), 'aggregate', round(e['plain_h2d_aggregate_gbs'], 1), 'e2e_input_gbs', round(e['e2e_input_gbs'], 1), 'wc', e['write_combined_input'])
print('weak', d.get('weak', {}).get('value'), d.get('weak', {}).get('gather_verified'))
for o in d.get('other_configs', []):
    print(' ', o['workload'], round(o['value']), 'frac', round(o['roofline']['frac'], 4), o.get('gather_verified'), o['parity']['identical'], '/', o['parity']['frames_checked'])
PY
done
