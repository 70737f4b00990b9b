#!/bin/sh
# round 2, call i (2 GPUs): launch-round overlap -- tests, then the sharded bench at N=2 and the N=1 line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.txt 2>&1 || { tail -8 gpurun_out/r2i_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2i_smoke.txt
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -8
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2i_pytest_2gpu.txt; cat gpurun_out/r2i_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 \
   2>gpurun_out/r2i_bench2.err | tail -1 > gpurun_out/r2i_bench2.json
tail -3 gpurun_out/r2i_bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2i_bench2.json'))
print('N', d['n_gpus'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4), d['roofline']['kernel_ms_per_step'])
print('gather_verified', d.get('gather_verified'), 'parity', d['parity']['identical'], '/', d['parity']['frames_checked'], 'weak', d['weak']['value'])
for o in d.get('other_configs', []):
    print(' ', o['workload'], round(o['value']), 'frac', round(o['roofline']['frac'], 4), o.get('gather_verified'), o['parity']['identical'], '/', o['parity']['frames_checked'])
PY
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2i_bench1.json
python -c "
import json; d = json.load(open('gpurun_out/r2i_bench1.json')); print('N1 fps', round(d['value']), 'frac', round(d['roofline']['frac'], 4))
for o in d.get('other_configs', []): print(' ', o['workload'][:50], round(o['value']), o['parity']['identical'], '/', o['parity']['frames_checked'])"
