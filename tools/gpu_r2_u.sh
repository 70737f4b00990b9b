#!/bin/sh
# round 2, call u: 1-GPU evidence on the final tree (instruction diet + flat-block shortcut) -- whole GPU suite, bench line, reference arm,
# ncu launch list and one --set full capture of k_encode_chunks (product build)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.txt 2>&1 || { tail -8 gpurun_out/r2u_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2u_smoke.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2u_pytest_1gpu.txt; cat gpurun_out/r2u_pytest_1gpu.txt
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2u_bench.err | tail -1 > gpurun_out/r2u_bench.json; tail -2 gpurun_out/r2u_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2u_bench.json'))
print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4), d['roofline']['kernel_ms_per_step'],
      'e2e', round(d['e2e']['value']), d['clocks'], d['parity']['identical'], '/', d['parity']['frames_checked'])
for o in d.get('other_configs', []):
    print(' ', o['workload'], round(o['value']), 'frac', round(o['roofline']['frac'], 4), o['parity']['identical'], '/', o['parity']['frames_checked'])
PY
python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/r2u_reference_arm.json; cut -c1-200 gpurun_out/r2u_reference_arm.json
B="python bench.py --steps 2 --warmup 3 --frames 40 --no-cpu-baseline --no-other-configs"
$B > gpurun_out/r2u_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 6 --csv --log-file gpurun_out/r2u_launches.csv $B > gpurun_out/r2u_ncu_launches.log 2>&1
$B > gpurun_out/r2u_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_encode -s 3 -c 1 -o gpurun_out/r2u_prof $B > gpurun_out/r2u_ncu_full.log 2>&1
tail -2 gpurun_out/r2u_ncu_full.log; grep -c k_ gpurun_out/r2u_launches.csv
