#!/bin/sh
# Round-end evidence, part 1: tests, bench (both arms), ncu launch list of the bench command.
# (tools/gpu_ncu_only.sh <tag> takes the full capture of the top kernel in a separate call.)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_pytest.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference_arm.json 2> gpurun_out/final_reference_arm.err
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -1 gpurun_out/final_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'], d['clocks'])"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c1-300
