#!/bin/sh
# round 2, call j: final 1-GPU test log with the final tree + the host topology behind the multi-GPU e2e ceiling
mkdir -p gpurun_out
{ lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)|thread"; free -g | head -2; nvidia-smi topo -m 2>/dev/null | head -14; } > gpurun_out/r2j_host_topology.txt 2>&1; cat gpurun_out/r2j_host_topology.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2j_pytest_1gpu.txt; cat gpurun_out/r2j_pytest_1gpu.txt
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2j_bench.err | tail -1 > gpurun_out/r2j_bench.json; tail -2 gpurun_out/r2j_bench.err
python -c "
import json; d = json.load(open('gpurun_out/r2j_bench.json')); print('fps', round(d['value']), 'frac', round(d['roofline']['frac'], 4), 'e2e', round(d['e2e']['value']), d['clocks'], d['parity']['identical'], '/', d['parity']['frames_checked'], list(d.keys()))"
python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
