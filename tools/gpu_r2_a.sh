#!/bin/sh
# round 2, first GPU call: parity suite, bench, content sweep, phase split with the profiling build
mkdir -p gpurun_out
# a tiny encode first: stop at once if the kernel faults
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.txt 2>&1 || { tail -5 gpurun_out/r2a_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2a_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.txt; cat gpurun_out/r2a_pytest.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2a_bench.err | tail -1 > gpurun_out/r2a_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2a_bench.json'))
print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4),
      d['roofline']['kernel_ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks'])
PY
python tools/content_sweep.py r2a 2>&1 | tail -8
cp build_variants/libm1cu_exp.so ec504_imageencoder_b200/libm1cu.so
for s in 0 2 1; do
  M1_DEBUG_SKIP=$s python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip', $s, 'enc_ms_per_step', round(d['roofline']['kernel_ms_per_step']['k_encode_chunks'],3))"
done | tee gpurun_out/r2a_phase_split.txt
