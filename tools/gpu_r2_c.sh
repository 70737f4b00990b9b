#!/bin/sh
# round 2, call c: A/B of the colour split variants, then the whole GPU suite and the full bench line on the default.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.txt 2>&1 || { tail -5 gpurun_out/r2c_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2c_smoke.txt
for v in split0 split1 split2; do
  cp build_variants/libm1cu_$v.so ec504_imageencoder_b200/libm1cu.so
  echo "== $v"
  python tools/content_sweep.py r2c_$v 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(r['content'], 'q', r['quality'], 'fps', round(r['frames_per_s']), 'enc_ms', round(r['encode_kernel_ms'], 3), 'frac', round(r['encode_kernel_roofline_frac'], 4))
"
done 2>&1 | tee gpurun_out/r2c_variants.txt
cp build_variants/libm1cu_split1.so ec504_imageencoder_b200/libm1cu.so
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.txt; cat gpurun_out/r2c_pytest.txt
python bench.py --steps 10 --warmup 3 2>gpurun_out/r2c_bench.err | tail -1 > gpurun_out/r2c_bench.json
tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_bench.json'))
print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4),
      d['roofline']['kernel_ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks'], d['parity'])
for o in d.get('other_configs', []):
    print(' ', o['workload'], round(o['value']), 'frac', round(o['roofline']['frac'], 4), o['parity']['identical'], '/', o['parity']['frames_checked'])
print(d.get('cpu_baseline'))
PY
python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > gpurun_out/r2c_reference.json; cut -c1-400 gpurun_out/r2c_reference.json
