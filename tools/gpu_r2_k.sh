#!/bin/sh
# round 2, call k: paired last chunks -- parity (variants first), then A/B by the tuning switch, then the bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.txt 2>&1 || { tail -8 gpurun_out/r2k_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2k_smoke.txt
timeout 1200 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2k_pytest_kernels.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2k_pairing_ab.txt
import torch, json
from ec504_imageencoder_b200 import M1Encoder
for (W, H, n, q, kind) in ((1920, 1080, 300, 12, 0), (1920, 1080, 300, 12, 1), (352, 240, 3000, 12, 0), (3840, 2160, 100, 12, 0)):
    for nop in (True, False, True, False):
        enc = M1Encoder(W, H, 3, 0, q, max_frames=n, no_tail_pairing=nop)
        rgb = enc.synth_rgb(12345, 0, n, kind); res = enc.alloc_outputs(n); enc.enable_timing(True)
        for _ in range(3): enc.encode_device(rgb, res=res, check=False)
        enc.check(); enc.kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): enc.encode_device(rgb, res=res, check=False)
        e1.record(); torch.cuda.synchronize(); enc.check()
        ms, _ = enc.kernel_times()
        print(W, H, 'kind', kind, 'pairing', not nop, 'enc_ms', round(ms[0] / 10, 4), 'step_ms', round(e0.elapsed_time(e1) / 10, 4), 'fps', round(n / (e0.elapsed_time(e1) / 10) * 1e3))
        enc.close(); del rgb, res
PY
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2k_pytest_1gpu.txt
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2k_bench.err | tail -1 > gpurun_out/r2k_bench.json; tail -2 gpurun_out/r2k_bench.err
python -c "
import json; d = json.load(open('gpurun_out/r2k_bench.json')); print('fps', round(d['value']), 'frac', round(d['roofline']['frac'], 4), d['roofline']['kernel_ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity']['identical'], '/', d['parity']['frames_checked'])
for o in d.get('other_configs', []): print(' ', o['workload'][:50], round(o['value']), round(o['roofline']['frac'], 4), o['parity']['identical'], '/', o['parity']['frames_checked'])"
