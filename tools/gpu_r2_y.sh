#!/bin/sh
# round 2, call y: scattered content (a quarter of the 8x8 pixel tiles are noise) -- parity of the new tests, then the flat-block
# shortcut on / off on it (the eight-lanes-per-block path is what runs there)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.txt 2>&1 || { tail -8 gpurun_out/r2y_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2y_smoke.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2y_pytest_kernels.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2y_scattered_ab.txt
import torch
from ec504_imageencoder_b200 import M1Encoder
for (W, H, n, q, kind) in ((1920, 1080, 300, 12, 4), (1920, 1080, 300, 5, 4), (1920, 1080, 300, 12, 0), (1920, 1080, 300, 12, 1)):
    for nfs in (True, False, True, False):
        enc = M1Encoder(W, H, 3, 0, q, max_frames=n, no_flat_skip=nfs)
        rgb = enc.synth_rgb(12345, 0, n, kind); res = enc.alloc_outputs(n); enc.enable_timing(True)
        for _ in range(3): enc.encode_device(rgb, res=res, check=False)
        enc.check(); enc.kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): enc.encode_device(rgb, res=res, check=False)
        e1.record(); torch.cuda.synchronize(); enc.check()
        ms, _ = enc.kernel_times()
        print(W, H, 'q', q, 'kind', kind, 'flat_skip', not nfs, 'enc_ms', round(ms[0] / 10, 4), 'step_ms', round(e0.elapsed_time(e1) / 10, 4), 'fps', round(n / (e0.elapsed_time(e1) / 10) * 1e3))
        enc.close(); del rgb, res
PY
