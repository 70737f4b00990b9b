#!/bin/sh
# round 2, call q: k_encode_chunks as resident CTAs pulling chunks from a device-side queue (tuning.work_queue) -- parity, A/B
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.txt 2>&1 || { tail -8 gpurun_out/r2q_smoke.txt; echo SMOKE_FAILED; exit 1; }
tail -1 gpurun_out/r2q_smoke.txt
timeout 1200 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2q_pytest_kernels.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2q_work_queue_ab.txt
import torch
from ec504_imageencoder_b200 import M1Encoder
for (W, H, n, q, kind) in ((1920, 1080, 300, 12, 0), (1920, 1080, 300, 12, 1), (352, 240, 3000, 12, 0), (3840, 2160, 75, 12, 0), (7680, 4320, 15, 12, 0)):
    for wq in (False, True, False, True):
        enc = M1Encoder(W, H, 3, 0, q, max_frames=n, work_queue=wq)
        rgb = enc.synth_rgb(12345, 0, n, kind); res = enc.alloc_outputs(n); enc.enable_timing(True)
        for _ in range(3): enc.encode_device(rgb, res=res, check=False)
        enc.check(); enc.kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): enc.encode_device(rgb, res=res, check=False)
        e1.record(); torch.cuda.synchronize(); enc.check()
        ms, _ = enc.kernel_times()
        print(W, H, 'kind', kind, 'work_queue', wq, 'enc_ms', round(ms[0] / 10, 4), 'step_ms', round(e0.elapsed_time(e1) / 10, 4), 'fps', round(n / (e0.elapsed_time(e1) / 10) * 1e3))
        enc.close(); del rgb, res
PY
