#!/usr/bin/env python3
"""1080p device-resident throughput by content kind (natural / noise / grey / r == g) and quality: the
integer colour path's fix-up queue makes k_encode_chunks content dependent, so the worst cases are
reported beside the default workload.  Same timing rules as bench.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ec504_imageencoder_b200 import M1Encoder, MODE_FULL  # noqa: E402

KINDS = {0: "natural", 1: "noise", 2: "grey", 3: "r==g"}


def run(W, H, n, q, kind, steps=10, warm=3):
    enc = M1Encoder(W, H, 3, MODE_FULL, q, max_frames=n)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        rgb = enc.synth_rgb(12345, 0, n, kind)
        res = enc.alloc_outputs(n)
        enc.enable_timing(True)
        for _ in range(warm):
            enc.encode_device(rgb, res=res, check=False)
        enc.check()
        enc.kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            enc.encode_device(rgb, res=res, check=False)
        e1.record(st)
        torch.cuda.synchronize()
        enc.check()
        kms, kn = enc.kernel_times()
        ms = e0.elapsed_time(e1) / steps
        payload = int(res.frame_bytes.to(torch.int64).sum().item()) / n
    fps = n / (ms * 1e-3)
    alg = 3 * W * H + payload + 4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    enc.close()
    return {"width": W, "height": H, "frames": n, "quality": q, "content": KINDS[kind], "ms_per_pass": ms,
            "frames_per_s": fps, "payload_bytes_per_frame": payload, "encode_kernel_ms": kms[0] / steps,
            "encode_kernel_roofline_frac": alg * n / (kms[0] / steps * 1e-3) / 1e9 / peak}


if __name__ == "__main__":
    out = []
    for cfg in [(1920, 1080, 300, 12, 0), (1920, 1080, 300, 12, 1), (1920, 1080, 300, 12, 2), (1920, 1080, 300, 12, 3),
                (1920, 1080, 300, 50, 1), (1920, 1080, 300, 50, 0)]:
        r = run(*cfg)
        out.append(r)
        print(json.dumps(r), flush=True)
    tag = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    with open(os.path.join(ROOT, "gpurun_out", f"content_{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)
