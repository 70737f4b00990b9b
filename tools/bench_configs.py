#!/usr/bin/env python3
"""Device-resident throughput of the other BASELINE.json configurations on ONE GPU (bench.py's JSON
line covers configs[1]; these are reported for context, same timing rules: inputs resident in HBM,
CUDA events on the launching stream, >= 3 warm-ups, inputs larger than L2).

  configs[0] 352x240   30 frames   q 12
  configs[2] 3840x2160 frames      q 5 / 12 / 50      (quality sweep; 60 frames = 1.5 GB per pass)
  configs[4] 7680x4320 15 frames   q 12               (the per-GPU share of the 8-GPU 8K stress)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ec504_imageencoder_b200 import M1Encoder, MODE_FULL, SYNTH_NATURAL, SYNTH_NOISE  # noqa: E402


def run(W, H, n, q, kind, steps=10, warm=3):
    enc = M1Encoder(W, H, 3, MODE_FULL, q, max_frames=n)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        rgb = enc.synth_rgb(12345, 0, n, kind)
        res = enc.alloc_outputs(n)
        for _ in range(warm):
            enc.encode_device(rgb, res=res, check=False)
        enc.check()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            enc.encode_device(rgb, res=res, check=False)
        e1.record(st)
        torch.cuda.synchronize()
        enc.check()
        ms = e0.elapsed_time(e1) / steps
        payload = int(res.frame_bytes.to(torch.int64).sum().item()) / n
    fps = n / (ms * 1e-3)
    alg = 3 * W * H + payload + 4
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    enc.close()
    return {"width": W, "height": H, "frames": n, "quality": q, "content": "noise" if kind else "natural",
            "ms_per_pass": ms, "frames_per_s": fps, "megapixels_per_s": fps * W * H / 1e6,
            "payload_bytes_per_frame": payload, "hbm_roofline_frac_whole_pass": alg * fps / 1e9 / peak}


def main():
    out = []
    for cfg in [(352, 240, 30, 12, SYNTH_NATURAL), (352, 240, 3000, 12, SYNTH_NATURAL),
                (1920, 1080, 300, 12, SYNTH_NATURAL), (1920, 1080, 300, 12, SYNTH_NOISE), (1920, 1080, 300, 50, SYNTH_NOISE),
                (3840, 2160, 60, 5, SYNTH_NATURAL), (3840, 2160, 60, 12, SYNTH_NATURAL), (3840, 2160, 60, 50, SYNTH_NATURAL),
                (7680, 4320, 15, 12, SYNTH_NATURAL)]:
        r = run(*cfg)
        out.append(r)
        print(json.dumps(r), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
