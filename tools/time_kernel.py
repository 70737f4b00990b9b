#!/usr/bin/env python3
"""Kernel-time only (no result checks): for timing experiments whose output may be garbage (tools/ builds).
usage: time_kernel.py [frames] [kind] -> prints k_encode_chunks / k_layout / k_stitch ms per pass"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ec504_imageencoder_b200 import M1Encoder, MODE_FULL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
enc = M1Encoder(1920, 1080, 3, MODE_FULL, 12, max_frames=n)
rgb = enc.synth_rgb(12345, 0, n, kind)
res = enc.alloc_outputs(n)
enc.enable_timing(True)
for _ in range(3):
    enc.encode_device(rgb, res=res, check=False)
torch.cuda.synchronize()
enc.kernel_times()
steps = 10
for _ in range(steps):
    enc.encode_device(rgb, res=res, check=False)
torch.cuda.synchronize()
ms, cnt = enc.kernel_times()
print("enc_ms %.4f layout_ms %.4f stitch_ms %.4f" % tuple(m / steps for m in ms))
