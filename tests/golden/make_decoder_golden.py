#!/usr/bin/env python3
"""Writes tests/golden/decoder_goldens.json: what FFmpeg (through cv2.VideoCapture, the decoder SURVEY.md section 4
names) makes of (a) the reference binary's own awesome_video.mpeg (tests/golden/refcompat_video.mpeg) and (b) the
FULL-mode stream of the SIF configuration (BASELINE configs[0], 30 synthetic 352x240 frames, quality 12) as the
oracle port produces it -- frame count, frame size and a SHA-256 over all decoded pixels.  The reference's
bitstream is not valid MPEG-1 video (README.md:144-145), so the pictures are garbage; the golden pins that the
garbage does not change.  Run from the repository root:  python tests/golden/make_decoder_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from test_decoder import decode_with_ffmpeg, sif_frames  # noqa: E402

import oracle  # noqa: E402


def summary(frames):
    return {"frames": len(frames), "height": int(frames[0].shape[0]), "width": int(frames[0].shape[1]),
            "sha256": hashlib.sha256(b"".join(f.tobytes() for f in frames)).hexdigest()}


if __name__ == "__main__":
    import cv2
    port = oracle.Port()
    out = {"decoder": "cv2.VideoCapture (FFmpeg backend), OpenCV " + cv2.__version__,
           "reference_binary_output": summary(decode_with_ffmpeg(open(os.path.join(HERE, "refcompat_video.mpeg"), "rb").read())),
           "full_mode_sif_30": summary(decode_with_ffmpeg(port.encode_stream(sif_frames(port), 12, oracle.MODE_FULL)))}
    with open(os.path.join(HERE, "decoder_goldens.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
