"""Decoder equivalence (north_star: "the output .mpeg must ... decode in FFmpeg/pl_mpeg exactly as the reference
output does"; reference README.md:140-145 names FFmpeg as its only test).  The decoder is FFmpeg through
cv2.VideoCapture, as SURVEY.md section 4 used it on the reference's file (30 frames of 144x88).

CPU tests decode (a) the reference binary's own output (tests/golden/refcompat_video.mpeg) and (b) the REF_COMPAT /
FULL streams of the oracle port (the very bytes the GPU path must produce, tests/test_gpu_parity.py) and compare
frame count, size and every pixel; GPU tests repeat it with the streams the CUDA path produced through the C
driver.  The FULL-mode result is pinned by tests/golden/decoder_goldens.json (tests/golden/make_decoder_golden.py)."""
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

os.environ.setdefault("OPENCV_FFMPEG_LOGLEVEL", "-8")      # the reference's syntax makes FFmpeg complain on every slice
cv2 = pytest.importorskip("cv2")

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
MODE_FULL, MODE_REF_COMPAT = 0, 1


def decode_with_ffmpeg(data: bytes):
    """All frames FFmpeg yields for the stream bytes (BGR uint8 arrays)."""
    with tempfile.NamedTemporaryFile(suffix=".mpeg", delete=False) as f:
        f.write(data)
        path = f.name
    try:
        cap = cv2.VideoCapture(path)
        assert cap.isOpened(), "FFmpeg cannot open the stream"
        frames = []
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            frames.append(fr.copy())
        cap.release()
    finally:
        os.unlink(path)
    return frames


def sif_frames(port):
    return np.stack([port.synth_rgb(12345, f, 352, 240, 0) for f in range(30)])


def fixture_frames():
    z = np.load(os.path.join(GOLD, "refcompat_inputs.npz"))
    images, frame_image = z["images"], z["frame_image"]
    frames = np.zeros((len(frame_image), 600, 400, 3), np.uint8)       # rows >= 144 are never read in REF_COMPAT
    for i, idx in enumerate(frame_image):
        frames[i, :144] = images[int(idx)]
    return frames


def _goldens():
    with open(os.path.join(GOLD, "decoder_goldens.json")) as f:
        return json.load(f)


def _summary(frames):
    return {"frames": len(frames), "height": int(frames[0].shape[0]), "width": int(frames[0].shape[1]),
            "sha256": hashlib.sha256(b"".join(f.tobytes() for f in frames)).hexdigest()}


def _same_pictures(a, b):
    assert len(a) == len(b) and len(a) > 0
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)


def test_reference_output_decodes_as_the_survey_saw_it():
    """FFmpeg opens the reference's awesome_video.mpeg as 30 frames of 144x88."""
    ref = decode_with_ffmpeg(open(os.path.join(GOLD, "refcompat_video.mpeg"), "rb").read())
    assert len(ref) == 30 and ref[0].shape == (88, 144, 3)
    assert _summary(ref) == _goldens()["reference_binary_output"]


def test_ref_compat_stream_decodes_exactly_like_the_reference_output(port):
    """Our REF_COMPAT stream differs from the reference's file only in the 4 trailer bytes per picture
    (uninitialised stack there, 00 00 01 b7 here): FFmpeg must not see a difference."""
    ref = decode_with_ffmpeg(open(os.path.join(GOLD, "refcompat_video.mpeg"), "rb").read())
    ours = decode_with_ffmpeg(port.encode_stream(fixture_frames(), 12, MODE_REF_COMPAT))
    _same_pictures(ref, ours)


def test_full_mode_stream_decode_is_pinned(port):
    """FULL mode (BASELINE configs[0]: 30 SIF frames): FFmpeg yields 30 frames of 352x240; the pixels are pinned."""
    got = decode_with_ffmpeg(port.encode_stream(sif_frames(port), 12, MODE_FULL))
    assert len(got) == 30 and got[0].shape == (240, 352, 3)
    assert _summary(got) == _goldens()["full_mode_sif_30"]


@pytest.mark.gpu
def test_gpu_streams_decode_like_the_reference(port):
    """The same two checks on the bytes the CUDA path wrote through the C driver (libencoder.so)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import hostlib
    ref = decode_with_ffmpeg(open(os.path.join(GOLD, "refcompat_video.mpeg"), "rb").read())
    _same_pictures(ref, decode_with_ffmpeg(hostlib.encode_frames_to_memory(fixture_frames(), 12, MODE_REF_COMPAT)))
    full = decode_with_ffmpeg(hostlib.encode_frames_to_memory(sif_frames(port), 12, MODE_FULL))
    assert _summary(full) == _goldens()["full_mode_sif_30"]
