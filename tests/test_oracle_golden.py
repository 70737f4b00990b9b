"""The oracle port (oracle/m1_oracle.c) against the committed golden vectors, which were produced
by the UNMODIFIED reference (tests/golden/make_golden.py).  Runs anywhere, no GPU, no reference."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "kat.json")) as f:
        return json.load(f)


def test_qmatrix(port, kat):
    for q, m in kat["qmatrix"].items():
        assert port.qmatrix(int(q)).tolist() == m, q
    assert port.qmatrix(12)[:8].tolist() == [33, 67, 79, 92, 108, 113, 121, 142]


def test_ac_table(port, kat):
    seen = 0
    for r in range(32):
        for a in range(41):
            want = kat["ac_table"].get(f"{r},{a}", "")
            assert port.ac_table_entry(r, a) == want, (r, a)
            seen += bool(want)
    assert seen == 110
    assert len(kat["ac_table"]["16,1"]) == 15      # the reference's (run 16, level 2) code is 15 bits


def test_block_bits(port, kat):
    n_bad = 0
    for b in kat["blocks"]:
        zz = np.array(b["zz"], np.int32)
        if b["bits"] is None:
            n_bad += 1
            with pytest.raises(ValueError):
                port.block_bits(zz, b["luma"])
        else:
            assert port.block_bits(zz, b["luma"]) == b["bits"], b
    assert len(kat["blocks"]) == 800 and n_bad < 100


def test_survey_kats(port):
    """Known answers quoted in SURVEY.md section 8(c)."""
    def zz(*pairs):
        a = np.zeros(64, np.int32)
        for i, v in enumerate(pairs):
            a[i] = v
        return a
    assert port.block_bits(zz(5, 0, 3, 2, 0, 0, -1), 1) == "101101000011010"
    assert port.block_bits(zz(0, 0, 3), 1) == "1000010010110"
    assert port.block_bits(zz(0, 0, 3), 0) == "000010010110"
    assert port.block_bits(zz(-3), 1) == "010110"
    assert port.block_bits(zz(), 1) == "10010"
    assert port.block_bits(zz(), 0) == "0010"
    assert port.block_bits(zz(61, 0, -2, 0, 0, 1), 1) == "111101111010010101110"
    assert port.block_bits(zz(255), 1) == "11111101111111110"
    assert port.block_bits(zz(256), 1) == "00010"
    for L, want in ((2, "00101"), (-2, "00101"), (40, "00000100000000101000"), (-40, "00000100000011011000"),
                    (-128, "0000010000001000000010000000")):
        assert port.block_bits(zz(0, L), 1) == "100" + want + "10", L      # one zero before it: z = 1
    a = np.zeros(64, np.int32)
    a[34] = -1                                        # 33 zeros before it when the DC is absent... run 34 -> r = 33
    assert port.block_bits(a, 1) == "100" + "000001" + "100001" + "11111111" + "10"


def test_dct(port, kat):
    for d in kat["dct"]:
        assert port.fdct8x8(np.array(d["in"], np.uint8)).tolist() == d["out"]


def test_colour_all_inputs(port, kat):
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    rgb = np.stack([r.ravel(), g.ravel(), b.ravel()], 1)
    planes = port.rgb_to_ycbcr(rgb)
    R, G, B = (rgb[:, i].astype(np.int64) for i in range(3))
    exact = {"Y": (299 * R + 587 * G + 114 * B) // 1000,
             "Cb": (128_000_000 - 168_736 * R - 331_264 * G + 500_000 * B) // 1_000_000,
             "Cr": (128_000_000 + 500_000 * R - 418_688 * G - 81_312 * B) // 1_000_000}
    for name, got in zip(("Y", "Cb", "Cr"), planes):
        want = kat["colour"][name]
        assert hashlib.sha256(got.tobytes()).hexdigest() == want["sha256"], name
        diff = np.nonzero(got.astype(np.int64) != exact[name])[0]
        assert diff.size == want["n_below_exact_floor"]
        assert diff[:64].tolist() == want["first_below"]
    # grey 128 -> Y 127, grey 255 -> Y 255, Cr(255,255,255) = 127
    y, cb, cr = port.rgb_to_ycbcr(np.array([[128, 128, 128], [255, 255, 255]], np.uint8))
    assert y.tolist() == [127, 255] and cr[1] == 127


def test_headers(port, kat):
    assert port.file_prologue().hex() == kat["file_prologue"]
    assert kat["file_prologue"] == "000001ba2100010001c33367000001bb0009c333670021ffe0e0e6"
    for p in kat["frame_prefix"]:
        assert port.frame_prefix(p["i"], p["W"], p["H"], p["mode"], p["payload"]).hex() == p["hex"], p


def test_pictures(port, kat):
    for p in kat["pictures"]:
        img = port.synth_rgb(p["seed"], p["frame"], p["W"], p["H"], p["kind"])
        assert hashlib.sha256(img.tobytes()).hexdigest() == p["rgb_sha256"], "synthetic generator changed"
        pay, lev = port.encode_picture(img, p["q"], p["mode"], want_levels=True)
        assert len(pay) == p["payload_bytes"], p
        assert hashlib.sha256(pay).hexdigest() == p["payload_sha256"], p
        assert hashlib.sha256(lev.tobytes()).hexdigest() == p["levels_sha256"], p
        if "payload_hex" in p:
            assert pay.hex() == p["payload_hex"]


def refcompat_expected_stream(port, images, frame_image, full_hw=(600, 400)):
    """prologue + per frame (44-byte prefix with the REAL picture size, payload of the cropped
    picture, 4 trailer bytes) -- what mpeg_encode_procedure writes (include/encoder.h:196-458)."""
    out = bytearray(port.file_prologue())
    payload_of = {}
    for i, idx in enumerate(frame_image):
        idx = int(idx)
        if idx not in payload_of:
            payload_of[idx] = port.encode_picture(images[idx], 12, 1)
        pay = payload_of[idx]
        out += port.frame_prefix(i, full_hw[1], full_hw[0], 1, len(pay)) + pay + b"\x00\x00\x01\xb7"
    return bytes(out)


def masked_equal(ours: bytes, ref: bytes, n_frames: int, sizes):
    """The reference writes 4 uninitialised bytes after every frame (include/encoder.h:456-458);
    they are the only bytes excluded from the comparison."""
    if len(ours) != len(ref):
        return False
    a, b = bytearray(ours), bytearray(ref)
    pos = 27
    for n in sizes:
        pos += 44 + n
        a[pos:pos + 4] = b[pos:pos + 4] = b"\0\0\0\0"
        pos += 4
    return a == b


def test_refcompat_stream_equals_reference_binary(port):
    """REF_COMPAT end to end: the byte stream of the reference's own ./encoder on images.zip."""
    z = np.load(os.path.join(GOLD, "refcompat_inputs.npz"))
    ref_video = open(os.path.join(GOLD, "refcompat_video.mpeg"), "rb").read()
    images, frame_image = z["images"], z["frame_image"]
    assert images.shape == (3, 144, 400, 3) and len(frame_image) == 30 and len(ref_video) == 18187
    ours = refcompat_expected_stream(port, images, frame_image)
    sizes = [len(port.encode_picture(images[int(i)], 12, 1)) for i in frame_image]
    assert masked_equal(ours, ref_video, 30, sizes)
    # header quirk: 400x600 is written as 144x88 (uint8 truncation, include/encoder.h:186-187)
    assert ref_video[27 + 16:27 + 28].hex() == "000001b309005814ffffe018"
