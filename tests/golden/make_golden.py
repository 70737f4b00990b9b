#!/usr/bin/env python3
"""Regenerates the golden fixtures under tests/golden/ by RUNNING THE REFERENCE in the dev
container (needs /root/reference and oracle/_ref built by `make -C oracle ref`).

  refcompat_video.mpeg   the file the reference's own binary (main.c, unmodified) writes for images.zip
  refcompat_inputs.npz   stb_image-decoded RGB of the three distinct JPEGs (rows 0..143: all the
                         reference's 96x144 loops read), plus the frame order the binary used
  kat.json               function-level known answers produced by the reference's functions
                         (oracle/_ref/libm1ref.so): quantiser matrices, AC table, block bit strings,
                         DCT blocks, colour-conversion exceptions and digests, headers, picture payloads

The GPU box has no /root/reference; the tests there read only these files.
"""
import ctypes as C
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF_ROOT = "/root/reference"


def stb_decode(path):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libstb.so"))
    lib.stbi_load.restype = C.POINTER(C.c_ubyte)
    lib.stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    p = lib.stbi_load(path.encode(), C.byref(w), C.byref(h), C.byref(ch), 0)
    assert p, path
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, ch.value)).copy()
    lib.stbi_image_free(p)
    return a


def make_refcompat():
    tmp = tempfile.mkdtemp(prefix="m1golden_")
    try:
        os.makedirs(os.path.join(tmp, "images"))
        with zipfile.ZipFile(os.path.join(REF_ROOT, "images.zip")) as z:
            for n in z.namelist():
                if n.lower().endswith((".jpg", ".jpeg")):
                    with open(os.path.join(tmp, "images", os.path.basename(n)), "wb") as f:
                        f.write(z.read(n))
        os.makedirs(os.path.join(tmp, "bitstreams"), exist_ok=True)
        out = subprocess.run([oracle.ref_encoder_binary()], cwd=tmp, capture_output=True, text=True, errors="replace")
        assert out.returncode == 0, out.returncode
        order = re.findall(r"Loaded image: (\S+) \(Width", out.stdout)
        video = open(os.path.join(tmp, "bitstreams", "awesome_video.mpeg"), "rb").read()
        distinct, index_of, frames = [], {}, []
        for name in order:
            data = open(os.path.join(tmp, "images", name), "rb").read()
            key = hashlib.md5(data).hexdigest()
            if key not in index_of:
                img = stb_decode(os.path.join(tmp, "images", name))
                assert img.shape == (600, 400, 3), img.shape
                index_of[key] = len(distinct)
                distinct.append(img[:144].copy())
            frames.append(index_of[key])
        # .bit side file of the first frame: digest only (720 008 bytes)
        bit = open(os.path.join(tmp, "bitstreams", "image_1.bit"), "rb").read()
        np.savez_compressed(os.path.join(HERE, "refcompat_inputs.npz"), images=np.stack(distinct),
                            frame_image=np.array(frames, np.int32), names=np.array(order),
                            full_shape=np.array([600, 400, 3], np.int32))
        open(os.path.join(HERE, "refcompat_video.mpeg"), "wb").write(video)
        return {"video_bytes": len(video), "frames": len(order), "bit_file_bytes": len(bit),
                "bit_file_header": list(bit[:8])}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_kat():
    ref, port = oracle.Ref(), oracle.Port()
    rng = np.random.default_rng(504)
    kat = {}
    kat["qmatrix"] = {str(q): ref.qmatrix(q).tolist() for q in (-5, 1, 5, 12, 25, 49, 50, 51, 75, 89, 90, 99, 100, 150)}
    kat["ac_table"] = {f"{r},{a}": ref.ac_table_entry(r, a) for r in range(32) for a in range(40) if ref.ac_table_entry(r, a)}
    kat["slice_header_bits"] = {str(v): ref.slice_header_bits(1, v) for v in (0, 1, 67, 174, 254, 255)}
    # block bit strings
    blocks = []
    for i in range(400):
        zz = np.zeros(64, np.int32)
        k = int(rng.integers(0, 9))
        pos = rng.choice(64, k, replace=False)
        zz[pos] = rng.integers(-255, 256, k)
        if i % 3 == 0:
            zz[0] = rng.integers(-2042, 2043)
        if i % 11 == 0:
            zz = rng.integers(-2, 3, 64).astype(np.int32)
        if i % 13 == 0:
            zz[:] = 0
            zz[int(rng.integers(1, 64))] = int(rng.choice([-255, -129, -128, -127, -41, -40, -2, -1, 1, 2, 39, 40, 127, 128, 255]))
        for luma in (1, 0):
            try:
                blocks.append({"zz": zz.tolist(), "luma": luma, "bits": ref.block_bits(zz, luma)})
            except ValueError:
                blocks.append({"zz": zz.tolist(), "luma": luma, "bits": None})
    kat["blocks"] = blocks
    # DCT
    dct = []
    for i in range(24):
        blk = rng.integers(0, 256, 64, dtype=np.uint8) if i % 3 else (rng.integers(0, 2, 64) * 255).astype(np.uint8)
        if i == 0:
            blk[:] = 255
        if i == 1:
            blk[:] = 0
        dct.append({"in": blk.tolist(), "out": ref.fdct8x8(blk).tolist()})
    kat["dct"] = dct
    # colour: digests over all 2^24 inputs + every input where the result differs from the exact floor
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    rgb = np.stack([r.ravel(), g.ravel(), b.ravel()], 1)
    Y, Cb, Cr = ref.rgb_to_ycbcr(rgb)
    R, G, B = (rgb[:, i].astype(np.int64) for i in range(3))
    ex_y = (299 * R + 587 * G + 114 * B) // 1000
    ex_cb = (128_000_000 - 168_736 * R - 331_264 * G + 500_000 * B) // 1_000_000
    ex_cr = (128_000_000 + 500_000 * R - 418_688 * G - 81_312 * B) // 1_000_000
    col = {}
    for name, got, ex in (("Y", Y, ex_y), ("Cb", Cb, ex_cb), ("Cr", Cr, ex_cr)):
        diff = np.nonzero(got.astype(np.int64) != ex)[0]
        assert np.all(got[diff].astype(np.int64) == ex[diff] - 1)
        col[name] = {"sha256": hashlib.sha256(got.tobytes()).hexdigest(), "n_below_exact_floor": int(diff.size),
                     "below_exact_floor_index_sha256": hashlib.sha256(diff.astype(np.uint32).tobytes()).hexdigest(),
                     "first_below": diff[:64].tolist()}
    kat["colour"] = col
    kat["file_prologue"] = ref.file_prologue().hex()
    kat["frame_prefix"] = [{"i": i, "W": W, "H": H, "mode": m, "payload": n, "hex": ref.frame_prefix(i, W, H, m, n).hex()}
                           for (i, W, H, m, n) in [(0, 400, 600, 1, 560), (1, 400, 600, 1, 561), (29, 400, 600, 1, 0),
                                                   (31, 352, 240, 0, 3225), (32, 352, 240, 0, 3225), (255, 1920, 1080, 0, 78710),
                                                   (256, 1920, 1080, 0, 102929), (300, 3840, 2160, 0, 400000),
                                                   (7, 7680, 4320, 0, 1600000)]]
    # pictures: synthetic input (oracle generator), reference output
    pics = []
    for (W, H, n, q, kind, mode) in [(352, 240, 3, 12, 0, 0), (352, 240, 2, 5, 1, 0), (352, 240, 2, 50, 1, 0),
                                     (64, 48, 2, 50, 1, 0), (33, 47, 1, 12, 1, 0), (100, 70, 1, 89, 0, 0),
                                     (400, 600, 2, 12, 1, 1), (400, 600, 1, 75, 0, 1), (640, 480, 1, 12, 0, 0)]:
        for f in range(n):
            img = port.synth_rgb(12345, f, W, H, kind)
            pay, lev = ref.encode_picture(img, q, mode, want_levels=True)
            e = {"W": W, "H": H, "frame": f, "q": q, "kind": kind, "mode": mode, "seed": 12345,
                 "rgb_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
                 "payload_sha256": hashlib.sha256(pay).hexdigest(), "payload_bytes": len(pay),
                 "levels_sha256": hashlib.sha256(lev.tobytes()).hexdigest()}
            if len(pay) <= 4096:
                e["payload_hex"] = pay.hex()
            pics.append(e)
    kat["pictures"] = pics
    return kat


def main():
    assert os.path.isdir(REF_ROOT) and oracle.Ref.available(), "needs /root/reference and `make -C oracle ref`"
    kat = make_kat()
    kat["refcompat"] = make_refcompat()
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, separators=(",", ":"))
    for n in ("kat.json", "refcompat_inputs.npz", "refcompat_video.mpeg"):
        print(n, os.path.getsize(os.path.join(HERE, n)))


if __name__ == "__main__":
    main()
