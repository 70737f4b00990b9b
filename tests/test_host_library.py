"""libencoder.so: the host-C side of the boundary.

CPU part: the per-stage compatibility functions (reference `make sharedlib` API) against the
oracle and the golden vectors.  GPU part: the driver (GPU hot path + host headers) against the
oracle's stream and the reference binary's own output file."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

REF_EXPORTS_HOT = """VLC_encode bitvector_clone bitvector_concat bitvector_expand_size bitvector_fwrite bitvector_init
bitvector_new bitvector_pos bitvector_print bitvector_put_binstring bitvector_put_bit bitvector_put_byte
bitvector_put_byte_ent bitvector_put_byte_off bitvector_toarray check_dimensions convert_rgb_to_ycbcr display_u8arr
encode_blk_coeff encode_block_end encode_block_header_i encode_coeff_sz_fast encode_macblk_address_value
encode_macblk_encoding_value encode_macroblock_end encode_macroblock_header_i equalize_coefficients extract_8x8_block
fast_DCT mpeg1_file_header mpeg1_gop mpeg1_packet_header mpeg1_picture_header mpeg1_sequence_end mpeg1_sequence_header
mpeg1_slice mpeg1_sys_header print_array quantization run_length_encode scale_quantization_matrix subsampling_420
write_to_bitstream zigzag_scanning mpeg_encode_procedure DCT IDCT fast_IDCT dequantization upsampling insert_8x8_block
convert_ycbcr_to_rgb concat_char
Q_MATRIX ZIGZAG_ORDER START_FILE START_PICTURE blk_coeff_1_f blk_coeff_1_n blk_coeff_end blk_rle_lookup blk_rle_table
dc_sz_chroma_table dc_sz_luma_table encoding_table mv_encoding_table slice_start_code""".split()


@pytest.fixture(scope="module")
def host():
    import subprocess
    root = os.path.dirname(HERE)
    from ec504_imageencoder_b200 import _native, hostlib
    _native.build_cuda()
    subprocess.run(["make", "-s", "-C", root, "sharedlib"], check=True)
    return hostlib


def test_exports(host):
    L = host.lib()
    for name in REF_EXPORTS_HOT:
        assert hasattr(L, name), name


def _block_bits(host, zz, luma):
    L = host.lib()
    z = (C.c_int * 64)(*[int(v) for v in zz])
    eq = (C.c_int * 64)()
    rle = (C.c_int * 132)()
    L.equalize_coefficients(z, eq)
    L.run_length_encode(eq, rle)
    bv = L.bitvector_new(b"", 8)
    L.encode_block_header_i(luma, rle, bv)
    L.encode_block_end(bv)
    return bv.contents.bitstring()


def test_block_syntax_against_golden(host, port):
    with open(os.path.join(GOLD, "kat.json")) as f:
        kat = json.load(f)
    n = 0
    for b in kat["blocks"]:
        if b["bits"] is None:
            continue
        assert _block_bits(host, b["zz"], b["luma"]) == b["bits"]
        n += 1
    assert n > 600
    L = host.lib()
    bv = L.bitvector_new(b"", 8)
    L.mpeg1_slice(1, 0, bv)
    L.encode_macroblock_header_i(1, 1, bv)
    assert bv.contents.bitstring() == kat["slice_header_bits"]["0"]
    bv = L.bitvector_new(b"", 8)
    L.encode_macroblock_header_i(70, 1, bv)          # two escapes + increment 4 + type
    assert bv.contents.bitstring() == "00000001000" * 2 + "0011" + "1"


def test_stage_functions_against_oracle(host, port):
    L = host.lib()
    rng = np.random.default_rng(11)
    for q in (1, 12, 50, 89, 100):
        m = (C.c_int * 64)()
        L.scale_quantization_matrix(m, q)
        assert list(m) == port.qmatrix(q).tolist()
    for i in range(300):
        blk = rng.integers(0, 256, 64, dtype=np.uint8) if i % 2 else (rng.integers(0, 2, 64) * 255).astype(np.uint8)
        d = (C.c_double * 64)()
        L.fast_DCT(blk.ctypes.data, d)
        want = port.fdct8x8(blk)
        assert [int(x) for x in d] == want.tolist()
        q = (5, 12, 50)[i % 3]
        qi, zz = (C.c_int * 64)(), (C.c_int * 64)()
        L.quantization(d, qi, q)
        L.zigzag_scanning(qi, zz)
        assert list(zz) == port.quant_zigzag(want, port.qmatrix(q)).tolist()
    # extract_8x8_block
    plane = rng.integers(0, 256, (40, 64), dtype=np.uint8)
    out = np.zeros((8, 8), np.uint8)
    L.extract_8x8_block(plane.ctypes.data, 64, 24, 16, out.ctypes.data)
    assert np.array_equal(out, plane[16:24, 24:32])


def test_headers_against_golden(host):
    with open(os.path.join(GOLD, "kat.json")) as f:
        kat = json.load(f)
    L = host.lib()
    buf = np.zeros(27, np.uint8)
    L.mpeg1_file_header(2202035, buf.ctypes.data)
    L.mpeg1_sys_header(2202035, 0xE6, buf[12:].ctypes.data)
    assert buf.tobytes().hex() == kat["file_prologue"]
    for p in kat["frame_prefix"]:
        hour = p["i"] & 0xFF
        out = np.zeros(44, np.uint8)
        L.mpeg1_packet_header(1 + 3600 * hour, out.ctypes.data)
        W, H = (p["W"] & 0xFF, p["H"] & 0xFF) if p["mode"] == 1 else (p["W"], p["H"])
        L.mpeg1_sequence_header(W, H, 1, 4, 3, out[16:].ctypes.data)
        L.mpeg1_gop(0, hour, 0, 0, 0, 1, 0, out[28:].ctypes.data)
        bid = np.zeros(4, np.uint8)
        L.mpeg1_picture_header(0, 1, 0xFFFF, bid.ctypes.data, out[36:].ctypes.data)
        fwd = (44 + p["payload"] - 8) & 0xFFFF
        out[4], out[5] = fwd >> 8, fwd & 0xFF
        assert out.tobytes().hex() == p["hex"], p
    end = np.zeros(4, np.uint8)
    L.mpeg1_sequence_end(end.ctypes.data)
    assert end.tobytes() == b"\x00\x00\x01\xb7"


def test_stream_templates_against_oracle(host, port):
    """m1_stream_templates (the header bytes handed to the device-side stream assembly): every one of the
    256 prefixes equals the oracle's frame prefix for an empty payload, in both modes; prologue and trailer."""
    for mode, (W, H) in ((0, (1920, 1080)), (0, (352, 240)), (1, (400, 600))):
        prefix, prologue, trailer = host.stream_templates(W, H, mode)
        assert prologue.tobytes() == port.file_prologue()
        assert trailer.tobytes() == b"\x00\x00\x01\xb7"
        for i in range(256):
            assert prefix[44 * i:44 * i + 44].tobytes() == port.frame_prefix(i, W, H, mode, 0), (mode, i)
        # the reference's clock wraps: picture 256 + i carries the headers of picture i
        assert port.frame_prefix(256 + 7, W, H, mode, 0) == port.frame_prefix(7, W, H, mode, 0)


def test_bitvector_semantics(host):
    L = host.lib()
    bv = L.bitvector_new(b"101", 3)                  # size is a capacity hint, length = strlen
    assert bv.contents.cap == 3 and bv.contents.bitstring() == "101"
    L.bitvector_put_byte_off(bv, 0b00010110, 5, 3)   # low five bits
    assert bv.contents.bitstring() == "101" + "10110"
    L.bitvector_put_byte(bv, bytes([0b11000000]), 2)          # top two bits
    L.bitvector_put_byte_ent(bv, bytes([0x0F]))
    assert bv.contents.bitstring() == "10110110" + "11" + "00001111"
    other = L.bitvector_new(b"0011", 8)
    L.bitvector_concat(bv, other)
    assert bv.contents.bitstring().endswith("00001111" + "0011") and bv.contents.cap == 22
    cl = L.bitvector_clone(bv)
    assert cl.contents.bitstring() == bv.contents.bitstring() and cl.contents.cursor == 22
    for run, level, want in ((2, 2, "000110"), (1, 1, "11"), (1, 2, "00101"), (2, 40, "00000100000100101000")):
        assert L.encode_blk_coeff(run, level, 0).contents.bitstring() == want
    assert not L.encode_blk_coeff(1, 256, 0)         # NULL, as the reference (source/vlc.c:383)


# ---------------------------------------------------------------------------------------------
# GPU part
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_driver_stream_equals_oracle(host, port):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    for (W, H, n, q, mode, kind) in ((352, 240, 5, 12, 0, 0), (100, 70, 3, 50, 0, 1), (400, 600, 3, 12, 1, 1)):
        frames = np.stack([port.synth_rgb(77, f, W, H, kind) for f in range(n)])
        assert host.encode_frames_to_memory(frames, q, mode) == port.encode_stream(frames, q, mode)


@pytest.mark.gpu
def test_driver_with_device_stream_assembly(host, port, monkeypatch):
    """M1_DEVICE_STREAM=1: the driver lets the GPU place headers, payloads and trailers
    (m1cu_encode_host_stream) and writes each batch with one call; same bytes as the oracle's stream,
    also across the driver's batches (at most 256 pictures each)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    monkeypatch.setenv("M1_DEVICE_STREAM", "1")
    for (W, H, n, q, mode, kind) in ((352, 240, 5, 12, 0, 0), (100, 70, 3, 50, 0, 1), (400, 600, 3, 12, 1, 1),
                                     (96, 64, 300, 12, 0, 0)):
        frames = np.stack([port.synth_rgb(77, f, W, H, kind) for f in range(n)])
        assert host.encode_frames_to_memory(frames, q, mode) == port.encode_stream(frames, q, mode)


@pytest.mark.gpu
def test_driver_on_reference_fixture(host, port, tmp_path):
    """REF_COMPAT end to end through the C driver entry point: the stream equals the file the
    reference's own binary wrote for images.zip, except the 4 uninitialised bytes per frame."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from test_oracle_golden import masked_equal
    z = np.load(os.path.join(GOLD, "refcompat_inputs.npz"))
    ref_video = open(os.path.join(GOLD, "refcompat_video.mpeg"), "rb").read()
    images, frame_image = z["images"], z["frame_image"]
    frames = np.zeros((len(frame_image), 600, 400, 3), np.uint8)       # rows >= 144 are never read in REF_COMPAT
    for i, idx in enumerate(frame_image):
        frames[i, :144] = images[int(idx)]
    ours = host.encode_frames_to_memory(frames, 12, 1)
    sizes = [len(port.encode_picture(images[int(i)], 12, 1)) for i in frame_image]
    assert masked_equal(ours, ref_video, 30, sizes)
    path = str(tmp_path / "v.mpeg")
    assert host.lib().m1_encode_frames_to_file(path.encode(), frames.ctypes.data, 30, 400, 600, 3, 12, 1) == 0
    assert open(path, "rb").read() == ours


@pytest.mark.gpu
def test_mpeg_encode_procedure_return_codes(host, tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    # unwritable output -> 1 (reference include/encoder.h:77-80)
    assert host.mpeg_encode_procedure(str(tmp_path / "imgs"), str(tmp_path / "bs"), str(tmp_path / "nodir" / "v.mpeg")) == 1
    # missing images folder -> created, 0 (:111-116); prologue already written (27 bytes)
    v = tmp_path / "v.mpeg"
    assert host.mpeg_encode_procedure(str(tmp_path / "imgs"), str(tmp_path / "bs"), str(v)) == 0
    assert (tmp_path / "imgs").is_dir() and (tmp_path / "bs").is_dir() and v.stat().st_size == 27
    # empty images folder -> -1 (:175-183); the reference has written the 27-byte pack + system header by then (:85-89)
    assert host.mpeg_encode_procedure(str(tmp_path / "imgs"), str(tmp_path / "bs"), str(v)) == -1
    assert v.stat().st_size == 27


def test_decode_helpers_against_reference(host):
    """the decoder-side exports, against the reference's own objects."""
    import oracle
    if not oracle.Ref.available():
        pytest.skip("oracle/_ref not built")
    L = host.lib()
    R = C.CDLL(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libm1ref.so"))
    rng = np.random.default_rng(8)
    for i in range(40):
        blk = rng.integers(0, 256, 64, dtype=np.uint8)
        a, b = np.zeros(64, np.float32), np.zeros(64, np.float32)
        L.DCT(blk.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p))
        R.DCT(blk.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a, b)
        pa, pb = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
        L.IDCT(a.ctypes.data_as(C.c_void_p), pa.ctypes.data_as(C.c_void_p))
        R.IDCT(a.ctypes.data_as(C.c_void_p), pb.ctypes.data_as(C.c_void_p))
        assert np.array_equal(pa, pb)
        q = rng.integers(-40, 41, 64).astype(np.int32)
        da, db = np.zeros(64, np.float64), np.zeros(64, np.float64)
        L.dequantization(q.ctypes.data_as(C.c_void_p), da.ctypes.data_as(C.c_void_p))
        R.dequantization(q.ctypes.data_as(C.c_void_p), db.ctypes.data_as(C.c_void_p))
        assert np.array_equal(da, db)
        fa, fb = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
        L.fast_IDCT(da.ctypes.data_as(C.c_void_p), fa.ctypes.data_as(C.c_void_p))
        R.fast_IDCT(da.ctypes.data_as(C.c_void_p), fb.ctypes.data_as(C.c_void_p))
        assert np.array_equal(fa, fb)
    W, H = 32, 16
    cb, cr = rng.integers(0, 256, (H // 2) * (W // 2), dtype=np.uint8), rng.integers(0, 256, (H // 2) * (W // 2), dtype=np.uint8)
    for lib_ in (L, R):
        lib_.upsampling.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        lib_.insert_8x8_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    outs = []
    for lib_ in (L, R):
        p, q2 = C.c_void_p(), C.c_void_p()
        lib_.upsampling(cb.ctypes.data, cr.ctypes.data, W, H, C.byref(p), C.byref(q2))
        outs.append((C.string_at(p, W * H), C.string_at(q2, W * H)))
    assert outs[0] == outs[1]
    assert np.array_equal(np.frombuffer(outs[0][0], np.uint8).reshape(H, W)[::2, ::2], cb.reshape(H // 2, W // 2))
    plane_a, plane_b = np.zeros((24, 40), np.uint8), np.zeros((24, 40), np.uint8)
    blk = rng.integers(0, 256, (8, 8), dtype=np.uint8)
    L.insert_8x8_block(plane_a.ctypes.data, 40, 16, 8, blk.ctypes.data)
    R.insert_8x8_block(plane_b.ctypes.data, 40, 16, 8, blk.ctypes.data)
    assert np.array_equal(plane_a, plane_b) and np.array_equal(plane_a[8:16, 16:24], blk)


@pytest.mark.gpu
def test_mpeg_encode_procedure_on_a_jpeg_folder(host, port, tmp_path):
    """The folder-scanning entry point end to end: JPEGs written here -> stb_image decode (host) ->
    GPU encode -> host headers -> .mpeg + image_N.bit side files, against the oracle's stream built
    from the same stb-decoded pixels in the same readdir order."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    PIL = pytest.importorskip("PIL.Image")
    L = host.lib()

    class Img(C.Structure):
        _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("data", C.POINTER(C.c_ubyte))]
    L.read_jpeg.restype = C.POINTER(Img)
    L.read_jpeg.argtypes = [C.c_char_p]
    L.free_image.argtypes = [C.POINTER(Img)]
    imgs = tmp_path / "images"
    imgs.mkdir()
    rng = np.random.default_rng(3)
    W, H = 208, 176
    for i in range(4):
        yy, xx = np.mgrid[0:H, 0:W]
        a = np.stack([(xx * 255 // W + 20 * i) % 256, (yy * 255 // H) % 256, ((xx + yy) // 2 + rng.integers(0, 30, (H, W))) % 256], -1)
        PIL.fromarray(a.astype(np.uint8)).save(str(imgs / f"pic_{i}.jpg"), quality=92)
    probe = L.read_jpeg(str(imgs / "pic_0.jpg").encode())
    if not probe:
        pytest.skip("libencoder.so was built without stb_image.h")
    L.free_image(probe)
    video = tmp_path / "out" / "v.mpeg"
    (tmp_path / "out").mkdir()
    assert host.mpeg_encode_procedure(str(imgs), str(tmp_path / "out"), str(video), 12) == 0
    frames = []
    for name in os.listdir(imgs):                      # readdir order, like the driver
        p = L.read_jpeg(str(imgs / name).encode())
        assert p and p.contents.channels == 3
        frames.append(np.ctypeslib.as_array(p.contents.data, shape=(H, W, 3)).copy())
        L.free_image(p)
    frames = np.stack(frames)
    assert video.read_bytes() == port.encode_stream(frames, 12, 1)          # REF_COMPAT is the driver's default
    # .bit side file: int32 width, height, then the full-resolution Y, Cb, Cr planes
    bit = (tmp_path / "out" / "image_1.bit").read_bytes()
    y, cb, cr = port.rgb_to_ycbcr(frames[0].reshape(-1, 3))
    assert bit == np.array([W, H], np.int32).tobytes() + y.tobytes() + cb.tobytes() + cr.tobytes()
    # M1_MODE=full: whole pictures
    os.environ["M1_MODE"] = "full"
    try:
        assert host.mpeg_encode_procedure(str(imgs), str(tmp_path / "out"), str(video), 12) == 0
    finally:
        del os.environ["M1_MODE"]
    assert video.read_bytes() == port.encode_stream(frames, 12, 0)


def _read_folder(L, folder, W, H):
    """The stb-decoded pictures of a folder in readdir order (what the driver sees); undecodable files skipped."""
    class Img(C.Structure):
        _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("data", C.POINTER(C.c_ubyte))]
    L.read_jpeg.restype = C.POINTER(Img)
    L.read_jpeg.argtypes = [C.c_char_p]
    L.free_image.argtypes = [C.POINTER(Img)]
    frames = []
    for name in os.listdir(folder):
        if ".jpg" not in name and ".jpeg" not in name:
            continue
        p = L.read_jpeg(os.path.join(str(folder), name).encode())
        if not p:
            continue
        assert (p.contents.width, p.contents.height, p.contents.channels) == (W, H, 3)
        frames.append(np.ctypeslib.as_array(p.contents.data, shape=(H, W, 3)).copy())
        L.free_image(p)
    return np.stack(frames)


@pytest.mark.gpu
def test_folder_pipeline_multi_batch_and_legacy_path(host, port, tmp_path, monkeypatch):
    """mpeg_encode_procedure decodes on worker threads into a two-slot pinned ring
    while the GPU encodes the previous batch.  600 small pictures = three ring turns; the bytes of the
    video and of every image_N.bit equal the oracle's and equal the decode-everything-first path
    (M1_DECODE_THREADS=0, the reference's order of work), also with a file stb cannot decode in the
    folder (skipped, as the reference skips it) and with one of a different size (-1, 27-byte file)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    PIL = pytest.importorskip("PIL.Image")
    L = host.lib()
    imgs = tmp_path / "images"
    imgs.mkdir()
    rng = np.random.default_rng(11)
    W, H, N = 64, 48, 600
    for i in range(N):
        a = (np.add.outer(np.arange(H) * 3 + i, np.arange(W) * 2)[..., None] + rng.integers(0, 40, (H, W, 3))) % 256
        PIL.fromarray(a.astype(np.uint8)).save(str(imgs / f"p{i:04d}.jpg"), quality=90)
    if not L.read_jpeg(str(imgs / "p0000.jpg").encode()):
        pytest.skip("libencoder.so was built without stb_image.h")
    monkeypatch.setenv("M1_MODE", "full")
    frames = _read_folder(L, imgs, W, H)
    assert len(frames) == N
    want = port.encode_stream(frames, 12, 0)

    def run(tag, **env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = tmp_path / tag
        out.mkdir()
        rc = host.mpeg_encode_procedure(str(imgs), str(out), str(out / "v.mpeg"), 12)
        for k in env:
            monkeypatch.delenv(k)
        return rc, out

    rc, out = run("stream")
    assert rc == 0 and (out / "v.mpeg").read_bytes() == want
    for i in (0, 1, 255, 256, 511, 512, N - 1):
        y, cb, cr = port.rgb_to_ycbcr(frames[i].reshape(-1, 3))
        assert (out / f"image_{i + 1}.bit").read_bytes() == np.array([W, H], np.int32).tobytes() + y.tobytes() + cb.tobytes() + cr.tobytes()
    assert len(list(out.glob("image_*.bit"))) == N
    rc, out1 = run("onethread", M1_DECODE_THREADS="1")
    assert rc == 0 and (out1 / "v.mpeg").read_bytes() == want
    rc, out0 = run("legacy", M1_DECODE_THREADS="0")
    assert rc == 0 and (out0 / "v.mpeg").read_bytes() == want
    assert (out0 / "image_300.bit").read_bytes() == (out / "image_300.bit").read_bytes()
    rc, outd = run("devstream", M1_DEVICE_STREAM="1")
    assert rc == 0 and (outd / "v.mpeg").read_bytes() == want

    # a file stb cannot decode: its header does not parse -> decode-first path, the file is skipped
    (imgs / "zz_broken.jpg").write_bytes(b"this is not a JPEG")
    frames2 = _read_folder(L, imgs, W, H)
    assert len(frames2) == N
    rc, outb = run("broken")
    assert rc == 0 and (outb / "v.mpeg").read_bytes() == port.encode_stream(frames2, 12, 0)
    (imgs / "zz_broken.jpg").unlink()

    # a picture of another size -> -1, and only the pack + system header in the file
    PIL.fromarray(np.zeros((H + 16, W, 3), np.uint8)).save(str(imgs / "odd.jpg"))
    rc, outm = run("mismatch")
    assert rc == -1 and (outm / "v.mpeg").stat().st_size == 27
