"""GPU parity: the CUDA path through the C ABI vs the oracle, bit for bit.

Integer/byte work => the bar is exact equality of (a) the synthetic input bytes, (b) the
zigzag-ordered quantised levels, (c) every payload byte, (d) the assembled stream.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import MODE_FULL, MODE_REF_COMPAT, SYNTH_NATURAL, SYNTH_NOISE, SYNTH_GREY, SYNTH_RG_EQUAL, SYNTH_SCATTERED  # noqa: E402


@pytest.fixture(scope="module")
def m1():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import ec504_imageencoder_b200 as m
    return m


def _encode_both(m1, port, W, H, n, q, kind, mode=MODE_FULL, seed=12345, first=0):
    enc = m1.M1Encoder(W, H, 3, mode, q, max_frames=n)
    rgb = enc.synth_rgb(seed, first, n, kind)
    res = enc.encode_device(rgb, want_levels=True)
    pay = res.payloads()
    host = rgb.cpu().numpy()
    lev = res.levels.cpu().numpy()
    for f in range(n):
        assert np.array_equal(host[f], port.synth_rgb(seed, first + f, W, H, kind))
        rp, rl = port.encode_picture(host[f], q, mode, want_levels=True)
        assert np.array_equal(lev[f], rl), f"levels differ: {W}x{H} q={q} kind={kind} frame={f}"
        assert pay[f] == rp, f"payload differs: {W}x{H} q={q} kind={kind} frame={f}"
    enc.close()
    return pay


@pytest.mark.parametrize("kind", [SYNTH_NATURAL, SYNTH_NOISE])
@pytest.mark.parametrize("q", [5, 12, 50])
def test_sif_config0(m1, port, q, kind):
    """BASELINE configs[0] geometry: 352x240, a few frames per quality."""
    _encode_both(m1, port, 352, 240, 3, q, kind)


@pytest.mark.parametrize("W,H", [(16, 16), (17, 16), (16, 17), (33, 47), (100, 70), (640, 480), (641, 479),
                                 (1, 1), (2050, 18)])
def test_ragged_sizes(m1, port, W, H):
    """Edge replication to the coded size, single/multiple chunks per slice."""
    _encode_both(m1, port, W, H, 2, 12, SYNTH_NOISE)
    _encode_both(m1, port, W, H, 1, 50, SYNTH_NATURAL)


@pytest.mark.parametrize("W,H", [(65535, 16), (65520, 32), (16, 65535), (32, 16384), (8192, 8192)])
def test_extreme_geometries(m1, port, W, H):
    """The largest width / height the context accepts (65535: 4096 macroblocks per slice = 256 chunks, or 4096 slices,
    whose vertical position byte wraps 16 times: source/mpeg1_blk.c:12-20 writes (vertical_pos + 1) & 0xff) and a
    67-megapixel square picture."""
    _encode_both(m1, port, W, H, 1, 12, SYNTH_NOISE)


@pytest.mark.parametrize("q", [1, 12, 50, 75, 89])
def test_ref_compat(m1, port, q):
    """Literal traversal of include/encoder.h:238-443 on a 400x600 picture (the fixture size)."""
    _encode_both(m1, port, 400, 600, 2, q, SYNTH_NOISE, MODE_REF_COMPAT)
    _encode_both(m1, port, 400, 600, 1, q, SYNTH_NATURAL, MODE_REF_COMPAT)
    _encode_both(m1, port, 96, 144, 1, q, SYNTH_NOISE, MODE_REF_COMPAT)


def test_1080p_one_frame(m1, port):
    pay = _encode_both(m1, port, 1920, 1080, 1, 12, SYNTH_NATURAL)
    assert 20_000 < len(pay[0]) < 400_000


def test_against_reference_functions(m1, ref):
    """Same comparison against the UNMODIFIED reference functions when oracle/_ref travelled."""
    W, H, q = 352, 240, 12
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, q, max_frames=2)
    rgb = enc.synth_rgb(7, 0, 2, SYNTH_NOISE)
    res = enc.encode_device(rgb, want_levels=True)
    pay, lev, host = res.payloads(), res.levels.cpu().numpy(), rgb.cpu().numpy()
    for f in range(2):
        rp, rl = ref.encode_picture(host[f], q, MODE_FULL, want_levels=True)
        assert np.array_equal(lev[f], rl)
        assert pay[f] == rp


def test_host_path_equals_device_path(m1, port):
    W, H, n = 352, 240, 4
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, 12, max_frames=n)
    rgb = enc.synth_rgb(99, 10, n, SYNTH_NATURAL)
    dev = enc.encode_device(rgb).payloads()
    hp, lev = enc.encode_host(rgb.cpu().numpy(), want_levels=True)
    assert hp == dev
    assert lev.shape == (n, enc.macroblocks, 6, 64)
    # 4-channel input: the 4th byte is ignored (source/image_processing.c:94-97 indexes i*channels)
    enc4 = m1.M1Encoder(W, H, 4, MODE_FULL, 12, max_frames=n)
    rgba = np.concatenate([rgb.cpu().numpy(), np.full((n, H, W, 1), 77, np.uint8)], axis=3)
    hp4, _ = enc4.encode_host(rgba)
    assert hp4 == dev


def test_color_all_2_24(m1, port):
    """Every RGB triple through the device colour conversion (the only floating point on the path)."""
    enc = m1.M1Encoder(4096, 4096, 3, MODE_FULL, 12, max_frames=1)
    v = torch.arange(256, dtype=torch.uint8, device="cuda")
    r, g, b = torch.meshgrid(v, v, v, indexing="ij")
    rgb = torch.stack([r, g, b], dim=-1).reshape(4096, 4096, 3).contiguous()
    Y, Cb, Cr = enc.ycbcr_planes(rgb)
    enc.check()
    oy, ocb, ocr = port.rgb_to_ycbcr(rgb.cpu().numpy().reshape(-1, 3))
    assert np.array_equal(Y.cpu().numpy().ravel(), oy)
    assert np.array_equal(Cb.cpu().numpy().ravel(), ocb)
    assert np.array_equal(Cr.cpu().numpy().ravel(), ocr)


def test_encode_kernel_all_colours_all_alignments(m1, port):
    """All 2^24 colours through k_encode_chunks' colour phase (not through the plane kernel): four
    4096x4096 pictures holding every colour once, shifted by 0..3 pixels so that every colour meets each of
    the four byte alignments of a packed 3-byte pixel.  Quality 89: the finest quantiser the reference can
    encode keeps single-sample errors visible in the levels.  (tests/test_gpu_variants.py repeats it on the
    integer-colour build.)"""
    g = np.arange(1 << 24, dtype=np.uint32)
    base = np.stack([g >> 16, (g >> 8) & 255, g & 255], axis=1).astype(np.uint8)
    enc = m1.M1Encoder(4096, 4096, 3, MODE_FULL, 89, max_frames=1)
    for shift in range(4):
        img = np.ascontiguousarray(np.roll(base, shift, axis=0).reshape(1, 4096, 4096, 3))
        res = enc.encode_device(torch.from_numpy(img).cuda(), want_levels=True)
        rp, rl = port.encode_picture(img[0], 89, MODE_FULL, want_levels=True)
        assert np.array_equal(res.levels[0].cpu().numpy(), rl), f"levels differ (shift {shift})"
        assert res.payloads()[0] == rp, f"payload differs (shift {shift})"
    enc.close()


@pytest.mark.parametrize("kind", [SYNTH_GREY, SYNTH_RG_EQUAL])
@pytest.mark.parametrize("W,H,q", [(352, 240, 12), (1920, 1080, 50), (100, 70, 89)])
def test_exact_quotient_worst_cases(m1, port, kind, W, H, q):
    """Grey and r == g pictures: every pixel is one of the colours where the reference's double chain can
    land one ulp below an integer, the worst case for any shortcut in the colour
    arithmetic (on the integer-colour build every 2x2 quad of every chunk is queued for recomputation)."""
    _encode_both(m1, port, W, H, 2, q, kind)


@pytest.mark.parametrize("W,H,q,mode", [(352, 240, 12, MODE_FULL), (1920, 1080, 12, MODE_FULL), (1920, 1080, 5, MODE_FULL),
                                        (640, 368, 20, MODE_FULL), (400, 600, 12, MODE_REF_COMPAT)])
def test_scattered_busy_tiles(m1, port, W, H, q, mode):
    """A quarter of the 8x8 pixel tiles are noise, the rest smooth: flat and busy blocks side by side in every warp, so
    the encoder's three block paths (DC only, eight lanes per busy block, one thread per block) all run inside one chunk."""
    _encode_both(m1, port, W, H, 2, q, SYNTH_SCATTERED, mode)


@pytest.mark.parametrize("channels", [3, 4])
def test_colour_exception_heavy_pictures(m1, port, channels):
    """Pictures made of the colours where the reference's double chain truncates below the exact
    rational value (greys, r == g, g == b, 1000 | 299r+587g+114b): every pixel takes the integer
    path's fix-up branch.  Quality 89 (the finest quantiser noise stays encodable at) keeps
    single-sample differences visible in the levels."""
    rng = np.random.default_rng(5)
    W, H, n = 352, 240, 6
    v = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    img = v.copy()
    img[0] = v[0][..., :1]                                     # greys
    img[1][..., 1] = img[1][..., 0]                            # r == g
    img[2][..., 2] = img[2][..., 1]                            # g == b
    img[3] = 0                                                 # black
    img[3][:, W // 2:] = 255                                   # and white
    base = 299 * v[4][..., 0].astype(np.int64) + 587 * v[4][..., 1].astype(np.int64)
    hit = (base[..., None] + 114 * np.arange(256)) % 1000 == 0      # at most one b per (r, g)
    ok = hit.any(axis=-1)
    img[4][..., 2] = np.where(ok, hit.argmax(axis=-1), v[4][..., 2]).astype(np.uint8)
    assert ok.mean() > 0.1
    # img[5] stays random
    if channels == 4:
        img = np.concatenate([img, rng.integers(0, 256, (n, H, W, 1), dtype=np.uint8)], axis=3)
    enc = m1.M1Encoder(W, H, channels, MODE_FULL, 89, max_frames=n)
    hp, lev = enc.encode_host(img, want_levels=True)
    for f in range(n):
        rp, rl = port.encode_picture(np.ascontiguousarray(img[f][..., :3]), 89, MODE_FULL, want_levels=True)
        assert np.array_equal(lev[f], rl), f"levels differ on frame {f}"
        assert hp[f] == rp, f"payload differs on frame {f}"
    # and the conversion itself, sample for sample
    enc1 = m1.M1Encoder(W, H, 3, MODE_FULL, 1, max_frames=1)
    for f in range(n):
        rgb = torch.from_numpy(np.ascontiguousarray(img[f][..., :3])).cuda()
        Y, Cb, Cr = enc1.ycbcr_planes(rgb)
        oy, ocb, ocr = port.rgb_to_ycbcr(np.ascontiguousarray(img[f][..., :3]).reshape(-1, 3))
        assert np.array_equal(Y.cpu().numpy().ravel(), oy)
        assert np.array_equal(Cb.cpu().numpy().ravel(), ocb)
        assert np.array_equal(Cr.cpu().numpy().ravel(), ocr)


def test_unencodable_level_reported(m1, port):
    """quality 91 + a half-block vertical step gives a coded AC level of 308: the reference
    returns NULL from encode_blk_coeff and crashes (source/vlc.c:383); the oracle refuses the
    input and the library reports M1CU_ERR_LEVEL instead."""
    W = H = 16
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, 91, max_frames=1)
    img = np.zeros((1, H, W, 3), np.uint8)
    img[0, 0:4] = 255
    img[0, 8:12] = 255
    with pytest.raises(ValueError):
        port.encode_picture(img[0], 91, MODE_FULL)
    with pytest.raises(m1.M1Error) as ei:
        enc.encode_host(img)
    assert ei.value.code == -4
    # one quality step lower the same picture is inside the reference's envelope and must match
    enc89 = m1.M1Encoder(W, H, 3, MODE_FULL, 89, max_frames=1)
    hp, _ = enc89.encode_host(img)
    assert hp[0] == port.encode_picture(img[0], 89, MODE_FULL)


def test_capacity_error_and_retry(m1, port):
    W, H = 64, 64
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, 50, max_frames=2)
    rgb = enc.synth_rgb(5, 0, 2, SYNTH_NOISE)
    tiny = enc.alloc_outputs(2, out_bytes=64)
    with pytest.raises(m1.M1Error) as ei:
        enc.encode_device(rgb, res=tiny)
    assert ei.value.code == -3
    ok = enc.encode_device(rgb)          # context still usable afterwards
    host = rgb.cpu().numpy()
    assert ok.payloads() == [port.encode_picture(host[f], 50, MODE_FULL) for f in range(2)]


@pytest.mark.parametrize("mode,W,H", [(MODE_FULL, 352, 240), (MODE_FULL, 100, 52), (MODE_REF_COMPAT, 400, 600)])
def test_device_stream_assembly(m1, port, mode, W, H):
    """m1cu_assemble_stream: the .mpeg bytes built on the device equal the oracle's stream (prologue, 44-byte
    prefixes with the packet length patched in, payloads, trailers), also for a batch that starts at another
    picture index and crosses the reference's 256-picture clock wrap, and without the prologue."""
    import torch
    n, q = 9, 12
    enc = m1.M1Encoder(W, H, 3, mode, q, max_frames=n)
    rgb = enc.synth_rgb(777, 0, n, SYNTH_NATURAL)
    res = enc.encode_device(rgb)
    out, nbytes = enc.assemble_stream(res)
    enc.check()
    got = out[:int(nbytes.item())].cpu().numpy().tobytes()
    want = port.encode_stream(rgb.cpu().numpy(), q, mode)
    assert got == want
    # same payloads as pictures 250 .. 258 of a longer sequence, appended to an existing file (no prologue)
    pay = res.payloads()
    out2, nb2 = enc.assemble_stream(res, first_frame_index=250, prologue=False)
    enc.check()
    got2 = out2[:int(nb2.item())].cpu().numpy().tobytes()
    want2 = b"".join(port.frame_prefix(250 + f, W, H, mode, len(pay[f])) + pay[f] + b"\x00\x00\x01\xb7" for f in range(n))
    assert got2 == want2
    # two calls into one buffer: pictures 0..8 with the prologue, then the same payloads again as pictures 9..17
    end1 = int(nbytes.item())
    big = torch.zeros((2 * len(want) + 64 + 15) // 16 * 16, dtype=torch.uint8, device=out.device)
    enc.assemble_stream(res, out=big)
    _, nb3 = enc.assemble_stream(res, first_frame_index=n, prologue=False, out=big, offset=end1)
    enc.check()
    want3 = want + b"".join(port.frame_prefix(n + f, W, H, mode, len(pay[f])) + pay[f] + b"\x00\x00\x01\xb7" for f in range(n))
    assert int(nb3.item()) == len(want3) and big[:len(want3)].cpu().numpy().tobytes() == want3
    # a stream buffer that is too small is reported, not overrun
    small = torch.zeros(((len(want) // 2) + 15) // 16 * 16, dtype=torch.uint8, device=out.device)
    enc.assemble_stream(res, out=small)
    with pytest.raises(m1.M1Error):
        enc.check()
    enc.close()


def test_host_stream_entry_point(m1, port):
    """m1cu_encode_host_stream through the C ABI: host pictures in, finished file bytes out (equal to the
    oracle's stream); a too small host buffer is an error, not an overrun."""
    import ctypes as C
    from ec504_imageencoder_b200 import _native, hostlib
    W, H, n, q = 176, 144, 6, 12
    frames = np.stack([port.synth_rgb(5, f, W, H, SYNTH_NOISE) for f in range(n)])
    want = port.encode_stream(frames, q, MODE_FULL)
    prefix, prologue, trailer = hostlib.stream_templates(W, H, MODE_FULL)
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, q, max_frames=n)
    lib = _native.m1cu()
    out = np.zeros(len(want) + 64, np.uint8)
    got = C.c_size_t(0)
    rc = lib.m1cu_encode_host_stream(enc._h, frames.ctypes.data, n, 0, prefix.ctypes.data, prologue.ctypes.data,
                                     trailer.ctypes.data, out.ctypes.data, out.size, C.byref(got))
    assert rc == 0 and out[:got.value].tobytes() == want
    small = np.zeros(len(want) // 2, np.uint8)
    rc = lib.m1cu_encode_host_stream(enc._h, frames.ctypes.data, n, 0, prefix.ctypes.data, prologue.ctypes.data,
                                     trailer.ctypes.data, small.ctypes.data, small.size, C.byref(got))
    assert rc == -3
    enc.close()
