"""The oracle port against the UNMODIFIED reference functions (oracle/_ref/libm1ref.so), live.
Skipped where oracle/_ref was not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

from oracle import MODE_FULL, MODE_REF_COMPAT


def test_tables_and_matrices(port, ref):
    for r in range(32):
        for a in range(41):
            assert port.ac_table_entry(r, a) == ref.ac_table_entry(r, a), (r, a)
    for q in range(-2, 104):
        assert (port.qmatrix(q) == ref.qmatrix(q)).all(), q


def test_colour_every_triple(port, ref):
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    rgb = np.stack([r.ravel(), g.ravel(), b.ravel()], 1)
    for a, c in zip(port.rgb_to_ycbcr(rgb), ref.rgb_to_ycbcr(rgb)):
        assert np.array_equal(a, c)
    rgba = np.concatenate([rgb[:100000], np.full((100000, 1), 9, np.uint8)], 1)
    for a, c in zip(port.rgb_to_ycbcr(rgba), ref.rgb_to_ycbcr(rgba)):
        assert np.array_equal(a, c)


def test_subsample(port, ref):
    rng = np.random.default_rng(3)
    for W, H in ((16, 16), (64, 48), (352, 240)):
        cb = rng.integers(0, 256, W * H, dtype=np.uint8)
        cr = rng.integers(0, 256, W * H, dtype=np.uint8)
        rcb, rcr = ref.subsample_420(cb, cr, W, H)
        assert np.array_equal(port.subsample_420(cb, W, H), rcb)
        assert np.array_equal(port.subsample_420(cr, W, H), rcr)


def test_dct_and_quant(port, ref):
    rng = np.random.default_rng(1)
    for i in range(1500):
        blk = rng.integers(0, 256, 64, dtype=np.uint8) if i % 2 else (rng.integers(0, 2, 64) * 255).astype(np.uint8)
        d = ref.fdct8x8(blk)
        assert (port.fdct8x8(blk) == d).all()
        assert ref.fdct_is_integral(blk)             # fast_DCT outputs are exact integers
        q = (1, 12, 50, 89, 100)[i % 5]
        assert (port.quant_zigzag(d, port.qmatrix(q)) == ref.quant_zigzag(d, q)).all()
    # extreme blocks stay inside the ranges the kernels rely on (|AC| <= 1022, 0 <= DC <= 2042)
    for blk in (np.full(64, 255, np.uint8), np.zeros(64, np.uint8), np.tile([255, 0], 32).astype(np.uint8),
                np.repeat([255, 0], 32).astype(np.uint8), ((np.indices((8, 8)).sum(0) % 2) * 255).astype(np.uint8).ravel()):
        d = ref.fdct8x8(blk)
        assert 0 <= d[0] <= 2042 and np.abs(d[1:]).max() <= 1022
        assert (port.fdct8x8(blk) == d).all()


def test_block_bits(port, ref):
    rng = np.random.default_rng(2)
    n = 0
    for i in range(6000):
        zz = np.zeros(64, np.int32)
        k = int(rng.integers(0, 9))
        zz[rng.choice(64, k, replace=False)] = rng.integers(-255, 256, k)
        if i % 3 == 0:
            zz[0] = rng.integers(-2042, 2043)
        if i % 7 == 0:
            zz = rng.integers(-3, 4, 64).astype(np.int32)
        for luma in (0, 1):
            try:
                want = ref.block_bits(zz, luma)
            except ValueError:
                with pytest.raises(ValueError):
                    port.block_bits(zz, luma)
                continue
            assert port.block_bits(zz, luma) == want
            n += 1
    assert n > 10000
    assert ref.slice_header_bits(1, 0) == "00000000000000000000000100000001" + "00001" + "0" + "11"


@pytest.mark.parametrize("W,H", [(352, 240), (64, 48), (100, 70), (33, 47), (96, 144), (400, 600)])
def test_pictures(port, ref, W, H):
    for kind in (0, 1):
        for q in (5, 12, 50, 89):
            img = port.synth_rgb(12345, 3, W, H, kind)
            pa, la = port.encode_picture(img, q, MODE_FULL, True)
            pb, lb = ref.encode_picture(img, q, MODE_FULL, True)
            assert pa == pb and (la == lb).all()
            if W >= 96 and H >= 144:
                pa, la = port.encode_picture(img, q, MODE_REF_COMPAT, True)
                pb, lb = ref.encode_picture(img, q, MODE_REF_COMPAT, True)
                assert pa == pb and (la == lb).all()


def test_headers(port, ref):
    assert port.file_prologue() == ref.file_prologue()
    for i in (0, 1, 5, 31, 32, 255, 256, 300, 1000):
        for mode in (0, 1):
            for W, H, n in ((400, 600, 12345), (1920, 1080, 99999), (352, 240, 0), (7680, 4320, 1 << 21)):
                assert port.frame_prefix(i, W, H, mode, n) == ref.frame_prefix(i, W, H, mode, n), (i, mode, W, H)
