"""GPU parity at the BASELINE.json geometries (configs[1]..[4]).  The oracle is run on a few
pictures per geometry (seconds each); the full-size batches are covered by size-independent
properties: batch independence, determinism, slice structure, offset/size consistency."""
import hashlib
import re

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import MODE_FULL, SYNTH_NATURAL, SYNTH_NOISE  # noqa: E402


@pytest.fixture(scope="module")
def m1():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import ec504_imageencoder_b200 as m
    return m


def _check_frames(m1, port, W, H, q, kind, frames, seed=12345, levels=True):
    n = len(frames)
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, q, max_frames=1)
    for f in frames:
        rgb = enc.synth_rgb(seed, f, 1, kind)
        res = enc.encode_device(rgb, want_levels=levels)
        host = rgb[0].cpu().numpy()
        rp, rl = port.encode_picture(host, q, MODE_FULL, want_levels=True)
        if levels:
            assert np.array_equal(res.levels[0].cpu().numpy(), rl), (W, H, q, kind, f)
        assert res.payloads()[0] == rp, (W, H, q, kind, f)
    enc.close()
    return n


@pytest.mark.parametrize("q", [5, 12, 50])
def test_config2_4k_quality_sweep(m1, port, q):
    """configs[2]: 3840x2160, low / default / high quality (all inside the reference's crash envelope)."""
    _check_frames(m1, port, 3840, 2160, q, SYNTH_NATURAL, [0])
    _check_frames(m1, port, 3840, 2160, q, SYNTH_NOISE, [299])


def test_config4_8k_slice_stress(m1, port):
    """configs[4]: 7680x4320 -- 480 macroblocks per slice = 15 chunks stitched at bit granularity."""
    _check_frames(m1, port, 7680, 4320, 12, SYNTH_NATURAL, [0])
    _check_frames(m1, port, 7680, 4320, 50, SYNTH_NOISE, [119], levels=False)


def test_config1_1080p_batch_properties(m1, port):
    """configs[1]: 300 frames of 1920x1080 in ONE call; oracle on a sample, properties on all."""
    W, H, n, q = 1920, 1080, 300, 12
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, q, max_frames=n)
    rgb = enc.synth_rgb(12345, 0, n, SYNTH_NATURAL)
    res = enc.encode_device(rgb)
    sizes = res.frame_bytes.cpu().numpy().astype(np.int64)
    offs = res.frame_offsets.cpu().numpy()
    # offsets: ascending, 16-byte aligned, consistent with the sizes
    assert offs[0] == 0 and np.all(offs % 16 == 0)
    assert np.array_equal(offs[1:], np.cumsum((sizes + 15) // 16 * 16))
    pays = res.payloads()
    # every payload: 68 byte-aligned slices, start codes 00 00 01 <row+1>, in order
    for f in (0, 1, 149, 150, 151, 299):
        starts = [m.start() for m in re.finditer(b"\x00\x00\x01[\x01-\x44]", pays[f])]
        rows = [pays[f][s + 3] for s in starts]
        assert [r for r in rows if r] and sorted(set(rows)) == list(range(1, 69)), f
    # oracle on a sample that straddles the internal launch rounds
    host = rgb[[0, 149, 150, 299]].cpu().numpy()
    for i, f in enumerate((0, 149, 150, 299)):
        assert pays[f] == port.encode_picture(host[i], q, MODE_FULL), f
    # batch independence: a frame encoded alone gives the same bytes as inside the batch
    one = m1.M1Encoder(W, H, 3, MODE_FULL, q, max_frames=1)
    for f in (7, 200):
        assert one.encode_device(rgb[f:f + 1]).payloads()[0] == pays[f]
    # determinism: a second pass over the same input reproduces every byte
    res2 = enc.encode_device(rgb)
    assert torch.equal(res2.frame_bytes, res.frame_bytes)
    end = int(offs[-1])
    assert hashlib.sha256(res2.out[:end].cpu().numpy().tobytes()).hexdigest() == \
           hashlib.sha256(res.out[:end].cpu().numpy().tobytes()).hexdigest()
    # host path == device path at full size
    hp, _ = enc.encode_host(rgb[:40].cpu().numpy())           # short call: simple path
    assert hp == pays[:40]
    hp, _ = enc.encode_host(rgb[:110].cpu().numpy())          # long call: pipelined upload / encode / download
    assert hp == pays[:110]
    hp, _ = enc.encode_host(rgb[:110].cpu().numpy())          # and again on the warmed-up context
    assert hp == pays[:110]


def test_config3_frame_ranges(m1, port):
    """configs[3]: frame-range sharding -- encoding [lo, hi) separately and concatenating equals one pass."""
    from ec504_imageencoder_b200.distributed import frame_range
    W, H, n, q = 1920, 1080, 24, 12
    enc = m1.M1Encoder(W, H, 3, MODE_FULL, q, max_frames=n)
    whole = enc.encode_device(enc.synth_rgb(12345, 0, n, SYNTH_NATURAL)).payloads()
    parts = []
    for r in range(8):
        lo, hi = frame_range(r, 8, n)
        parts += enc.encode_device(enc.synth_rgb(12345, lo, hi - lo, SYNTH_NATURAL)).payloads()
    assert parts == whole
