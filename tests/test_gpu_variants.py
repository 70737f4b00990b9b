"""Parity of the non-default work partitions of k_encode_chunks (m1cu_create_ex / m1cu_tuning: chunk sizes,
equal chunks, small bit windows that force the multi-window path).  The product library reads no
environment variable; the knobs are arguments."""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (1080p: 120 macroblocks per slice = 7 x 16 + 8, SIF: 22 = 16 + 6, 640: 40 = 2 x 16 + 8 -> paired last chunks by default;
#  352x240 has an odd number of slices, so its last slice's short chunk stays alone; 89 = long blocks: multi-pass windows)
CASES = ((352, 240, 3, 12, 0), (1920, 1080, 2, 12, 0), (640, 480, 2, 50, 1), (1920, 1080, 1, 5, 1), (48, 32, 2, 12, 1),
         (1920, 1088, 1, 89, 1), (352, 256, 2, 89, 1))


@pytest.mark.parametrize("tuning", [{"chunk_mbs": 7}, {"chunk_mbs": 1}, {"chunk_mbs": 11}, {"chunk_even": True},
                                    {"win_words": 8}, {"win_words": 5, "chunk_mbs": 3}, {"win_words": 6, "chunk_mbs": 5},
                                    {"batch_frames": 1}, {"batch_frames": 2, "chunk_mbs": 5},
                                    {"no_tail_pairing": True}, {"no_tail_pairing": True, "win_words": 8},
                                    {"chunk_mbs": 14}, {"chunk_mbs": 14, "win_words": 4}, {"chunk_mbs": 9, "win_words": 7},
                                    {"no_flat_skip": True}, {"no_flat_skip": True, "chunk_mbs": 5, "win_words": 6}])
def test_kernel_variant(tuning, port):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import M1Encoder
    for (W, H, n, q, kind) in CASES:
        enc = M1Encoder(W, H, 3, 0, q, max_frames=n, **tuning)
        rgb = enc.synth_rgb(4242, 5, n, kind)
        res = enc.encode_device(rgb, want_levels=True)
        res2 = enc.encode_device(rgb)                       # production variant (no levels)
        pay, pay2, host, lev = res.payloads(), res2.payloads(), rgb.cpu().numpy(), res.levels.cpu().numpy()
        for f in range(n):
            rp, rl = port.encode_picture(host[f], q, 0, want_levels=True)
            assert np.array_equal(lev[f], rl), (W, H, f)
            assert pay[f] == rp and pay2[f] == rp, (W, H, f)
        enc.close()


@pytest.mark.parametrize("quality", [5, 12, 20, 50, 89])
def test_flat_block_skip_at_the_boundary(port, quality):
    """Blocks whose samples span at most m1_flat_range(quality) skip the DCT (only their DC coefficient is computed, the
    others are queued and transformed by the whole CTA).  Pictures whose 8x8 blocks span R - 3 .. R + 3 grey levels around
    random bases (so passes and failures mix inside every warp, and whole macroblocks of each kind occur), plus colour
    noise of the same amplitudes: levels and bytes against the oracle, with and without the shortcut."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import M1Encoder
    R = {5: 42, 12: 16, 20: 10, 50: 2, 89: 0}[quality]      # m1_flat_range; the library uses it from 6 up (quality 50, 89: no shortcut)
    W, H, n = 640, 368, 4
    rng = np.random.default_rng(quality)
    frames = np.zeros((n, H, W, 3), np.uint8)
    for f in range(n):
        by, bx = H // 8, W // 8
        span = rng.integers(max(R - 3, 0), R + 4, (by, bx))
        if f == 1:
            span = np.repeat(np.repeat(rng.integers(max(R - 3, 0), R + 4, (by // 2, bx // 2)), 2, 0), 2, 1)   # per macroblock
        base = rng.integers(0, 256 - (R + 4), (by, bx))
        px = base.repeat(8, 0).repeat(8, 1) + (rng.random((H, W)) * (span.repeat(8, 0).repeat(8, 1) + 1)).astype(np.int64)
        if f < 2:
            frames[f] = px[:, :, None]                                             # grey: Y ~ the pattern, chroma flat
        else:
            frames[f] = np.clip(px[:, :, None] + rng.integers(0, R // 2 + 2, (H, W, 3)), 0, 255)
    want = [port.encode_picture(frames[f], quality, 0, want_levels=True) for f in range(n)]
    for nfs in (False, True):
        enc = M1Encoder(W, H, 3, 0, quality, max_frames=n, no_flat_skip=nfs)
        rgb = torch.from_numpy(frames).cuda()
        res = enc.encode_device(rgb, want_levels=True)
        res2 = enc.encode_device(rgb)
        pay, pay2, lev = res.payloads(), res2.payloads(), res.levels.cpu().numpy()
        for f in range(n):
            assert np.array_equal(lev[f], want[f][1]), (quality, nfs, f)
            assert pay[f] == want[f][0] and pay2[f] == want[f][0], (quality, nfs, f)
        enc.close()


def test_flat_range_is_reported():
    """m1cu_flat_range: the bound of csrc/m1cu_quant.h per quality (checked against the oracle in tests/test_block_host.py); the
    library uses it from 6 grey levels up and not at all with no_flat_skip."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import M1Encoder
    for q, want in ((1, 218), (5, 42), (12, 16), (20, 10), (30, 6), (40, -1), (50, -1), (89, -1)):
        enc = M1Encoder(64, 48, 3, 0, q, max_frames=1)
        assert enc.flat_range == want, (q, enc.flat_range)
        enc.close()
    enc = M1Encoder(64, 48, 3, 0, 12, max_frames=1, no_flat_skip=True)
    assert enc.flat_range == -1
    enc.close()


def test_random_geometries_qualities_contents(port):
    """Sixty random combinations of picture size (aligned and ragged: fast and generic colour loads), quality (1 .. 80: flat
    ranges from 218 grey levels down to none), content kind (all five generators) and chunk size: levels and bytes against the
    oracle.  Every block path of the encoder (DC only, eight lanes per block, one thread per block) occurs many times over."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import M1Encoder
    rng = np.random.default_rng(20261019)
    for case in range(60):
        W = int(rng.choice([16 * int(rng.integers(1, 26)), int(rng.integers(1, 400))]))
        H = int(rng.choice([16 * int(rng.integers(1, 14)), int(rng.integers(1, 220))]))
        q = int(rng.integers(1, 81))
        kind = int(rng.integers(0, 5))
        cm = int(rng.choice([0, 0, int(rng.integers(1, 17))]))
        n = int(rng.integers(1, 4))
        enc = M1Encoder(W, H, 3, 0, q, max_frames=n, chunk_mbs=cm)
        rgb = enc.synth_rgb(1000 + case, case, n, kind)
        res = enc.encode_device(rgb, want_levels=True)
        pay2 = enc.encode_device(rgb).payloads()
        pay, host, lev = res.payloads(), rgb.cpu().numpy(), res.levels.cpu().numpy()
        for f in range(n):
            rp, rl = port.encode_picture(host[f], q, 0, want_levels=True)
            assert np.array_equal(lev[f], rl), (case, W, H, q, kind, cm, f)
            assert pay[f] == rp and pay2[f] == rp, (case, W, H, q, kind, cm, f)
        enc.close()


def test_launch_rounds_overlap(port):
    """A call with more pictures than one launch round holds (batch_frames) runs the layout + stitch of a round on a
    side stream beside the next round's chunk encoder, on two sets of staging buffers: 23 pictures in rounds of 4,
    twice in a row on the same context (buffer reuse across calls), levels and bytes against the oracle; the
    host-buffer paths on top of it."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import M1Encoder
    W, H, n, q = 352, 240, 23, 12
    enc = M1Encoder(W, H, 3, 0, q, max_frames=n, batch_frames=4)
    rgb = enc.synth_rgb(31, 0, n, 0)
    host = rgb.cpu().numpy()
    want = [port.encode_picture(host[f], q, 0, want_levels=True) for f in range(n)]
    for _ in range(2):
        res = enc.encode_device(rgb, want_levels=True)
        pay, lev = res.payloads(), res.levels.cpu().numpy()
        for f in range(n):
            assert np.array_equal(lev[f], want[f][1]), f
            assert pay[f] == want[f][0], f
    res_a = enc.alloc_outputs(n)
    res_b = enc.alloc_outputs(9)
    enc.encode_device(rgb, res=res_a, check=False)            # two calls queued back to back, no host sync between
    enc.encode_device(rgb[:9], res=res_b, check=False)
    enc.check()
    assert res_a.payloads() == [w[0] for w in want]
    assert res_b.payloads() == [w[0] for w in want[:9]]
    hp, _ = enc.encode_host(host)
    assert hp == [w[0] for w in want]
    enc.close()


def test_tuning_arguments_are_validated():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import M1Encoder, M1Error
    for bad in ({"chunk_mbs": 17}, {"chunk_mbs": -1}, {"win_words": 3}, {"win_words": 513}):
        with pytest.raises(M1Error):
            M1Encoder(64, 64, 3, 0, 12, max_frames=1, **bad)


INT_SNIPPET = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import oracle
from ec504_imageencoder_b200 import M1Encoder
P = oracle.Port()
g = np.arange(1 << 24, dtype=np.uint32)
base = np.stack([g >> 16, (g >> 8) & 255, g & 255], axis=1).astype(np.uint8)
enc = M1Encoder(4096, 4096, 3, 0, 89, max_frames=1)
for shift in range(4):                                   # every colour in each of the four byte alignments
    img = np.ascontiguousarray(np.roll(base, shift, axis=0).reshape(1, 4096, 4096, 3))
    res = enc.encode_device(torch.from_numpy(img).cuda(), want_levels=True)
    rp, rl = P.encode_picture(img[0], 89, 0, want_levels=True)
    assert np.array_equal(res.levels[0].cpu().numpy(), rl), shift
    assert res.payloads()[0] == rp, shift
enc.close()
for (W, H, n, q, kind) in ((352, 240, 2, 12, 2), (1920, 1080, 1, 50, 3), (100, 70, 2, 89, 2), (1920, 1080, 2, 12, 0), (641, 479, 2, 12, 1)):
    enc = M1Encoder(W, H, 3, 0, q, max_frames=n)
    rgb = enc.synth_rgb(4242, 5, n, kind)
    res = enc.encode_device(rgb, want_levels=True)
    pay, host, lev = res.payloads(), rgb.cpu().numpy(), res.levels.cpu().numpy()
    for f in range(n):
        rp, rl = P.encode_picture(host[f], q, 0, want_levels=True)
        assert np.array_equal(lev[f], rl) and pay[f] == rp, (W, H, kind, f)
print("VARIANT_OK")
""" % ROOT


@pytest.mark.parametrize("variant", ["intcolour", "mixcolour"])
def test_integer_colour_build_variants(variant):
    """The integer colour path (m1cu_colour.cuh, -DM1_COLOUR_SPLIT=0 / 1) is not the product's arithmetic (it measured
    slower) but stays bit-exact: all 2^24 colours in every byte alignment through the encode
    kernel and its fix-up queue, plus grey / r == g pictures where every quad is queued."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    lib = os.path.join(ROOT, "build_variants", f"libm1cu_{variant}.so")
    if not os.path.exists(lib):
        pytest.skip(f"{lib} not built (tools/build_experiments.sh {variant})")
    out = subprocess.run([sys.executable, "-c", INT_SNIPPET], env=dict(os.environ, M1CU_LIB=lib), capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "VARIANT_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("device_stream", [False, True])
def test_sharded_example_matches_oracle(tmp_path, device_stream):
    """examples/encode_sharded.py: frame-range sharding + NCCL gather to rank 0 + host headers (or, with
    --device-stream, the file image assembled on rank 0's GPU) gives the oracle's byte stream.  Runs with 2
    ranks when the box has 2 GPUs, else with 1."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    n = 2 if torch.cuda.device_count() >= 2 else 1
    out = str(tmp_path / "s.mpeg")
    script = os.path.join(ROOT, "examples", "encode_sharded.py")
    args = ["--frames", "7", "--width", "352", "--height", "240", "--out", out, "--verify"]
    if device_stream:
        args.append("--device-stream")
    if n == 1:
        cmd = [sys.executable, script] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
               "--master-addr", "127.0.0.1", "--master-port", "29591", script] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_VERIFY_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
