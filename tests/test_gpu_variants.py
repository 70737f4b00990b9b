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

CASES = ((352, 240, 3, 12, 0), (1920, 1080, 2, 12, 0), (640, 480, 2, 50, 1), (1920, 1080, 1, 5, 1), (48, 32, 2, 12, 1))


@pytest.mark.parametrize("tuning", [{"chunk_mbs": 7}, {"chunk_mbs": 1}, {"chunk_mbs": 11}, {"chunk_even": True},
                                    {"win_words": 8}, {"win_words": 5, "chunk_mbs": 3}, {"win_words": 6, "chunk_mbs": 5}])
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


def test_tuning_arguments_are_validated():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from ec504_imageencoder_b200 import M1Encoder, M1Error
    for bad in ({"chunk_mbs": 17}, {"chunk_mbs": -1}, {"win_words": 3}, {"win_words": 513}):
        with pytest.raises(M1Error):
            M1Encoder(64, 64, 3, 0, 12, max_frames=1, **bad)


def test_product_library_reads_no_environment():
    """no getenv in the product build of the CUDA library."""
    import re
    lib = open(os.path.join(ROOT, "ec504_imageencoder_b200", "libm1cu.so"), "rb").read()
    for knob in (b"M1_DEBUG_SKIP", b"M1_PAD_SMEM", b"M1_CHUNK_MBS", b"M1_WIN_WORDS", b"M1_WS", b"M1_PERSIST", b"M1_CHUNK_EVEN"):
        assert knob not in lib, knob


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
