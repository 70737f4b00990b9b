"""Parity of the non-default kernel variants (selected by environment knobs that are read once per
process, hence the subprocesses): the warp-specialised persistent kernel (M1_WS=1) and other chunk
sizes (M1_CHUNK_MBS)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import oracle
from ec504_imageencoder_b200 import M1Encoder
P = oracle.Port()
for (W, H, n, q, kind) in ((352, 240, 3, 12, 0), (1920, 1080, 2, 12, 0), (640, 480, 2, 50, 1), (1920, 1080, 1, 5, 1), (48, 32, 2, 12, 1)):
    enc = M1Encoder(W, H, 3, 0, q, max_frames=n)
    rgb = enc.synth_rgb(4242, 5, n, kind)
    res = enc.encode_device(rgb, want_levels=True)
    res2 = enc.encode_device(rgb)                       # production variant (no levels)
    pay, pay2, host, lev = res.payloads(), res2.payloads(), rgb.cpu().numpy(), res.levels.cpu().numpy()
    for f in range(n):
        rp, rl = P.encode_picture(host[f], q, 0, want_levels=True)
        assert np.array_equal(lev[f], rl), (W, H, f)
        assert pay[f] == rp and pay2[f] == rp, (W, H, f)
print("VARIANT_OK")
""" % ROOT


@pytest.mark.parametrize("env", [{"M1_WS": "1"}, {"M1_CHUNK_MBS": "7"}, {"M1_CHUNK_MBS": "1"}, {"M1_WS": "1", "M1_CHUNK_MBS": "11"},
                                 {"M1_CHUNK_EVEN": "1"}, {"M1_WIN_WORDS": "8"}, {"M1_WIN_WORDS": "5", "M1_CHUNK_MBS": "3"},
                                 {"M1_WS": "1", "M1_WIN_WORDS": "8"}, {"M1_PERSIST": "1"},
                                 {"M1_PERSIST": "1", "M1_WIN_WORDS": "6", "M1_CHUNK_MBS": "5"}])
def test_kernel_variant(env):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    e = dict(os.environ, **env)
    out = subprocess.run([sys.executable, "-c", SNIPPET], env=e, capture_output=True, text=True, timeout=600)
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
