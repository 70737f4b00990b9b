"""Two-GPU test of the NVLink peer-memory gather (distributed.PeerGather): every rank's stitch kernel
writes into rank 0's memory; rank 0 must see exactly the payloads a single GPU produces for the whole
sequence.  Needs >= 2 GPUs on the box (skipped otherwise); runs examples/encode_sharded.py --peer under
torchrun and checks the assembled stream against the oracle (--verify)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("staged", [False, True])
@pytest.mark.parametrize("frames", [8, 7])
def test_peer_gather_two_gpus(tmp_path, frames, staged):
    _run_peer(tmp_path, frames, staged, False)


def test_peer_gather_device_stream(tmp_path):
    """Same exchange, and rank 0 turns the gathered segments into the file image on its GPU."""
    _run_peer(tmp_path, 7, True, True)


def _run_peer(tmp_path, frames, staged, device_stream):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "seq.mpeg"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "examples", "encode_sharded.py"), "--frames", str(frames), "--width", "352",
           "--height", "240", "--out", str(out), "--verify", "--peer"] + (["--staged"] if staged else []) \
        + (["--device-stream"] if device_stream else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_VERIFY_OK" in r.stdout
