"""N > 1 host logic on the CPU: frame-range partition and the gather of sizes + compressed segments
to rank 0, world_size 2 and 3 over gloo.  The payloads here come from the oracle (this is a test of
the plumbing, not of the encoder)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def test_frame_range_partition():
    from ec504_imageencoder_b200.distributed import frame_range
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 300, 8000, 8001):
            ranges = [frame_range(r, world, n) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from ec504_imageencoder_b200.distributed import frame_range, gather_to_rank0
        P = oracle.Port()
        W, H = 64, 48
        lo, hi = frame_range(rank, world, n_frames)
        pays = [P.encode_picture(P.synth_rgb(5, f, W, H, 1), 50, 0) for f in range(lo, hi)]
        # lay the payloads out like an EncodedBatch: 16-byte-aligned offsets
        offs, pos = [], 0
        for p in pays:
            offs.append(pos)
            pos += (len(p) + 15) & ~15
        offs.append(pos)
        buf = np.zeros(max(pos, 1) + 64, np.uint8)
        for o, p in zip(offs, pays):
            buf[o:o + len(p)] = np.frombuffer(p, np.uint8)
        g = gather_to_rank0(torch.from_numpy(buf), torch.tensor([len(p) for p in pays], dtype=torch.int32),
                            torch.tensor(offs, dtype=torch.int64),
                            [frame_range(r, world, n_frames)[1] - frame_range(r, world, n_frames)[0] for r in range(world)])
        if rank == 0:
            want = [P.encode_picture(P.synth_rgb(5, f, W, H, 1), 50, 0) for f in range(n_frames)]
            q.put(g.payloads() == want)
        else:
            assert g is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 7), (3, 4), (2, 1)])
def test_gather_to_rank0_gloo(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_frame_range_partitions_baseline_configs():
    """BASELINE configs[3] (8000 frames over 2 / 4 / 8 ranks) and configs[4] (120 frames over 8): the ranges are
    contiguous, disjoint, ordered and cover the sequence; uneven counts differ by at most one frame."""
    from ec504_imageencoder_b200.distributed import frame_range
    for total, worlds in ((8000, (1, 2, 4, 8)), (120, (2, 4, 8)), (7, (2, 3, 8)), (0, (2,))):
        for w in worlds:
            edges = [frame_range(k, w, total) for k in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[k][1] == edges[k + 1][0] for k in range(w - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
