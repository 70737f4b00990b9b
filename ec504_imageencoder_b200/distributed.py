"""Frame-range sharding across the GPUs of one box (north_star / SURVEY.md section 8e).

I-frames are independent, so a sequence of N pictures is split into contiguous ranges, one per
rank; every rank encodes its range with its own M1Encoder.  The only exchange is the gather of the
per-frame byte counts and the compressed segments to rank 0 (NCCL over NVLink on the GPU box, gloo
in the CPU tests), where they are concatenated in rank order = frame order.  No picture data and no
intermediate ever crosses GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def frame_range(rank: int, world: int, n_frames: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: rank k gets [k*N/world, (k+1)*N/world)."""
    return rank * n_frames // world, (rank + 1) * n_frames // world


@dataclass
class Gathered:
    """Rank 0's view after the gather: for every rank its per-frame sizes, 16-byte-aligned
    offsets inside its segment, and the segment bytes."""
    sizes: list            # world x int32 tensor [frames of that rank]
    offsets: list          # world x int64 tensor [frames + 1]
    segments: list         # world x uint8 tensor (segment bytes, padded between frames)

    def payloads(self) -> list[bytes]:
        out = []
        for s, o, seg in zip(self.sizes, self.offsets, self.segments):
            s, o, b = s.cpu().tolist(), o.cpu().tolist(), seg.cpu().numpy()
            out += [b[o[i]:o[i] + s[i]].tobytes() for i in range(len(s))]
        return out


def gather_to_rank0(out: torch.Tensor, frame_bytes: torch.Tensor, frame_offsets: torch.Tensor,
                    frames_per_rank: list[int], recv: torch.Tensor | None = None, group=None) -> Gathered | None:
    """Collective.  `out`/`frame_bytes`/`frame_offsets` are this rank's EncodedBatch fields;
    frames_per_rank[k] = number of frames rank k encoded (known from frame_range on every rank).
    Returns a Gathered on rank 0, None elsewhere.  `recv` (rank 0, optional) is a preallocated
    uint8 buffer for the remote segments."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = out.device
    nmax = max(frames_per_rank)
    # 1. metadata: sizes and offsets of every rank, padded to the longest range (tiny)
    meta = torch.zeros(2 * nmax + 1, dtype=torch.int64, device=dev)
    n = frames_per_rank[rank]
    meta[:n] = frame_bytes[:n].to(torch.int64)
    meta[nmax:nmax + n + 1] = frame_offsets[:n + 1]
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    # segment lengths must be known on the host to post the receives: ONE small device->host copy
    ends = torch.stack([m[nmax + frames_per_rank[k]] for k, m in enumerate(metas)]).cpu().tolist()
    # 2. payload segments: grouped send/recv into rank 0
    if rank == 0:
        need = sum(ends[1:])
        if recv is None or recv.numel() < need:
            recv = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
        ops, pos, views = [], 0, [out[:ends[0]]]
        for k in range(1, world):
            v = recv[pos:pos + ends[k]]
            views.append(v)
            if ends[k]:
                ops.append(dist.P2POp(dist.irecv, v, k, group=group))
            pos += ends[k]
    else:
        ops = [dist.P2POp(dist.isend, out[:ends[rank]], 0, group=group)] if ends[rank] else []
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    if rank != 0:
        return None
    return Gathered(sizes=[m[:frames_per_rank[k]].to(torch.int32) for k, m in enumerate(metas)],
                    offsets=[m[nmax:nmax + frames_per_rank[k] + 1] for k, m in enumerate(metas)],
                    segments=views)
