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
            s, o = s.cpu().tolist(), o.cpu().tolist()
            if not s:
                continue
            b = seg[:o[-1]].cpu().numpy()
            out += [b[o[i]:o[i] + s[i]].tobytes() for i in range(len(s))]
        return out


    def assemble_stream(self, enc, first_frame_index: int = 0, prologue: bool = True) -> torch.Tensor:
        """Rank 0: the ordered .mpeg image of all ranks' pictures, assembled on the device
        (M1Encoder.assemble_stream / m1cu_assemble_stream, one call per rank's segment, each continuing
        where the previous one ended).  Returns the uint8 CUDA tensor holding exactly the stream bytes."""
        from .encoder import EncodedBatch
        counts = [int(s.numel()) for s in self.sizes]
        ends = [int(o[-1].item()) if c else 0 for o, c in zip(self.offsets, counts)]     # one small sync per rank
        cap = 32 + sum(48 * c + e for c, e in zip(counts, ends))
        out = torch.empty((cap + 15) // 16 * 16, dtype=torch.uint8, device=self.segments[0].device)
        off, index, first = 0, int(first_frame_index), True
        for s, o, seg, c in zip(self.sizes, self.offsets, self.segments, counts):
            if c == 0:
                continue
            b = EncodedBatch(out=seg, frame_bytes=s.contiguous(), frame_offsets=o.contiguous(), levels=None)
            _, end = enc.assemble_stream(b, n_frames=c, first_frame_index=index, prologue=prologue and first,
                                         out=out, offset=off)
            off, index, first = int(end.item()), index + c, False
        enc.check()
        return out[:off]


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


class _RawCuda:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class PeerGather:
    """Fused stitch + gather over NVLink peer memory (NCCL process group, one process per GPU).

    Rank 0 owns `slots` x `world` receive regions of `region_bytes`; it exports the allocation as a
    CUDA IPC handle and every other rank maps it.  `batch(slot)` returns an EncodedBatch whose `out`
    IS this rank's region in rank 0's memory, so `M1Encoder.encode_device(rgb, res=batch)` makes
    k_stitch write the finished payload bytes straight to rank 0 -- the compressed segments are never
    copied again.  `finish(batch)` is the per-step collective: one small all_gather of the per-frame
    sizes and offsets, which doubles as the completion fence (it is stream-ordered after every rank's
    k_stitch, and a kernel's peer writes have landed when it completes).

    Slot reuse: rank 0 must enqueue whatever consumes slot s (e.g. the device->host copy) on the
    stream it calls `finish` from BEFORE its next `finish`; the peers cannot leave that collective, and
    therefore cannot start overwriting a slot, until rank 0 has entered it.  With `slots` = 2 that
    gives rank 0 one full step to consume a result.

    `staged=True` keeps k_stitch local and lets `finish` push the payload bytes to rank 0 with a small
    copy kernel on the CURRENT stream before the fence.  Called from a high-priority side stream that
    waits on the encode, the push overlaps the next step's encode; the direct mode cannot (at 8 GPUs
    all seven remote stitches share rank 0's NVLink ingest, ~0.25 ms per 165 MB, on the critical path).
    Direct = one pass fewer over the payload and the lowest latency for a single batch; staged = the
    better steady-state throughput.

    CUDA refuses to open an IPC handle in the exporting process, so this needs world >= 2 real
    processes; `gather_to_rank0` is the backend-neutral path (gloo in the CPU tests)."""

    def __init__(self, enc, n_frames: int, slots: int = 2, region_bytes: int | None = None, group=None,
                 staged: bool = False):
        from . import _native
        self.lib = _native.m1cu()
        self.enc, self.n, self.slots, self.group = enc, int(n_frames), int(slots), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = enc.device
        self.staged = bool(staged)
        self.region = int(region_bytes) if region_bytes else enc.typical_out_bytes(self.n)
        self.region = (self.region + 255) & ~255
        total = self.region * self.world * self.slots
        import ctypes as C
        handle = [None]
        self._base, self._mapped = None, None
        if self.rank == 0:
            self._base = self.lib.m1cu_device_alloc(total)
            if not self._base:
                raise MemoryError(f"PeerGather: cannot allocate {total} bytes on rank 0")
            buf = C.create_string_buffer(64)
            if self.lib.m1cu_ipc_export(self._base, buf):
                raise RuntimeError("PeerGather: " + (self.lib.m1cu_last_error(None) or b"").decode())
            handle[0] = buf.raw
        dist.broadcast_object_list(handle, src=0, group=group)
        if self.rank == 0:
            base = self._base
        else:
            p = C.c_void_p()
            if self.lib.m1cu_ipc_open(self.device, handle[0], C.byref(p)):
                raise RuntimeError("PeerGather: " + (self.lib.m1cu_last_error(None) or b"").decode())
            self._mapped = p.value
            base = p.value
        dev = torch.device("cuda", self.device)
        # rank 0 sees its regions as an ordinary tensor; the other ranks only ever pass the mapped peer
        # address to the C ABI (torch would attribute that pointer to rank 0's device)
        self._all = torch.as_tensor(_RawCuda(base, total), device=dev) if self.rank == 0 else None
        self._batches = []
        for s in range(self.slots):
            off = (s * self.world + self.rank) * self.region
            if self.rank == 0:
                b = enc.alloc_outputs(self.n, out_bytes=16)
                b.out = self._all[off:off + self.region]
            elif self.staged:
                b = enc.alloc_outputs(self.n, out_bytes=self.region)     # local; pushed by finish()
            else:
                b = enc.alloc_outputs(self.n, out_bytes=16)
                b.out_ptr, b.out_cap = base + off, self.region
            self._batches.append(b)
        self._peer = [base + (s * self.world + self.rank) * self.region for s in range(self.slots)]
        self._meta = [torch.zeros(2 * self.n + 1, dtype=torch.int64, device=dev) for _ in range(self.slots)]
        self._metas = [[torch.empty(2 * self.n + 1, dtype=torch.int64, device=dev) for _ in range(self.world)]
                       for _ in range(self.slots)]

    def batch(self, slot: int):
        """EncodedBatch of this rank for `slot`; its `out` lives in rank 0's memory."""
        return self._batches[slot % self.slots]

    def finish(self, slot: int, frames_per_rank: list[int] | None = None):
        """Collective, stream-ordered after this rank's encode into `slot`.  Returns a Gathered on rank 0
        (views, no host synchronisation), None elsewhere.  frames_per_rank[k] = pictures rank k encoded
        (default: n_frames everywhere)."""
        s = slot % self.slots
        b, meta, n = self._batches[s], self._meta[s], self.n
        cnt = frames_per_rank if frames_per_rank is not None else [n] * self.world
        if self.staged and self.rank != 0 and cnt[self.rank] > 0:
            self.enc.push_payloads(b, self._peer[s], self.region, cnt[self.rank])
        meta[:n] = b.frame_bytes
        meta[n:] = b.frame_offsets
        dist.all_gather(self._metas[s], meta, group=self.group)
        if self.rank != 0:
            return None
        segs = [self._all[(s * self.world + k) * self.region:(s * self.world + k + 1) * self.region]
                for k in range(self.world)]
        return Gathered(sizes=[m[:cnt[k]].to(torch.int32) for k, m in enumerate(self._metas[s])],
                        offsets=[m[n:n + cnt[k] + 1] for k, m in enumerate(self._metas[s])], segments=segs)

    def close(self):
        """Collective: every rank unmaps before rank 0 frees."""
        self._batches, self._all = [], None
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        if self._mapped:
            self.lib.m1cu_ipc_close(self.device, self._mapped)
            self._mapped = None
        dist.barrier(group=self.group)
        if self._base:
            self.lib.m1cu_device_free(self._base)
            self._base = None
