#!/usr/bin/env python3
"""Frame-range sharded encode of one sequence on N GPUs of one box (BASELINE configs[3]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        examples/encode_sharded.py --frames 64 --width 1920 --height 1080 --out /tmp/seq.mpeg [--verify]

Every rank encodes its contiguous frame range on its own GPU; per-frame byte counts and the
compressed segments are gathered to rank 0 over NCCL (or, with --peer, written by every rank's
stitch kernel straight into rank 0's memory over NVLink: distributed.PeerGather); rank 0 adds the host-side headers (with the
GLOBAL frame index, which drives the time stamps) and writes the .mpeg.  --verify compares the file
with the oracle's stream for the same synthetic frames (test infrastructure, CPU, slow).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ec504_imageencoder_b200 import M1Encoder, MODE_FULL, SYNTH_NATURAL, hostlib  # noqa: E402
from ec504_imageencoder_b200.distributed import Gathered, PeerGather, frame_range, gather_to_rank0  # noqa: E402


def frame_prefix(L, index, W, H, payload_bytes):
    """The 44 bytes in front of a picture's payload, built with libencoder's header functions
    (reference include/encoder.h:196-231 + the length patch :448-454)."""
    out = np.zeros(44, np.uint8)
    hour = index & 0xFF
    L.mpeg1_packet_header(1 + 3600 * hour, out.ctypes.data)
    L.mpeg1_sequence_header(W, H, 1, 4, 3, out[16:].ctypes.data)
    L.mpeg1_gop(0, hour, 0, 0, 0, 1, 0, out[28:].ctypes.data)
    bid = np.zeros(4, np.uint8)
    L.mpeg1_picture_header(0, 1, 0xFFFF, bid.ctypes.data, out[36:].ctypes.data)
    fwd = (44 + payload_bytes - 8) & 0xFFFF
    out[4], out[5] = fwd >> 8, fwd & 0xFF
    return out.tobytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--quality", type=int, default=12)
    ap.add_argument("--out", default="/tmp/m1_sharded.mpeg")
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--peer", action="store_true", help="payloads go to rank 0 through NVLink peer memory (PeerGather)")
    ap.add_argument("--staged", action="store_true", help="with --peer: local stitch + push kernel instead of a remote stitch")
    ap.add_argument("--device-stream", action="store_true",
                    help="rank 0 assembles the final file image on its GPU (m1cu_assemble_stream) instead of on the host")
    a = ap.parse_args()

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = frame_range(rank, world, a.frames)
    counts = [frame_range(r, world, a.frames)[1] - frame_range(r, world, a.frames)[0] for r in range(world)]
    enc = M1Encoder(a.width, a.height, 3, MODE_FULL, a.quality, max_frames=max(1, hi - lo), device=local)
    rgb = enc.synth_rgb(12345, lo, max(1, hi - lo), SYNTH_NATURAL)[: hi - lo]      # this rank's pictures
    pg = None
    if world > 1 and a.peer:
        pg = PeerGather(enc, max(1, max(counts)), slots=1, staged=a.staged)
        res = pg.batch(0)
        if hi > lo:
            enc.encode_device(rgb.contiguous(), res=res)
        g = pg.finish(0, counts)
    elif world > 1:
        res = enc.encode_device(rgb.contiguous()) if hi > lo else enc.alloc_outputs(1)
        g = gather_to_rank0(res.out, res.frame_bytes, res.frame_offsets, counts)
    else:
        res = enc.encode_device(rgb.contiguous())
        g = Gathered(sizes=[res.frame_bytes], offsets=[res.frame_offsets], segments=[res.out])
    image = payloads = None
    if rank == 0 and a.device_stream:
        image = g.assemble_stream(enc).cpu().numpy().tobytes()      # headers from libencoder, bytes placed by the GPU
    elif rank == 0:
        payloads = g.payloads()

    ok = True
    if rank == 0 and image is not None:
        with open(a.out, "wb") as f:
            f.write(image)
    elif rank == 0:
        L = hostlib.lib()
        pro = np.zeros(27, np.uint8)
        L.mpeg1_file_header(2202035, pro.ctypes.data)
        L.mpeg1_sys_header(2202035, 0xE6, pro[12:].ctypes.data)
        with open(a.out, "wb") as f:
            f.write(pro.tobytes())
            for i, p in enumerate(payloads):
                f.write(frame_prefix(L, i, a.width, a.height, len(p)))
                f.write(p)
                f.write(b"\x00\x00\x01\xb7")
    if rank == 0:
        size = os.path.getsize(a.out)
        print(f"rank 0: {a.frames} frames from {world} GPU(s) -> {a.out} ({size} bytes)")
        if a.verify:
            import oracle
            P = oracle.Port()
            frames = np.stack([P.synth_rgb(12345, i, a.width, a.height, SYNTH_NATURAL) for i in range(a.frames)])
            ok = open(a.out, "rb").read() == P.encode_stream(frames, a.quality, MODE_FULL)
            print("SHARDED_VERIFY_OK" if ok else "SHARDED_VERIFY_MISMATCH")
    if pg is not None:
        pg.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
