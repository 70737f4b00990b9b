"""Python host layer over the C ABI (include/m1cu.h).

PyTorch is used for plumbing only: device buffers, streams and (in bench.py) torch.distributed.
Every byte of output is produced by the CUDA kernels in csrc/ through the C ABI; there is no
CPU fallback and no oracle on this path.

Reference interface mirrored: the per-picture loop body of `mpeg_encode_procedure`
(reference include/encoder.h:216-445) and its quality argument (include/encoder.h:20,
default 12 from main.c:16).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _native

MODE_FULL = 0        # raster macroblocks, 4:2:0 chroma (BASELINE configs)
MODE_REF_COMPAT = 1  # literal traversal of include/encoder.h:238-443 (drop-in byte parity)
SYNTH_NATURAL, SYNTH_NOISE, SYNTH_GREY, SYNTH_RG_EQUAL, SYNTH_SCATTERED = 0, 1, 2, 3, 4
DEFAULT_QUALITY = 12  # reference main.c:16

_ERRORS = {-1: "bad argument", -2: "CUDA failure / no device", -3: "output capacity",
           -4: "coded AC level outside the reference's encodable range (|L| >= 256)"}


class M1Error(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"m1cu error {code} ({_ERRORS.get(code, '?')}): {detail}")


def qmatrix(quality: int) -> np.ndarray:
    """scale_quantization_matrix (reference source/image_processing.c:314-343), raster order."""
    out = np.zeros(64, np.int32)
    _native.m1cu().m1cu_qmatrix(int(quality), out.ctypes.data)
    return out


@dataclass
class EncodedBatch:
    """Device-resident result of one encode call."""
    out: torch.Tensor            # uint8, payloads at 16-byte-aligned offsets
    frame_bytes: torch.Tensor    # int32 [n] (bit pattern of uint32)
    frame_offsets: torch.Tensor  # int64 [n+1]
    levels: torch.Tensor | None  # int16 [n, macroblocks, 6, 64] or None
    out_ptr: int | None = None   # raw device address used instead of `out` (peer memory of another
    out_cap: int | None = None   # rank, see distributed.PeerGather) and its capacity in bytes

    def payloads(self) -> list[bytes]:
        """Copies the payloads to the host (synchronises)."""
        if self.out_ptr is not None:
            raise RuntimeError("this batch writes into another rank's memory; read it there (Gathered.payloads)")
        sizes = self.frame_bytes.cpu().numpy().astype(np.int64)
        offs = self.frame_offsets.cpu().numpy()
        end = int(offs[-1])
        buf = self.out[:end].cpu().numpy()
        return [buf[int(o):int(o) + int(s)].tobytes() for o, s in zip(offs[:-1], sizes)]


class M1Encoder:
    """One context per GPU.  Geometry and quality are fixed at construction."""

    def __init__(self, width: int, height: int, channels: int = 3, mode: int = MODE_FULL,
                 quality: int = DEFAULT_QUALITY, max_frames: int = 64, device: int | None = None,
                 chunk_mbs: int = 0, chunk_even: bool = False, win_words: int = 0, no_tail_pairing: bool = False,
                 batch_frames: int = 0, no_flat_skip: bool = False):
        self.lib = _native.m1cu()
        if self.lib.m1cu_device_count() == 0 or not torch.cuda.is_available():
            raise M1Error(-2, "no CUDA device: the encode path has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.width, self.height, self.channels = int(width), int(height), int(channels)
        self.mode, self.quality, self.max_frames = int(mode), int(quality), int(max_frames)
        h = C.c_void_p()
        # work-partition knobs (m1cu_tuning; tests and sweeps only, no effect on the bytes)
        tuning = (C.c_int * 6)(int(chunk_mbs), int(bool(chunk_even)), int(win_words), int(bool(no_tail_pairing)), int(batch_frames),
                               int(bool(no_flat_skip)))
        rc = self.lib.m1cu_create_ex(C.byref(h), self.device, self.width, self.height, self.channels,
                                     self.mode, self.quality, self.max_frames, tuning)
        if rc != 0:
            raise M1Error(rc, (self.lib.m1cu_last_error(None) or b"").decode())
        self._h = h
        self.macroblocks = self.lib.m1cu_macroblocks_per_frame(h)
        self.flat_range = self.lib.m1cu_flat_range(h)        # blocks spanning at most this skip the DCT (-1: none)
        self.frame_bytes_in = self.lib.m1cu_frame_bytes_in(h)
        self.payload_bound = self.lib.m1cu_payload_bound(h)
        self._stream = None

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.m1cu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        raise M1Error(rc, (self.lib.m1cu_last_error(self._h) or b"").decode())

    def _bind_stream(self):
        # torch's default stream has handle 0 (the legacy default stream); the C ABI reads NULL as
        # "use your own stream", so name the legacy stream explicitly (cudaStreamLegacy == 0x1).
        s = torch.cuda.current_stream(self.device).cuda_stream or 1
        if s != self._stream:
            rc = self.lib.m1cu_set_stream(self._h, C.c_void_p(s))
            if rc:
                self._err(rc)
            self._stream = s

    @property
    def launches(self) -> int:
        return int(self.lib.m1cu_launch_count(self._h))

    def enable_timing(self, on: bool = True):
        self.lib.m1cu_enable_timing(self._h, int(bool(on)))

    def kernel_times(self):
        """(ms[3], launches[3]) for k_encode_chunks / k_layout / k_stitch since the last call."""
        ms = (C.c_double * 3)()
        n = (C.c_ulonglong * 3)()
        rc = self.lib.m1cu_kernel_times(self._h, ms, n)
        if rc:
            self._err(rc)
        return list(ms), [int(x) for x in n]

    def typical_out_bytes(self, n_frames: int) -> int:
        return int(self.lib.m1cu_typical_out_bytes(self._h, int(n_frames)))

    # -- device-resident hot path ------------------------------------------------------------
    def alloc_outputs(self, n_frames: int, want_levels: bool = False, out_bytes: int | None = None) -> EncodedBatch:
        dev = torch.device("cuda", self.device)
        cap = self.typical_out_bytes(n_frames) if out_bytes is None else int(out_bytes)
        return EncodedBatch(
            out=torch.empty(cap, dtype=torch.uint8, device=dev),
            frame_bytes=torch.empty(n_frames, dtype=torch.int32, device=dev),
            frame_offsets=torch.empty(n_frames + 1, dtype=torch.int64, device=dev),
            levels=(torch.empty((n_frames, self.macroblocks, 6, 64), dtype=torch.int16, device=dev)
                    if want_levels else None))

    def encode_device(self, rgb: torch.Tensor, res: EncodedBatch | None = None, want_levels: bool = False,
                      check: bool = True) -> EncodedBatch:
        """rgb: uint8 CUDA tensor [n, H, W, C] (contiguous).  Asynchronous unless check=True."""
        assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        n = rgb.shape[0]
        assert tuple(rgb.shape[1:]) == (self.height, self.width, self.channels), rgb.shape
        if res is None:
            res = self.alloc_outputs(n, want_levels)
        self._bind_stream()
        out_ptr = res.out.data_ptr() if res.out_ptr is None else res.out_ptr
        out_cap = res.out.numel() if res.out_ptr is None else res.out_cap
        rc = self.lib.m1cu_encode_device(self._h, rgb.data_ptr(), n, out_ptr, out_cap,
                                         res.frame_bytes.data_ptr(), res.frame_offsets.data_ptr(),
                                         res.levels.data_ptr() if res.levels is not None else None)
        if rc:
            self._err(rc)
        if check:
            self.check()
        return res

    def push_payloads(self, res: EncodedBatch, dst_ptr: int, dst_cap: int, n_frames: int | None = None):
        """Copies the payload bytes of a finished encode to `dst_ptr` (e.g. a peer mapping) with a small
        kernel on torch's CURRENT stream; the caller orders that stream after the encode."""
        s = torch.cuda.current_stream(self.device).cuda_stream or 1
        rc = self.lib.m1cu_push_payloads(self._h, C.c_void_p(s), dst_ptr, int(dst_cap), res.out.data_ptr(),
                                         res.frame_offsets.data_ptr(),
                                         int(res.frame_bytes.numel() if n_frames is None else n_frames))
        if rc:
            self._err(rc)

    def assemble_stream(self, res: EncodedBatch, n_frames: int | None = None, first_frame_index: int = 0,
                        prologue: bool = True, out: torch.Tensor | None = None, offset: int = 0):
        """The bytes of the .mpeg file for a finished encode, assembled on the device (m1cu_assemble_stream):
        [27-byte prologue] + per picture 44-byte prefix, payload, 4-byte trailer.  The header bytes come from the
        host C library (hostlib.stream_templates).  Writing starts at byte `offset` of `out` (continue a stream
        with the end offset of the previous call).  Returns (uint8 CUDA tensor, the end offset as a 1-element
        int64 CUDA tensor); asynchronous on the encoder's stream."""
        from . import hostlib
        n = int(res.frame_bytes.numel() if n_frames is None else n_frames)
        if getattr(self, "_tmpl", None) is None:
            self._tmpl = hostlib.stream_templates(self.width, self.height, self.mode)
        prefix, prolog, trailer = self._tmpl
        if out is None:
            cap = int(offset) + 32 + n * 48 + int(res.out.numel() if res.out_ptr is None else res.out_cap)
            out = torch.empty((cap + 15) // 16 * 16, dtype=torch.uint8, device=res.frame_bytes.device)
        nbytes = torch.zeros(1, dtype=torch.int64, device=res.frame_bytes.device)
        self._bind_stream()
        out_ptr = res.out.data_ptr() if res.out_ptr is None else res.out_ptr
        rc = self.lib.m1cu_assemble_stream(self._h, out_ptr, res.frame_bytes.data_ptr(), res.frame_offsets.data_ptr(), n,
                                           int(first_frame_index), prefix.ctypes.data,
                                           prolog.ctypes.data if prologue else None, trailer.ctypes.data,
                                           out.data_ptr(), out.numel(), int(offset), nbytes.data_ptr())
        if rc:
            self._err(rc)
        return out, nbytes

    def check(self):
        rc = self.lib.m1cu_check(self._h)
        if rc:
            self._err(rc)

    # -- host-buffer path (what the C driver uses) ---------------------------------------------
    def encode_host(self, rgb: np.ndarray | torch.Tensor, want_levels: bool = False, out: np.ndarray | None = None,
                    copy: bool = True):
        """rgb: uint8 host array/tensor [n, H, W, C].  Returns (payload list, levels or None).
        copy=False returns numpy views into the output buffer instead of bytes objects."""
        if isinstance(rgb, torch.Tensor):
            assert not rgb.is_cuda
            src_ptr, n, keep = rgb.data_ptr(), rgb.shape[0], rgb
            assert rgb.is_contiguous() and rgb.dtype == torch.uint8
        else:
            keep = np.ascontiguousarray(rgb, dtype=np.uint8)
            src_ptr, n = keep.ctypes.data, keep.shape[0]
        assert tuple(keep.shape[1:]) == (self.height, self.width, self.channels), keep.shape
        cap = self.typical_out_bytes(n)
        for _ in range(2):
            buf = out if out is not None and out.size >= cap else np.empty(cap, np.uint8)
            sizes = np.zeros(n, np.uint32)
            lev = np.empty((n, self.macroblocks, 6, 64), np.int16) if want_levels else None
            total = C.c_size_t(0)
            self._bind_stream()
            rc = self.lib.m1cu_encode_host(self._h, src_ptr, n, buf.ctypes.data, buf.size, sizes.ctypes.data,
                                           lev.ctypes.data if lev is not None else None, C.byref(total))
            if rc == -3 and cap < self.payload_bound * n:
                cap = self.payload_bound * n
                out = None
                continue
            if rc:
                self._err(rc)
            break
        offs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
        if copy:
            payloads = [buf[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)]
        else:
            payloads = [buf[int(offs[i]):int(offs[i + 1])] for i in range(n)]
        return payloads, lev

    # -- utilities ---------------------------------------------------------------------------------
    def synth_rgb(self, seed: int, first_frame: int, n_frames: int, kind: int = SYNTH_NATURAL) -> torch.Tensor:
        dev = torch.device("cuda", self.device)
        t = torch.empty((n_frames, self.height, self.width, 3), dtype=torch.uint8, device=dev)
        self._bind_stream()
        rc = self.lib.m1cu_synth_rgb(self._h, int(seed) & 0xFFFFFFFF, int(first_frame), int(n_frames), int(kind),
                                     t.data_ptr())
        if rc:
            self._err(rc)
        return t

    def ycbcr_planes(self, rgb: torch.Tensor):
        """Full-resolution Y, Cb, Cr planes of ONE picture [H, W, C] (the .bit side-file content)."""
        assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        dev = rgb.device
        planes = [torch.empty((self.height, self.width), dtype=torch.uint8, device=dev) for _ in range(3)]
        self._bind_stream()
        rc = self.lib.m1cu_ycbcr_planes(self._h, rgb.data_ptr(), *[p.data_ptr() for p in planes])
        if rc:
            self._err(rc)
        return planes
