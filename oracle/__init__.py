"""ctypes bindings for the parity checkers under oracle/.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (ec504_imageencoder_b200) never does.

  Port   -- oracle/libm1oracle.so, our C restatement (m1_oracle.c); builds anywhere.
  Ref    -- oracle/_ref/libm1ref.so, the unmodified reference sources under our generalised
            driver (ref_driver.c); built in the dev container, shipped prebuilt to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_FULL, MODE_REF_COMPAT = 0, 1
SYNTH_NATURAL, SYNTH_NOISE, SYNTH_GREY, SYNTH_RG_EQUAL, SYNTH_SCATTERED = 0, 1, 2, 3, 4

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_i16p = C.POINTER(C.c_int16)


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc).  `ref` is attempted only when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class _Base:
    prefix = ""

    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path)

    def _f(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype, fn.argtypes = restype, argtypes
        return fn

    # -- shared helpers -------------------------------------------------------------------
    def qmatrix(self, q: int) -> np.ndarray:
        out = np.zeros(64, np.int32)
        self._qmatrix(int(q), _p(out, _i32p))
        return out

    def fdct8x8(self, blk) -> np.ndarray:
        blk = _u8(blk).reshape(64)
        out = np.zeros(64, np.int32)
        self._fdct(_p(blk, _u8p), _p(out, _i32p))
        return out

    def block_bits(self, zz, is_luma: bool) -> str:
        zz = np.ascontiguousarray(zz, dtype=np.int32).reshape(64)
        buf = C.create_string_buffer(2048)
        n = self._block_bits(_p(zz, _i32p), int(bool(is_luma)), buf, 2048)
        if n < 0:
            raise ValueError(f"block not encodable by the reference (rc={n})")
        return buf.value.decode()

    def ac_table_entry(self, r: int, a: int) -> str:
        buf = C.create_string_buffer(32)
        self._ac_entry(int(r), int(a), buf)
        return buf.value.decode()

    def file_prologue(self) -> bytes:
        out = np.zeros(27, np.uint8)
        self._prologue(_p(out, _u8p))
        return out.tobytes()

    def frame_prefix(self, frame_index: int, W: int, H: int, mode: int, payload_bytes: int) -> bytes:
        out = np.zeros(44, np.uint8)
        self._prefix(int(frame_index), W, H, mode, int(payload_bytes), _p(out, _u8p))
        return out.tobytes()


class Port(_Base):
    """Our restatement (oracle/m1_oracle.c)."""
    prefix = "m1o_"

    def __init__(self):
        path = os.environ.get("M1_ORACLE_LIB") or os.path.join(HERE, "libm1oracle.so")   # M1_ORACLE_LIB: sanitizer build
        if not os.path.exists(path):
            build(ref=False)
        super().__init__(path)
        self._qmatrix = self._f("qmatrix", None, [C.c_int, _i32p])
        self._fdct = self._f("fdct8x8", None, [_u8p, _i32p])
        self._block_bits = self._f("block_bits", C.c_int, [_i32p, C.c_int, C.c_char_p, C.c_int])
        self._ac_entry = self._f("ac_table_entry", C.c_int, [C.c_int, C.c_int, C.c_char_p])
        self._prologue = self._f("file_prologue", C.c_int, [_u8p])
        self._prefix = self._f("frame_prefix", C.c_int, [C.c_long, C.c_int, C.c_int, C.c_int, C.c_long, _u8p])
        self._color = self._f("rgb_to_ycbcr", None, [_u8p, C.c_int, C.c_long, _u8p, _u8p, _u8p])
        self._sub = self._f("subsample_420", None, [_u8p, C.c_int, C.c_int, _u8p])
        self._qz = self._f("quant_zigzag", None, [_i32p, _i32p, _i32p])
        self._pic = self._f("encode_picture", C.c_long,
                            [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _u8p, C.c_long, _i16p])
        self._nmb = self._f("picture_macroblocks", C.c_long, [C.c_int, C.c_int, C.c_int])
        self._stream = self._f("encode_stream", C.c_long,
                               [_u8p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_long])
        self._synth = self._f("synth_rgb", None, [C.c_uint32, C.c_long, C.c_int, C.c_int, C.c_int, _u8p])

    def rgb_to_ycbcr(self, rgb):
        rgb = _u8(rgb)
        ch = rgb.shape[-1]
        n = rgb.size // ch
        Y, Cb, Cr = (np.zeros(n, np.uint8) for _ in range(3))
        self._color(_p(rgb, _u8p), ch, n, _p(Y, _u8p), _p(Cb, _u8p), _p(Cr, _u8p))
        return Y, Cb, Cr

    def subsample_420(self, plane, W, H):
        plane = _u8(plane)
        out = np.zeros((W // 2) * (H // 2), np.uint8)
        self._sub(_p(plane, _u8p), W, H, _p(out, _u8p))
        return out

    def quant_zigzag(self, dct, qm):
        dct = np.ascontiguousarray(dct, np.int32).reshape(64)
        qm = np.ascontiguousarray(qm, np.int32).reshape(64)
        zz = np.zeros(64, np.int32)
        self._qz(_p(dct, _i32p), _p(qm, _i32p), _p(zz, _i32p))
        return zz

    def macroblocks(self, W, H, mode=MODE_FULL) -> int:
        return int(self._nmb(W, H, mode))

    def encode_picture(self, rgb, quality=12, mode=MODE_FULL, want_levels=False, qm=None):
        """rgb: (H, W, C) uint8.  Returns payload bytes (and levels [mb,6,64] int16)."""
        rgb = _u8(rgb)
        H, W, ch = rgb.shape
        qm = self.qmatrix(quality) if qm is None else np.ascontiguousarray(qm, np.int32)
        nmb = self.macroblocks(W, H, mode)
        cap = 64 + nmb * 700
        out = np.zeros(cap, np.uint8)
        lev = np.zeros((nmb, 6, 64), np.int16) if want_levels else None
        n = self._pic(_p(rgb, _u8p), W, H, ch, mode, _p(qm, _i32p), _p(out, _u8p), cap, _p(lev, _i16p))
        if n < 0:
            raise ValueError(f"oracle encode_picture failed rc={n}")
        payload = out[:n].tobytes()
        return (payload, lev) if want_levels else payload

    def encode_stream(self, frames, quality=12, mode=MODE_FULL) -> bytes:
        """frames: (N, H, W, C) uint8 -> whole .mpeg image (trailer 00 00 01 b7)."""
        frames = _u8(frames)
        N, H, W, ch = frames.shape
        cap = 64 + N * (64 + self.macroblocks(W, H, mode) * 700)
        out = np.zeros(cap, np.uint8)
        n = self._stream(_p(frames, _u8p), N, W, H, ch, mode, int(quality), _p(out, _u8p), cap)
        if n < 0:
            raise ValueError(f"oracle encode_stream failed rc={n}")
        return out[:n].tobytes()

    def synth_rgb(self, seed, frame_index, W, H, kind=SYNTH_NATURAL):
        out = np.zeros((H, W, 3), np.uint8)
        self._synth(int(seed) & 0xFFFFFFFF, int(frame_index), W, H, kind, _p(out, _u8p))
        return out


class Ref(_Base):
    """The unmodified reference functions under oracle/ref_driver.c."""
    prefix = "m1ref_"

    def __init__(self, opt: str = "O0"):
        path = os.path.join(HERE, "_ref", "libm1ref.so" if opt == "O0" else "libm1ref_O2.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        super().__init__(path)
        self.opt = opt
        self._qmatrix = self._f("qmatrix", None, [C.c_int, _i32p])
        self._fdct = self._f("fdct8x8", None, [_u8p, _i32p])
        self._fdct_int = self._f("fdct_is_integral", C.c_int, [_u8p])
        self._block_bits = self._f("block_bits", C.c_int, [_i32p, C.c_int, C.c_char_p, C.c_int])
        self._slice_bits = self._f("slice_header_bits", C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_int])
        self._ac_entry = self._f("ac_table_entry", C.c_int, [C.c_int, C.c_int, C.c_char_p])
        self._prologue = self._f("file_prologue", C.c_int, [_u8p])
        self._prefix = self._f("frame_prefix", C.c_int, [C.c_long, C.c_int, C.c_int, C.c_int, C.c_long, _u8p])
        self._color = self._f("rgb_to_ycbcr", None, [_u8p, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _u8p])
        self._sub = self._f("subsample_420", None, [_u8p, _u8p, C.c_int, C.c_int, _u8p, _u8p])
        self._qz = self._f("quant_zigzag", None, [_i32p, C.c_int, _i32p])
        self._pic = self._f("encode_picture", C.c_long,
                            [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_long, _i16p])
        self._time = self._f("time_pictures", C.c_double,
                             [_u8p, C.c_long, C.c_int, C.c_int, C.c_int, _u8p, C.c_long, C.POINTER(C.c_long)])

    @staticmethod
    def available(opt: str = "O0") -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libm1ref.so" if opt == "O0" else "libm1ref_O2.so"))

    def rgb_to_ycbcr(self, rgb):
        rgb = _u8(rgb)
        ch = rgb.shape[-1]
        n = rgb.size // ch
        Y, Cb, Cr = (np.zeros(n, np.uint8) for _ in range(3))
        self._color(_p(rgb, _u8p), ch, n, 1, _p(Y, _u8p), _p(Cb, _u8p), _p(Cr, _u8p))
        return Y, Cb, Cr

    def subsample_420(self, cb, cr, W, H):
        cb, cr = _u8(cb), _u8(cr)
        ocb = np.zeros((W // 2) * (H // 2), np.uint8)
        ocr = np.zeros_like(ocb)
        self._sub(_p(cb, _u8p), _p(cr, _u8p), W, H, _p(ocb, _u8p), _p(ocr, _u8p))
        return ocb, ocr

    def fdct_is_integral(self, blk) -> bool:
        blk = _u8(blk).reshape(64)
        return bool(self._fdct_int(_p(blk, _u8p)))

    def quant_zigzag(self, dct, quality):
        dct = np.ascontiguousarray(dct, np.int32).reshape(64)
        zz = np.zeros(64, np.int32)
        self._qz(_p(dct, _i32p), int(quality), _p(zz, _i32p))
        return zz

    def slice_header_bits(self, quant_scale, vertical_pos) -> str:
        buf = C.create_string_buffer(256)
        self._slice_bits(quant_scale, vertical_pos, buf, 256)
        return buf.value.decode()

    def encode_picture(self, rgb, quality=12, mode=MODE_FULL, want_levels=False):
        rgb = _u8(rgb)
        H, W, ch = rgb.shape
        nmb = 54 if mode == MODE_REF_COMPAT else ((W + 15) // 16) * ((H + 15) // 16)
        cap = 64 + nmb * 700
        out = np.zeros(cap, np.uint8)
        lev = np.zeros((nmb, 6, 64), np.int16) if want_levels else None
        n = self._pic(_p(rgb, _u8p), W, H, ch, mode, int(quality), _p(out, _u8p), cap, _p(lev, _i16p))
        if n < 0:
            raise ValueError(f"reference encode_picture failed rc={n}")
        payload = out[:n].tobytes()
        return (payload, lev) if want_levels else payload

    def time_pictures(self, frames, quality=12):
        """Seconds the reference's single-threaded functions take for frames (N,H,W,3), FULL."""
        frames = _u8(frames)
        N, H, W, ch = frames.shape
        assert ch == 3
        cap = 64 + ((W + 15) // 16) * ((H + 15) // 16) * 700
        scratch = np.zeros(cap, np.uint8)
        tot = C.c_long(0)
        s = self._time(_p(frames, _u8p), N, W, H, int(quality), _p(scratch, _u8p), cap, C.byref(tot))
        if s < 0:
            raise ValueError("reference timing run failed")
        return float(s), int(tot.value)


def ref_encoder_binary() -> str | None:
    p = os.path.join(HERE, "_ref", "encoder")
    return p if os.path.exists(p) else None
