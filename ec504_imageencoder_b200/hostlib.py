"""ctypes view of libencoder.so: the reference's `make sharedlib` API (host C) plus the driver.

`mpeg_encode_procedure` and `encode_frames_*` run the per-picture hot path on the GPU through
libm1cu.so; the per-stage functions are host-C compatibility entry points (see include/*.h)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native

_lib = None


class BitVector(C.Structure):
    """struct bitvector (include/bit_vector.h)."""
    _fields_ = [("value", C.POINTER(C.c_char)), ("bits", C.c_longlong), ("cursor", C.c_longlong), ("cap", C.c_longlong)]

    def bitstring(self) -> str:
        raw = C.string_at(self.value, (self.cap + 7) // 8)
        return "".join(str((raw[k >> 3] >> (7 - (k & 7))) & 1) for k in range(self.cap))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    _native.m1cu()                                   # libencoder.so links against libm1cu.so
    path = os.environ.get("M1_HOSTLIB") or _native.LIB_ENCODER      # M1_HOSTLIB: the sanitizer build (tools/run_host_sanitizers.sh)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `make sharedlib`")
    L = C.CDLL(path)
    bvp, u8p, ip = C.POINTER(BitVector), C.c_void_p, C.c_void_p
    sig = {
        "mpeg_encode_procedure": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
        "m1_encode_frames_to_file": (C.c_int, [C.c_char_p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "m1_encode_frames_to_memory": (C.c_long, [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_long]),
        "m1_stream_templates": (None, [C.c_int, C.c_int, C.c_int, u8p, u8p, u8p]),
        "bitvector_new": (bvp, [C.c_char_p, C.c_longlong]),
        "bitvector_put_bit": (None, [bvp, C.c_char]),
        "bitvector_put_binstring": (None, [bvp, C.c_char_p]),
        "bitvector_put_byte_off": (None, [bvp, C.c_ubyte, C.c_char, C.c_char]),
        "bitvector_put_byte": (None, [bvp, C.c_char, C.c_char]),
        "bitvector_put_byte_ent": (None, [bvp, C.c_char]),
        "bitvector_concat": (None, [bvp, bvp]),
        "bitvector_clone": (bvp, [bvp]),
        "bitvector_pos": (C.c_longlong, [bvp, C.c_longlong]),
        "bitvector_toarray": (C.c_int, [bvp, C.c_char_p]),
        "fast_DCT": (None, [u8p, C.c_void_p]),
        "scale_quantization_matrix": (None, [ip, C.c_int]),
        "quantization": (None, [C.c_void_p, ip, C.c_int]),
        "zigzag_scanning": (None, [ip, ip]),
        "equalize_coefficients": (None, [ip, ip]),
        "run_length_encode": (C.c_void_p, [ip, ip]),
        "encode_block_header_i": (None, [C.c_ubyte, ip, bvp]),
        "encode_block_end": (None, [bvp]),
        "encode_macroblock_header_i": (None, [C.c_uint, C.c_short, bvp]),
        "mpeg1_slice": (None, [C.c_uint8, C.c_uint8, bvp]),
        "encode_blk_coeff": (bvp, [C.c_int, C.c_int, C.c_int]),
        "encode_macblk_address_value": (bvp, [C.c_int]),
        "mpeg1_file_header": (None, [C.c_uint32, u8p]),
        "mpeg1_sys_header": (None, [C.c_uint32, C.c_uint8, u8p]),
        "mpeg1_packet_header": (None, [C.c_uint32, u8p]),
        "mpeg1_sequence_header": (None, [C.c_uint16, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8, u8p]),
        "mpeg1_sequence_end": (None, [u8p]),
        "mpeg1_gop": (None, [C.c_uint8] * 7 + [u8p]),
        "mpeg1_picture_header": (None, [C.c_uint16, C.c_uint8, C.c_uint16, u8p, u8p]),
        "extract_8x8_block": (None, [u8p, C.c_int, C.c_int, C.c_int, u8p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def encode_frames_to_memory(frames: np.ndarray, quality: int = 12, mode: int = 0) -> bytes:
    """frames: uint8 [n, H, W, C] -> the whole .mpeg stream image (GPU encode, host headers)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, H, W, ch = frames.shape
    cap = 64 + n * (64 + ((W + 15) // 16) * ((H + 15) // 16) * 700)
    out = np.empty(cap, np.uint8)
    got = lib().m1_encode_frames_to_memory(frames.ctypes.data, n, W, H, ch, int(quality), int(mode), out.ctypes.data, cap)
    if got < 0:
        raise RuntimeError(f"m1_encode_frames_to_memory failed ({got})")
    return out[:got].tobytes()


def stream_templates(width: int, height: int, mode: int = 0):
    """(prefix256 [256*44], prologue [27], trailer [4]) as uint8 arrays: the header bytes, written by the host
    C library's reference-API functions, that M1Encoder.assemble_stream hands to the device."""
    prefix, prologue, trailer = np.zeros(256 * 44, np.uint8), np.zeros(27, np.uint8), np.zeros(4, np.uint8)
    lib().m1_stream_templates(int(width), int(height), int(mode), prefix.ctypes.data, prologue.ctypes.data, trailer.ctypes.data)
    return prefix, prologue, trailer


def mpeg_encode_procedure(images_folder: str, bitstream_folder: str, video_path: str, quality_factor: int = 12) -> int:
    """The reference's entry point (include/encoder.h:20): same arguments and return codes."""
    return int(lib().mpeg_encode_procedure(images_folder.encode(), bitstream_folder.encode(), video_path.encode(),
                                           int(quality_factor)))
