"""Loader / builder for the in-tree native libraries.

  libm1cu.so      CUDA kernels + the C ABI of include/m1cu.h           (nvcc, sm_100a)
  libencoder.so   host C: the reference's `make sharedlib` API + mpeg_encode_procedure,
                  calling the CUDA path through m1cu.h only            (gcc)

Both are built in-tree (they travel to the GPU box with the snapshot).  There is no CPU
fallback: if libm1cu.so is missing or no GPU is present, compute calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_M1CU = os.path.join(PKG, "libm1cu.so")
LIB_ENCODER = os.path.join(PKG, "libencoder.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
CUDA_SOURCES = ["m1cu_kernels.cu", "m1cu_api.cu"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libm1cu.so")


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in ("m1cu_common.cuh", "m1cu_kernels.h", "m1cu_tables.h", "m1cu_block.cuh",
                                                   "m1cu_quant.h", "m1cu_colour.cuh")]
    deps.append(os.path.join(ROOT, "include", "m1cu.h"))
    if not force and _newer(LIB_M1CU, deps):
        return LIB_M1CU
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_M1CU, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_M1CU


def build_host(force: bool = False) -> str:
    """gcc build of the host C library (reference API + driver) via the top-level Makefile."""
    subprocess.run(["make", "-s", "-C", ROOT, "sharedlib"] + (["-B"] if force else []), check=True)
    return LIB_ENCODER


_m1cu = None


def m1cu() -> C.CDLL:
    """The CUDA library with typed entry points.  Raises if it has not been built."""
    global _m1cu
    if _m1cu is not None:
        return _m1cu
    path = os.environ.get("M1CU_LIB") or LIB_M1CU       # M1CU_LIB: a tools/build_experiments.sh variant (tests only)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the encode path)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, u8p, i32p, u32p, u64p, i16p = (C.c_void_p,) * 6
    sig = {
        "m1cu_abi_version": (C.c_int, []),
        "m1cu_device_count": (C.c_int, []),
        "m1cu_qmatrix": (C.c_int, [C.c_int, i32p]),
        "m1cu_last_error": (C.c_char_p, [vp]),
        "m1cu_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "m1cu_create_ex": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "m1cu_destroy": (C.c_int, [vp]),
        "m1cu_set_stream": (C.c_int, [vp, vp]),
        "m1cu_synchronize": (C.c_int, [vp]),
        "m1cu_macroblocks_per_frame": (C.c_int, [vp]),
        "m1cu_flat_range": (C.c_int, [vp]),
        "m1cu_frame_bytes_in": (C.c_size_t, [vp]),
        "m1cu_payload_bound": (C.c_size_t, [vp]),
        "m1cu_typical_out_bytes": (C.c_size_t, [vp, C.c_int]),
        "m1cu_encode_device": (C.c_int, [vp, u8p, C.c_int, u8p, C.c_size_t, u32p, u64p, i16p]),
        "m1cu_check": (C.c_int, [vp]),
        "m1cu_encode_host": (C.c_int, [vp, u8p, C.c_int, u8p, C.c_size_t, u32p, i16p, C.POINTER(C.c_size_t)]),
        "m1cu_ycbcr_planes": (C.c_int, [vp, u8p, u8p, u8p, u8p]),
        "m1cu_host_batch_planes": (C.c_int, [vp, C.c_int, u8p, C.c_size_t]),
        "m1cu_synth_rgb": (C.c_int, [vp, C.c_uint32, C.c_long, C.c_int, C.c_int, u8p]),
        "m1cu_launch_count": (C.c_ulonglong, [vp]),
        "m1cu_enable_timing": (C.c_int, [vp, C.c_int]),
        "m1cu_kernel_times": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
        "m1cu_device_alloc": (vp, [C.c_size_t]),
        "m1cu_device_free": (None, [vp]),
        "m1cu_pinned_alloc": (vp, [C.c_size_t]),
        "m1cu_pinned_free": (None, [vp]),
        "m1cu_pinned_alloc_wc": (vp, [C.c_size_t]),
        "m1cu_memcpy_h2d": (C.c_int, [vp, vp, C.c_size_t]),
        "m1cu_memcpy_d2h": (C.c_int, [vp, vp, C.c_size_t]),
        "m1cu_ipc_export": (C.c_int, [vp, C.c_char_p]),
        "m1cu_ipc_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
        "m1cu_ipc_close": (C.c_int, [C.c_int, vp]),
        "m1cu_push_payloads": (C.c_int, [vp, vp, u8p, C.c_size_t, u8p, u64p, C.c_int]),
        "m1cu_encode_host_stream": (C.c_int, [vp, u8p, C.c_int, C.c_long, vp, vp, vp, u8p, C.c_size_t, C.POINTER(C.c_size_t)]),
        "m1cu_assemble_stream": (C.c_int, [vp, u8p, u32p, u64p, C.c_int, C.c_long, vp, vp, vp, u8p, C.c_size_t, C.c_size_t, u64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._m1_symbols = tuple(sig)
    _m1cu = lib
    return lib


M1CU_SYMBOLS = (
    "m1cu_abi_version", "m1cu_device_count", "m1cu_qmatrix", "m1cu_last_error", "m1cu_create", "m1cu_create_ex",
    "m1cu_destroy", "m1cu_set_stream", "m1cu_synchronize", "m1cu_macroblocks_per_frame", "m1cu_flat_range",
    "m1cu_frame_bytes_in", "m1cu_payload_bound", "m1cu_typical_out_bytes", "m1cu_encode_device",
    "m1cu_check", "m1cu_encode_host", "m1cu_ycbcr_planes", "m1cu_host_batch_planes", "m1cu_synth_rgb", "m1cu_launch_count",
    "m1cu_enable_timing", "m1cu_kernel_times",
    "m1cu_device_alloc", "m1cu_device_free", "m1cu_pinned_alloc", "m1cu_pinned_free", "m1cu_pinned_alloc_wc",
    "m1cu_memcpy_h2d", "m1cu_memcpy_d2h",
    "m1cu_ipc_export", "m1cu_ipc_open", "m1cu_ipc_close", "m1cu_push_payloads", "m1cu_assemble_stream",
    "m1cu_encode_host_stream",
)
