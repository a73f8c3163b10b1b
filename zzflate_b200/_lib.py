"""ctypes binding of the C-ABI in include/zzgpu.h (libzzflate_b200.so, built in-tree by build.py).

The library has no CPU fallback: loading fails loudly when the shared object is missing, and every
compute entry point returns ZZGPU_E_NO_DEVICE (raised here as ZzGpuError) without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
# ZZFLATE_B200_LIB selects a diagnostic build of the same library (e.g. -DZZ_PHASE_TIMING); never a fallback
LIB_PATH = Path(os.environ.get("ZZFLATE_B200_LIB") or (PKG_DIR / "libzzflate_b200.so"))

OK, E_NO_DEVICE, E_CUDA, E_ARG, E_CAPACITY, E_NOMEM = range(6)
MEM_HOST, MEM_DEVICE = 0, 1

u8p = C.POINTER(C.c_uint8)


class ZzGpuError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"zzgpu status {status}: {message}")
        self.status = status


class Stats(C.Structure):
    _fields_ = [
        ("chunks", C.c_uint64), ("stored_chunks", C.c_uint64), ("matches", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("device_ms", C.c_float), ("total_ms", C.c_float),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("stage_ms", C.c_float * 10), ("stage_launches", C.c_uint32 * 10),
    ]

STAGES = ["cand", "parse", "huff", "offs", "emit", "cksum", "fixed", "gather", "info", "lz"]


# every symbol include/zzgpu.h declares; tests check that the library exports all of them
SYMBOLS = [
    "zzgpu_init", "zzgpu_shutdown", "zzgpu_device_count", "zzgpu_strerror", "zzgpu_last_error", "zzgpu_bound",
    "zzgpu_deflate", "zzgpu_deflate_ex", "zzgpu_checksums", "zzgpu_adler32_combine", "zzgpu_crc32_combine",
    "zzgpu_debug_chunk", "zzgpu_set_option", "zzgpu_get_counter", "zzgpu_deflate_hist", "zzgpu_deflate_sink",
    "zzgpu_deflate_hold", "zzgpu_fetch", "zzgpu_release", "zzgpu_deflate_mode",
]

SINK_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_size_t, C.c_void_p)

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "zzflate_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    lib.zzgpu_init.restype = C.c_int; lib.zzgpu_init.argtypes = [C.c_int]
    lib.zzgpu_shutdown.restype = None
    lib.zzgpu_device_count.restype = C.c_int
    lib.zzgpu_strerror.restype = C.c_char_p; lib.zzgpu_strerror.argtypes = [C.c_int]
    lib.zzgpu_last_error.restype = C.c_char_p
    lib.zzgpu_bound.restype = C.c_size_t; lib.zzgpu_bound.argtypes = [C.c_size_t, C.c_int, C.c_uint32]
    lib.zzgpu_deflate.restype = C.c_int
    lib.zzgpu_deflate.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int,
                                  C.c_int, C.c_uint32, C.c_uint32,
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(Stats)]
    lib.zzgpu_deflate_ex.restype = C.c_int
    lib.zzgpu_deflate_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                     C.c_void_p, C.c_size_t, C.c_int,
                                     C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(Stats)]
    lib.zzgpu_deflate_mode.restype = C.c_int
    lib.zzgpu_deflate_mode.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                       C.c_void_p, C.c_size_t, C.c_int,
                                       C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(Stats)]
    lib.zzgpu_checksums.restype = C.c_int
    lib.zzgpu_checksums.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.zzgpu_adler32_combine.restype = C.c_uint32
    lib.zzgpu_adler32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_size_t]
    lib.zzgpu_crc32_combine.restype = C.c_uint32
    lib.zzgpu_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    lib.zzgpu_set_option.restype = C.c_int; lib.zzgpu_set_option.argtypes = [C.c_char_p, C.c_int]
    lib.zzgpu_get_counter.restype = C.c_longlong; lib.zzgpu_get_counter.argtypes = [C.c_char_p]
    lib.zzgpu_deflate_hist.restype = C.c_int
    lib.zzgpu_deflate_hist.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                       C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_size_t)]
    lib.zzgpu_deflate_sink.restype = C.c_int
    lib.zzgpu_deflate_sink.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                       SINK_FN, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(Stats)]
    lib.zzgpu_deflate_hold.restype = C.c_int
    lib.zzgpu_deflate_hold.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(Stats)]
    lib.zzgpu_fetch.restype = C.c_int
    lib.zzgpu_fetch.argtypes = [C.c_void_p, C.c_size_t, SINK_FN, C.c_void_p, C.c_size_t]
    lib.zzgpu_release.restype = None
    lib.zz_c_partition.restype = C.c_int
    lib.zz_c_partition.argtypes = [C.c_size_t, C.c_int, C.c_uint32, C.POINTER(C.c_uint64), C.c_int]
    lib.zzgpu_debug_chunk.restype = C.c_int
    lib.zzgpu_debug_chunk.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64,
                                      C.POINTER(C.c_uint16), C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint32), u8p, C.POINTER(C.c_uint32)]
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != OK:
        lib = load()
        msg = lib.zzgpu_strerror(status).decode()
        detail = lib.zzgpu_last_error().decode()
        raise ZzGpuError(status, f"{msg} ({detail})" if detail else msg)
