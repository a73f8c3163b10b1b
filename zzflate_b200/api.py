"""Python mirror of the reference's public interface (zzflate.h / encoder.h / crc.h) over the C-ABI.

Names, argument meaning and error behaviour follow the reference so that tests read like its own
(zztest/Test.cpp): ``ZzFlateEncode`` returns ``None`` where the C++ call reports ``*destLen = ~0``.
Everything computes on the GPU through libzzflate_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, Stats, ZzGpuError, check

DEFAULT_CHUNK = 65536
DEFAULT_DICT = 32768
_ERR = C.c_size_t(-1).value


class Format(enum.IntEnum):       # zzflate.h:8
    Zlib = 0
    Gzip = 1
    Deflate = 2


@dataclass
class Config:                     # zzflate.h:10-15
    format: Format = Format.Zlib
    level: int = 2
    threaded: bool = False


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def bound(n: int, level: int = 2, chunk: int = DEFAULT_CHUNK) -> int:
    """Worst-case raw deflate size for n input bytes (zzgpu_bound)."""
    return _lib.load().zzgpu_bound(n, level, chunk)


def ZzFlateEncode(source, config: Config, dest_len: Optional[int] = None) -> Optional[bytes]:
    """zzflate.cpp:225 -- returns the stream, or None when the C++ API reports ``*destLen = ~0``."""
    lib = _lib.load()
    src = _as_u8(source)
    cap = dest_len if dest_len is not None else bound(src.size, min(config.level, 3)) + 32
    dest = np.empty(max(cap, 1), dtype=np.uint8)
    lib.zz_c_encode.restype = C.c_size_t
    lib.zz_c_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
    w = lib.zz_c_encode(dest.ctypes.data, cap, src.ctypes.data, src.size, int(config.format), int(config.level),
                        int(config.threaded))
    if w == _ERR:
        return None
    return dest[:w].tobytes()


def ZzFlateEncodeToCallback(source, config: Config, callback: Callable[[bytes], bool]) -> None:
    """zzflate.cpp:197 -- header, data pieces, trailer are delivered in order; return value ignored."""
    lib = _lib.load()
    src = _as_u8(source)
    CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_size_t, C.c_void_p)

    def tramp(ptr, count, _user):
        callback(C.string_at(ptr, count))
        return 0

    cb = CB(tramp)
    lib.zz_c_encode_to_callback.restype = None
    lib.zz_c_encode_to_callback.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, CB, C.c_void_p]
    lib.zz_c_encode_to_callback(src.ctypes.data, src.size, int(config.format), int(config.level), int(config.threaded),
                                cb, None)


def adler32x(start_value: int, data) -> int:          # adler.cpp:17
    lib = _lib.load()
    a = _as_u8(data)
    lib.zz_c_adler32x.restype = C.c_uint32
    lib.zz_c_adler32x.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
    return lib.zz_c_adler32x(start_value, a.ctypes.data, a.size)


def combine(first: int, second: int, len_second: int) -> int:      # adler.cpp:5
    return _lib.load().zzgpu_adler32_combine(first, second, len_second)


def crc32(data, start_value: int = 0) -> int:          # crc.cpp:24
    lib = _lib.load()
    a = _as_u8(data)
    lib.zz_c_crc32.restype = C.c_uint32
    lib.zz_c_crc32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    return lib.zz_c_crc32(a.ctypes.data, a.size, start_value)


def crc32_combine(crc1: int, crc2: int, len2: int) -> int:
    return _lib.load().zzgpu_crc32_combine(crc1, crc2, len2)


def deflate_raw(source, level: int = 2, chunk: int = DEFAULT_CHUNK, dict_size: int = DEFAULT_DICT,
                history: int = 0, final: bool = True, checksums: bool = True, mode: int = 0):
    """zzgpu_deflate_ex on host buffers.  `source` holds `history` bytes of preceding stream first.
    Returns (bytes, adler_start0, crc, Stats)."""
    lib = _lib.load()
    src = _as_u8(source)
    n = src.size - history
    cap = bound(n, level, chunk) + 16
    dst = np.empty(cap, dtype=np.uint8)
    out_len = C.c_size_t(0); a0 = C.c_uint32(0); crc = C.c_uint32(0); st = Stats()
    check(lib.zzgpu_deflate_mode(src.ctypes.data + history, n, history, int(final), MEM_HOST, dst.ctypes.data, cap, MEM_HOST,
                                 level, chunk, dict_size, 3 if checksums else 0, mode,
                                 C.byref(out_len), C.byref(a0), C.byref(crc), C.byref(st)))
    return dst[: out_len.value].tobytes(), a0.value, crc.value, st


def deflate_device(src_ptr: int, n: int, dst_ptr: int, cap: int, level: int = 2, chunk: int = DEFAULT_CHUNK,
                   dict_size: int = DEFAULT_DICT, history: int = 0, final: bool = True, checksums: int = 0):
    """zzgpu_deflate_ex on device-resident buffers (raw pointers, e.g. torch.Tensor.data_ptr()).
    Returns (out_len, adler_start0, crc, Stats)."""
    lib = _lib.load()
    out_len = C.c_size_t(0); a0 = C.c_uint32(0); crc = C.c_uint32(0); st = Stats()
    check(lib.zzgpu_deflate_ex(src_ptr, n, history, int(final), MEM_DEVICE, dst_ptr, cap, MEM_DEVICE,
                               level, chunk, dict_size, checksums,
                               C.byref(out_len), C.byref(a0), C.byref(crc), C.byref(st)))
    return out_len.value, a0.value, crc.value, st


def debug_chunk(source, chunk_index: int, level: int = 2, chunk: int = DEFAULT_CHUNK, dict_size: int = DEFAULT_DICT):
    """Intermediate per-chunk products of the level>=2 pipeline (zzgpu_debug_chunk)."""
    lib = _lib.load()
    src = _as_u8(source)
    cand = np.zeros(chunk, dtype=np.uint16)
    tokens = np.zeros(3 * 20000, dtype=np.uint32)
    ntok = C.c_uint32(0)
    hist = np.zeros(316, dtype=np.uint32)
    lengths = np.zeros(336, dtype=np.uint8)
    info = np.zeros(4, dtype=np.uint32)
    check(lib.zzgpu_debug_chunk(src.ctypes.data, src.size, MEM_HOST, level, chunk, dict_size, chunk_index,
                                cand.ctypes.data_as(C.POINTER(C.c_uint16)),
                                tokens.ctypes.data_as(C.POINTER(C.c_uint32)), 20000, C.byref(ntok),
                                hist.ctypes.data_as(C.POINTER(C.c_uint32)),
                                lengths.ctypes.data_as(C.POINTER(C.c_uint8)),
                                info.ctypes.data_as(C.POINTER(C.c_uint32))))
    return {"cand": cand, "matches": tokens[: 3 * ntok.value].reshape(-1, 3).copy(), "hist": hist,
            "lit_len": lengths[:286].astype(int), "dist_len": lengths[286:316].astype(int),
            "meta_len": lengths[316:335].astype(int),
            "block_type": int(info[0]), "hdr_bits": int(info[1]), "out_bytes": int(info[2]), "total_bits": int(info[3])}


def ZzFlateDecode(stream, fmt: Format = Format.Zlib, max_len: Optional[int] = None, dictionary=None) -> Optional[bytes]:
    """include/decoder.h -- the host inflater that fills the role of the reference's stub decoder.h.  Returns the
    inflated bytes, or None on any error (bad header / block / code / distance / checksum, truncated input)."""
    lib = _lib.load()
    src = _as_u8(stream)
    cap = max_len if max_len is not None else max(1024, src.size * 64)
    dest = np.empty(max(cap, 1), dtype=np.uint8)
    d = _as_u8(dictionary) if dictionary is not None and len(dictionary) else None
    lib.zz_c_decode.restype = C.c_size_t
    lib.zz_c_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
    status = C.c_int(0)
    w = lib.zz_c_decode(dest.ctypes.data, cap, src.ctypes.data, src.size, int(fmt), d.ctypes.data if d is not None else None,
                        d.size if d is not None else 0, C.byref(status))
    if w == _ERR:
        return None
    return dest[:w].tobytes()


def device_count() -> int:
    return _lib.load().zzgpu_device_count()


def encode_ptr(dest_ptr: int, cap: int, src_ptr: int, n: int, config: Config) -> Optional[int]:
    """ZzFlateEncode on caller-owned HOST buffers given as raw addresses (e.g. pinned torch tensors).
    Returns bytes written or None on error (*destLen = ~0)."""
    lib = _lib.load()
    lib.zz_c_encode.restype = C.c_size_t
    lib.zz_c_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
    w = lib.zz_c_encode(dest_ptr, cap, src_ptr, n, int(config.format), int(config.level), int(config.threaded))
    return None if w == _ERR else w
