"""Synthetic inputs of SURVEY 8(d): order-2 Markov text, splitmix64 random bytes, repetitive buffers.

Bench / test support.  The generators are C (csrc/zz_synth.c, built by build.py into libzzsynth.so)
because the Markov chain is sequential per 1 MiB segment; the trained model ships as data/markov2.npz
(made by tools/make_markov_model.py from the reference's English corpus files).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
SYNTH_LIB = PKG_DIR / "libzzsynth.so"
MODEL = PKG_DIR / "data" / "markov2.npz"

SEED_TEXT = 0x5EED0001
SEED_RANDOM = 0x5EED0002
SEED_PATTERN = 0x5EED0003

_lib = None
_model = None


def _load():
    global _lib, _model
    if _lib is None:
        if not SYNTH_LIB.exists():
            from . import build as _b
            _b.build_synth()
        _lib = C.CDLL(str(SYNTH_LIB))
        _lib.zz_synth_markov.restype = None
        _lib.zz_synth_markov.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint8, C.c_uint8, C.c_int]
        _lib.zz_synth_random.restype = None
        _lib.zz_synth_random.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int]
    if _model is None:
        m = np.load(MODEL)
        row_off = np.ascontiguousarray(m["row_off"], dtype=np.uint32)
        counts = m["counts"].astype(np.uint64)
        total = np.cumsum(counts)
        row_start_total = np.concatenate([[0], total])[row_off[:-1]]          # cumulative count before each row
        row_of_entry = np.repeat(np.arange(65536), np.diff(row_off.astype(np.int64)))
        cum = (total - row_start_total[row_of_entry]).astype(np.uint32)       # inclusive, per row
        _model = (row_off, np.ascontiguousarray(m["syms"], dtype=np.uint8), np.ascontiguousarray(cum),
                  int(m["start"][0]), int(m["start"][1]))
    return _lib, _model


def _threads(threads):
    return threads if threads else max(1, (os.cpu_count() or 1))


def markov_text(n: int, seg0: int = 0, threads: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """n bytes of the text-like stream starting at 1 MiB segment `seg0` (n, seg0 shard a longer stream)."""
    lib, (row_off, syms, cum, s0, s1) = _load()
    dst = out if out is not None else np.empty(n, dtype=np.uint8)
    assert dst.size >= n and dst.dtype == np.uint8 and dst.flags.c_contiguous
    lib.zz_synth_markov(dst.ctypes.data, n, seg0, row_off.ctypes.data, syms.ctypes.data, cum.ctypes.data, s0, s1,
                        _threads(threads))
    return dst[:n]


def random_bytes(n: int, seed: int = SEED_RANDOM, threads: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    lib, _ = _load()
    dst = out if out is not None else np.empty(n, dtype=np.uint8)
    lib.zz_synth_random(dst.ctypes.data, n, seed, _threads(threads))
    return dst[:n]


def repetitive(n: int, kind: str = "pattern") -> np.ndarray:
    """'zeros' or a 1000-byte splitmix64 pattern (seed 0x5EED0003) repeated."""
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    pat = random_bytes(1000, SEED_PATTERN, threads=1)
    reps = (n + 999) // 1000
    return np.tile(pat, reps)[:n].copy()


def workload(name: str, n: int, seg0: int = 0) -> np.ndarray:
    if name == "text":
        return markov_text(n, seg0)
    if name == "random":
        return random_bytes(n)
    if name == "zeros":
        return repetitive(n, "zeros")
    if name == "pattern":
        return repetitive(n, "pattern")
    raise ValueError(name)
