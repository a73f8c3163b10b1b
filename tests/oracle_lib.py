"""ctypes bindings for the TEST-ONLY checkers.

* ``oracle``  : oracle/_build/libzzoracle.so -- the plain-C restatement (oracle/zz_oracle.c)
* ``ref``     : oracle/_ref/libzzref.so      -- the unmodified reference behind oracle/ref_shim.cpp
                (present only where it was built, i.e. in the dev container; it travels to the GPU
                box as a prebuilt file but is never required there)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "_build" / "libzzoracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libzzref.so"
REFERENCE_DIR = Path("/root/reference")

ZLIB, GZIP, DEFLATE = 0, 1, 2
PAD = 64  # readable slack after every input handed to the checkers (reference over-reads 8 bytes)

u8p = C.POINTER(C.c_uint8)


def build_oracle(force: bool = False) -> None:
    """Compile the C restatement (and the reference shim when /root/reference exists)."""
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (ROOT / "oracle" / "zz_oracle.c").stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "oracle"])
    if REFERENCE_DIR.exists() and (force or not REF_SO.exists()
                                   or REF_SO.stat().st_mtime < (ROOT / "oracle" / "ref_shim.cpp").stat().st_mtime):
        subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "ref"])


class ChunkInfo(C.Structure):
    _fields_ = [
        ("defects", C.c_int), ("block_type", C.c_int), ("n_records", C.c_int), ("n_matches", C.c_int),
        ("block_bits", C.c_int64),
        ("lit_freq", C.c_int * 286), ("dist_freq", C.c_int * 30),
        ("lit_len", C.c_int * 286), ("dist_len", C.c_int * 30), ("meta_len", C.c_int * 19),
        ("records", C.POINTER(C.c_uint32)), ("max_records", C.c_int),
        ("matches", C.POINTER(C.c_uint32)), ("max_matches", C.c_int),
    ]


def _padded(data) -> np.ndarray:
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    buf = np.zeros(a.size + PAD, dtype=np.uint8)
    buf[: a.size] = a
    return buf


def _ptr(a: np.ndarray, off: int = 0):
    return C.cast(a.ctypes.data + off, u8p)


class Oracle:
    def __init__(self):
        build_oracle()
        self.lib = lib = C.CDLL(str(ORACLE_SO))
        lib.zzo_chunk_encode.restype = C.c_size_t
        lib.zzo_chunk_encode.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, C.c_int, u8p, C.c_size_t,
                                         C.POINTER(ChunkInfo)]
        lib.zzo_stream_chunked.restype = C.c_size_t
        lib.zzo_stream_chunked.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                           C.POINTER(C.c_int)]
        lib.zzo_stream_chunked_mt.restype = C.c_size_t
        lib.zzo_stream_chunked_mt.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_int, C.c_size_t,
                                              C.c_size_t, C.c_int]
        lib.zzo_stream_reference.restype = C.c_size_t
        lib.zzo_stream_reference.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.zzo_bound.restype = C.c_size_t
        lib.zzo_bound.argtypes = [C.c_size_t, C.c_int, C.c_size_t]
        lib.zzo_chunk_candidates.restype = None
        lib.zzo_chunk_candidates.argtypes = [u8p, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint16)]
        for name in ("zzo_adler32", "zzo_adler32x_literal"):
            f = getattr(lib, name); f.restype = C.c_uint32; f.argtypes = [C.c_uint32, u8p, C.c_size_t]
        lib.zzo_combine.restype = C.c_uint32; lib.zzo_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_size_t]
        lib.zzo_crc32.restype = C.c_uint32; lib.zzo_crc32.argtypes = [u8p, C.c_size_t, C.c_uint32]
        lib.zzo_crc32_combine.restype = C.c_uint32; lib.zzo_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        lib.zzo_calc_lengths.restype = None
        lib.zzo_calc_lengths.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.zzo_calc_lengths_iters.restype = C.c_int
        lib.zzo_calc_lengths_iters.argtypes = lib.zzo_calc_lengths.argtypes
        lib.zzo_from_lengths.restype = C.c_int
        lib.zzo_from_lengths.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), u8p, C.c_int]
        lib.zzo_generate.restype = None
        lib.zzo_generate.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int32)]
        lib.zzo_reverse.restype = C.c_uint; lib.zzo_reverse.argtypes = [C.c_uint, C.c_int]
        lib.zzo_find_distance.restype = C.c_int; lib.zzo_find_distance.argtypes = [C.c_int]
        lib.zzo_read_lut.restype = C.c_int; lib.zzo_read_lut.argtypes = [C.c_int]
        lib.zzo_hash.restype = C.c_uint; lib.zzo_hash.argtypes = [u8p]
        lib.zzo_bitstream_kat.restype = C.c_size_t
        lib.zzo_bitstream_kat.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.c_int, C.c_int, u8p, C.c_size_t]
        self.prefix = "zzo"

    # ---- streams ----
    def bound(self, n, level, chunk=65536):
        return self.lib.zzo_bound(n, level, chunk)

    def stream_chunked(self, data, fmt=ZLIB, level=2, chunk=65536, dict_size=32768, threads=1):
        src = _padded(data); n = src.size - PAD
        cap = self.bound(n, level, chunk)
        out = np.empty(cap, dtype=np.uint8)
        defects = C.c_int(0)
        if threads > 1:
            w = self.lib.zzo_stream_chunked_mt(_ptr(out), cap, _ptr(src), n, fmt, level, chunk, dict_size, threads)
        else:
            w = self.lib.zzo_stream_chunked(_ptr(out), cap, _ptr(src), n, fmt, level, chunk, dict_size, C.byref(defects))
        if w == C.c_size_t(-1).value:
            raise RuntimeError("oracle stream_chunked failed")
        return out[:w].tobytes(), defects.value

    def stream_reference(self, data, fmt=ZLIB, level=2):
        src = _padded(data); n = src.size - PAD
        cap = n + n // 8 + 1024
        out = np.empty(cap, dtype=np.uint8)
        defects = C.c_int(0)
        w = self.lib.zzo_stream_reference(_ptr(out), cap, _ptr(src), n, fmt, level, C.byref(defects))
        if w == C.c_size_t(-1).value:
            raise RuntimeError("oracle stream_reference failed")
        return out[:w].tobytes(), defects.value

    def chunk_encode(self, buf: np.ndarray, off: int, n: int, dict_size: int, level: int, final: bool,
                     want_tokens: bool = False):
        """buf: padded full input (np.uint8, >= PAD slack); chunk = buf[off:off+n]."""
        cap = n * 9 // 8 + 64
        out = np.empty(cap, dtype=np.uint8)
        info = ChunkInfo()
        recs = matches = None
        if want_tokens:
            recs = np.zeros(3 * 20000, dtype=np.uint32); matches = np.zeros(3 * 20000, dtype=np.uint32)
            info.records = recs.ctypes.data_as(C.POINTER(C.c_uint32)); info.max_records = 20000
            info.matches = matches.ctypes.data_as(C.POINTER(C.c_uint32)); info.max_matches = 20000
        w = self.lib.zzo_chunk_encode(_ptr(buf, off), n, dict_size, off, level, int(final), _ptr(out), cap, C.byref(info))
        if w == C.c_size_t(-1).value:
            raise RuntimeError("oracle chunk_encode overflow")
        res = {"bytes": out[:w].tobytes(), "defects": info.defects, "block_type": info.block_type,
               "block_bits": info.block_bits, "n_records": info.n_records, "n_matches": info.n_matches,
               "lit_freq": np.array(info.lit_freq), "dist_freq": np.array(info.dist_freq),
               "lit_len": np.array(info.lit_len), "dist_len": np.array(info.dist_len),
               "meta_len": np.array(info.meta_len)}
        if want_tokens:
            res["records"] = recs[: 3 * info.n_records].reshape(-1, 3).copy()
            res["matches"] = matches[: 3 * info.n_matches].reshape(-1, 3).copy()
        return res

    def chunk_candidates(self, buf: np.ndarray, off: int, n: int, dict_size: int) -> np.ndarray:
        cand = np.zeros(max(n, 1), dtype=np.uint16)
        self.lib.zzo_chunk_candidates(_ptr(buf, off), n, dict_size, cand.ctypes.data_as(C.POINTER(C.c_uint16)))
        return cand[:n]

    # ---- checksums ----
    def adler32(self, data, start=1):
        a = _padded(data); return self.lib.zzo_adler32(start, _ptr(a), a.size - PAD)

    def adler32x_literal(self, data, start=1):
        a = _padded(data); return self.lib.zzo_adler32x_literal(start, _ptr(a), a.size - PAD)

    def combine(self, first, second, len_second):
        return self.lib.zzo_combine(first, second, len_second)

    def crc32(self, data, start=0):
        a = _padded(data); return self.lib.zzo_crc32(_ptr(a), a.size - PAD, start)

    def crc32_combine(self, c1, c2, len2):
        return self.lib.zzo_crc32_combine(c1, c2, len2)

    # ---- huffman ----
    def calc_lengths(self, freqs, max_len, want_iters=False):
        f = (C.c_int * len(freqs))(*[int(x) for x in freqs]); out = (C.c_int * len(freqs))()
        fn = getattr(self.lib, self.prefix + "_calc_lengths_iters", None) if want_iters else None
        if fn is not None:
            it = fn(f, len(freqs), max_len, out)
            return list(out), it
        getattr(self.lib, self.prefix + "_calc_lengths")(f, len(freqs), max_len, out)
        return list(out)

    def from_lengths(self, lengths, freqs19=None):
        l = (C.c_int * len(lengths))(*[int(x) for x in lengths])
        f = (C.c_int * 19)(*(freqs19 or [0] * 19))
        rec = np.zeros(2 * (len(lengths) + 2), dtype=np.uint8)
        n = getattr(self.lib, self.prefix + "_from_lengths")(l, len(lengths), f, _ptr(rec), len(lengths) + 2)
        return [tuple(int(v) for v in rec[2 * i: 2 * i + 2]) for i in range(n)], list(f)

    def generate(self, lengths):
        l = (C.c_int * len(lengths))(*[int(x) for x in lengths]); out = (C.c_int32 * (2 * len(lengths)))()
        getattr(self.lib, self.prefix + "_generate")(l, len(lengths), out)
        return [(out[2 * i], out[2 * i + 1] & 0xFFFFFFFF) for i in range(len(lengths))]

    def tables(self):
        lib = self.lib
        lc = (C.c_int16 * 259)(); le = (C.c_int8 * 259)(); lb = (C.c_int8 * 259)()
        order = (C.c_uint8 * 19)(); ed = (C.c_uint8 * 30)(); el = (C.c_uint8 * 286)(); db = (C.c_uint16 * 30)()
        cf = (C.c_int32 * 572)(); lf = (C.c_int32 * 518)(); df = (C.c_int32 * 60)()
        getattr(lib, self.prefix + "_tables")(lc, le, lb, order, ed, el, db, cf, lf, df)
        return {"length_code": list(lc), "length_extra": list(le), "length_extra_bits": list(lb), "order": list(order),
                "extra_dist": list(ed), "extra_len": list(el), "dist_base": list(db),
                "codes_f": list(cf), "lcodes_f": list(lf), "dcodes_f": list(df)}

    def bitstream_kat(self, pairs, flush=True):
        bits = (C.c_uint64 * len(pairs))(*[p[0] for p in pairs]); cnt = (C.c_int * len(pairs))(*[p[1] for p in pairs])
        buf = np.zeros(128, dtype=np.uint8)
        n = getattr(self.lib, self.prefix + "_bitstream_kat")(bits, cnt, len(pairs), int(flush), _ptr(buf), 100)
        return buf[:n].tobytes(), buf.tobytes()


class Reference(Oracle):
    """The unmodified reference (oracle/_ref/libzzref.so).  Raises FileNotFoundError where it was not built."""

    def __init__(self):  # noqa: super().__init__ deliberately not called (different library)
        if REFERENCE_DIR.exists():
            build_oracle()
        if not REF_SO.exists():
            raise FileNotFoundError(str(REF_SO))
        self.lib = lib = C.CDLL(str(REF_SO))
        self.prefix = "zzref"
        lib.zzref_encode.restype = C.c_size_t
        lib.zzref_encode.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        lib.zzref_encode_callback.restype = C.c_size_t
        lib.zzref_encode_callback.argtypes = lib.zzref_encode.argtypes
        lib.zzref_chunk_encode.restype = C.c_size_t
        lib.zzref_chunk_encode.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, u8p, C.c_size_t]
        lib.zzref_chunk_tokens.restype = C.c_int
        lib.zzref_chunk_tokens.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, u8p, C.c_size_t,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.c_int,
                                           C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.zzref_adler32x.restype = C.c_uint32; lib.zzref_adler32x.argtypes = [C.c_uint32, u8p, C.c_size_t]
        lib.zzref_combine.restype = C.c_uint32; lib.zzref_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_size_t]
        lib.zzref_crc32.restype = C.c_uint32; lib.zzref_crc32.argtypes = [u8p, C.c_size_t, C.c_uint32]
        lib.zzref_calc_lengths.restype = None
        lib.zzref_calc_lengths.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.zzref_from_lengths.restype = C.c_int
        lib.zzref_from_lengths.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), u8p, C.c_int]
        lib.zzref_generate.restype = None
        lib.zzref_generate.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int32)]
        lib.zzref_reverse.restype = C.c_uint; lib.zzref_reverse.argtypes = [C.c_uint, C.c_int]
        lib.zzref_find_distance.restype = C.c_int; lib.zzref_find_distance.argtypes = [C.c_int]
        lib.zzref_read_lut.restype = C.c_int; lib.zzref_read_lut.argtypes = [C.c_int]
        lib.zzref_bitstream_kat.restype = C.c_size_t
        lib.zzref_bitstream_kat.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.c_int, C.c_int, u8p, C.c_size_t]

    def encode(self, data, fmt=ZLIB, level=2, threaded=False, callback=False, cap=None):
        src = _padded(data); n = src.size - PAD
        cap = cap if cap is not None else n + n // 4 + 4096
        out = np.zeros(cap + 64, dtype=np.uint8)
        fn = self.lib.zzref_encode_callback if callback else self.lib.zzref_encode
        w = fn(_ptr(out), cap, _ptr(src), n, fmt, level, int(threaded))
        if w == C.c_size_t(-1).value:
            return None
        return out[:w].tobytes()

    def chunk_encode(self, buf, off, n, dict_size, level, final, want_tokens=False):
        cap = n * 9 // 8 + 4096
        out = np.zeros(cap, dtype=np.uint8)
        if not want_tokens or level < 2:
            w = self.lib.zzref_chunk_encode(_ptr(buf, off), n, dict_size, level, int(final), _ptr(out), cap)
            return {"bytes": out[:w].tobytes()}
        recs = np.zeros(3 * 20000, dtype=np.uint32)
        lit = (C.c_int32 * 572)(); dist = (C.c_int32 * 60)(); olen = C.c_size_t(0)
        cnt = self.lib.zzref_chunk_tokens(_ptr(buf, off), n, dict_size, level, int(final), _ptr(out), cap,
                                          C.byref(olen), recs.ctypes.data_as(C.POINTER(C.c_uint32)), 20000, lit, dist)
        return {"bytes": out[: olen.value].tobytes(), "records": recs[: 3 * cnt].reshape(-1, 3).copy(),
                "lit_len": np.array(lit[0::2]), "dist_len": np.array(dist[0::2])}

    def adler32x(self, data, start=1):
        a = _padded(data); return self.lib.zzref_adler32x(start, _ptr(a), a.size - PAD)

    def combine(self, first, second, len_second):
        return self.lib.zzref_combine(first, second, len_second)

    def crc32(self, data, start=0):
        a = _padded(data); return self.lib.zzref_crc32(_ptr(a), a.size - PAD, start)


_oracle = None
_ref = None


def oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def reference() -> Reference:
    global _ref
    if _ref is None:
        _ref = Reference()
    return _ref


def have_reference() -> bool:
    try:
        reference()
        return True
    except (FileNotFoundError, OSError):
        return False
