"""include/decoder.h: the host inflater in the place of the reference's stub (zzflate/decoder.h:4-33).  Host code, so it is
checked without a GPU: against zlib's own streams (stored / fixed / dynamic blocks, all levels), against the oracle's
E-mode streams of the golden inputs (every format and level), with a dictionary (a shard of a longer stream), and on
damaged input."""
import zlib

import numpy as np
import pytest

from oracle_lib import DEFLATE, GZIP, ZLIB


@pytest.fixture(scope="module")
def zz():
    import zzflate_b200
    from zzflate_b200 import build
    build.build()
    return zzflate_b200


def _inputs(golden):
    rng = np.random.default_rng(5)
    yield b""
    yield b"a"
    yield b"hello hello hello hello"
    yield bytes(100000)
    yield rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    yield (rng.integers(0, 4, 50000, dtype=np.uint8) + 65).astype(np.uint8).tobytes()
    for case in ("alice29", "kennedy", "ptt5", "mixed"):
        yield golden.input(case)


def test_inflates_zlib_streams(zz, golden):
    for data in _inputs(golden):
        for level in (0, 1, 6, 9):
            for fmt, wbits in ((zz.Format.Zlib, 15), (zz.Format.Gzip, 31), (zz.Format.Deflate, -15)):
                c = zlib.compressobj(level, zlib.DEFLATED, wbits)
                stream = c.compress(data) + c.flush()
                assert zz.ZzFlateDecode(stream, fmt, max_len=len(data) + 16) == data, (len(data), level, fmt)
        c = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)        # fixed Huffman blocks
        stream = c.compress(data) + c.flush()
        assert zz.ZzFlateDecode(stream, zz.Format.Deflate, max_len=len(data) + 16) == data


def test_inflates_the_oracles_streams(zz, oracle, golden):
    for case in ("alice29", "kennedy", "ptt5", "pattern", "zeros", "random", "mixed", "hello", "one", "zero512"):
        data = golden.input(case)
        for level in (0, 1, 2):
            for fmt, ofmt in ((zz.Format.Zlib, ZLIB), (zz.Format.Gzip, GZIP), (zz.Format.Deflate, DEFLATE)):
                stream, _ = oracle.stream_chunked(data, ofmt, level)
                assert zz.ZzFlateDecode(stream, fmt, max_len=len(data) + 16) == data, (case, level, fmt)


def test_dictionary_and_members(zz, oracle, golden):
    data = golden.input("markov")
    whole, _ = oracle.stream_chunked(data, DEFLATE, 2)
    # a shard of a longer stream: its matches reach into the history before it
    c = zlib.compressobj(6, zlib.DEFLATED, -15, zdict=data[:40000][-32768:])
    shard = c.compress(data[40000:]) + c.flush()
    assert zz.ZzFlateDecode(shard, zz.Format.Deflate, max_len=len(data), dictionary=data[:40000]) == data[40000:]
    assert zz.ZzFlateDecode(shard, zz.Format.Deflate, max_len=len(data)) is None                  # distance beyond the start
    # concatenated gzip members (RFC 1952 2.2)
    import gzip
    two = gzip.compress(data[:50000]) + gzip.compress(data[50000:])
    assert zz.ZzFlateDecode(two, zz.Format.Gzip, max_len=len(data) + 16) == data
    assert zz.ZzFlateDecode(whole, zz.Format.Deflate, max_len=len(data)) == data


def test_rejects_damaged_streams(zz, golden):
    data = golden.input("alice29")[:50000]
    good = zlib.compress(data, 6)
    assert zz.ZzFlateDecode(good, zz.Format.Zlib, max_len=len(data)) == data
    assert zz.ZzFlateDecode(good[:-1], zz.Format.Zlib, max_len=len(data)) is None               # truncated trailer
    assert zz.ZzFlateDecode(good[: len(good) // 2], zz.Format.Zlib, max_len=len(data)) is None  # truncated body
    bad = bytearray(good); bad[-2] ^= 1
    assert zz.ZzFlateDecode(bytes(bad), zz.Format.Zlib, max_len=len(data)) is None              # Adler-32 mismatch
    bad = bytearray(good); bad[0] = 0x79
    assert zz.ZzFlateDecode(bytes(bad), zz.Format.Zlib, max_len=len(data)) is None              # header check
    assert zz.ZzFlateDecode(good, zz.Format.Zlib, max_len=len(data) - 1) is None                # output too small
    g = bytearray(__import__("gzip").compress(data)); g[-5] ^= 0x10
    assert zz.ZzFlateDecode(bytes(g), zz.Format.Gzip, max_len=len(data)) is None                # CRC-32 mismatch
