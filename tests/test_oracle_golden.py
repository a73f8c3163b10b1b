"""The CPU restatement (oracle/zz_oracle.c) against golden vectors produced by the unmodified reference
(tests/golden/vectors.npz).  Runs without a GPU and without /root/reference."""
import zlib

import numpy as np
import pytest

from oracle_lib import DEFLATE, GZIP, ZLIB, _padded

S, D = 65536, 32768


def test_golden_has_cases(golden):
    assert len(golden.cases) >= 30


@pytest.mark.parametrize("level", [0, 1, 2])
def test_whole_stream_matches_reference(oracle, golden, level):
    """zzo_stream_reference == ZzFlateEncode(Zlib, level, threaded=false) wherever the reference inflates."""
    checked = 0
    for case in golden.cases:
        data = golden.input(case)
        want, ok = golden.stream(case, level)
        got, defects = oracle.stream_reference(data, ZLIB, level)
        if ok:
            assert got == want, (case, level)
            checked += 1
        else:
            assert defects != 0, (case, level)            # the restatement knows the reference is defective here
            assert zlib.decompress(got) == data           # ... and is itself correct
    assert checked >= 25


@pytest.mark.parametrize("level", [0, 1, 2])
def test_chunk_streams_match_reference(oracle, golden, level):
    """E(c) of SURVEY A.7: byte identity per chunk with the reference Encoder primed with the dictionary."""
    identical = defective = 0
    for case in golden.cases:
        data = golden.input(case)
        buf = _padded(data)
        want, well = golden.chunks(case, level)
        for i, off in enumerate(range(0, len(data), S)):
            ln = min(S, len(data) - off)
            r = oracle.chunk_encode(buf, off, ln, min(D, off), level, off + ln == len(data))
            if well[i]:
                assert r["bytes"] == want[i], (case, level, i)
                identical += 1
            else:
                assert r["defects"] != 0, (case, level, i)
                defective += 1
    assert identical >= 40
    if level == 0:
        assert defective == 0


@pytest.mark.parametrize("level", [0, 1, 2, 3])
@pytest.mark.parametrize("fmt,wbits", [(ZLIB, 15), (GZIP, 31), (DEFLATE, -15)])
def test_chunked_stream_inflates(oracle, golden, level, fmt, wbits):
    for case in golden.cases:
        data = golden.input(case)
        s, _ = oracle.stream_chunked(data, fmt, level)
        assert zlib.decompress(s, wbits) == data, (case, level, fmt)


def test_chunked_stream_is_concatenation_of_chunks(oracle, golden):
    for case in ("alice29", "mixed", "zeros"):
        data = golden.input(case); buf = _padded(data)
        parts = []
        for off in range(0, len(data), S):
            ln = min(S, len(data) - off)
            parts.append(oracle.chunk_encode(buf, off, ln, min(D, off), 2, off + ln == len(data))["bytes"])
        s, _ = oracle.stream_chunked(data, DEFLATE, 2)
        assert s == b"".join(parts)
        mt, _ = oracle.stream_chunked(data, DEFLATE, 2, threads=3)
        assert mt == s


def test_empty_input_policy(oracle):
    # R7: the reference emits no block at all; the restatement emits one final empty stored block
    s, _ = oracle.stream_chunked(b"", ZLIB, 2)
    assert s == bytes([0x78, 0x01, 0x01, 0x00, 0x00, 0xFF, 0xFF, 0, 0, 0, 1])
    assert zlib.decompress(s) == b""


def test_small_chunk_sizes(oracle, golden):
    data = golden.input("alice29")[:50000]
    for chunk, dict_size in [(4096, 2048), (1024, 32768), (8192, 0), (32768, 32768)]:
        for level in (1, 2):
            s, _ = oracle.stream_chunked(data, DEFLATE, level, chunk, dict_size)
            assert zlib.decompress(s, -15) == data


def test_huffman_lengths_known_answers(oracle, golden):
    freqs, lens, shape = golden["huff/freqs"], golden["huff/lengths"], golden["huff/shape"]
    iters = 0
    for f, l, (n, limit) in zip(freqs, lens, shape):
        got, it = oracle.calc_lengths(f[:n].tolist(), int(limit), want_iters=True)
        assert got == l[:n].tolist()
        iters += it > 1
    assert iters >= 10           # the frequency-floor limiter (huffman.cpp:122-154) was exercised


def test_rle_known_answers(oracle, golden):
    for l, r, f in zip(golden["rle/lengths"], golden["rle/records"], golden["rle/freqs"]):
        n, cnt = int(f[20]), int(f[19])
        recs, f19 = oracle.from_lengths(l[:n].tolist())
        assert len(recs) == cnt
        assert recs == [tuple(x) for x in r[:cnt].tolist()]
        assert f19 == f[:19].tolist()


def test_checksum_known_answers(oracle, golden):
    kat = bytes([0, 1, 23, 30, 4, 69, 145, 32, 216])        # zztest/Test.cpp:301-313
    a_all, a_first, a_last0, a_comb, crc = [int(x) for x in golden["cksum/kat"]]
    assert oracle.adler32(kat, 1) == a_all == a_comb
    assert oracle.combine(oracle.adler32(kat[:5], 1), oracle.adler32(kat[5:], 0), 4) == a_all
    assert oracle.adler32(kat[:5], 1) == a_first and oracle.adler32(kat[5:], 0) == a_last0
    assert oracle.crc32(kat) == crc == zlib.crc32(kat)
    for name, row in zip(golden["cksum/names"], golden["cksum/cases"]):
        d = golden.input(str(name))
        start_a = 0x12345678 % 65521 | (77 << 16)
        assert [oracle.adler32(d, 1), oracle.crc32(d), oracle.adler32(d, start_a), oracle.crc32(d, 0xDEADBEEF)] == [int(x) for x in row]
        assert oracle.adler32x_literal(d, 1) == int(row[0])
        assert oracle.adler32(d, 1) == zlib.adler32(d)


def test_crc_combine_is_new_but_consistent(oracle):
    rng = np.random.default_rng(7)
    for n1, n2 in [(0, 5), (5, 0), (1, 1), (1000, 77), (65536, 65536), (123457, 3)]:
        a = rng.integers(0, 256, n1, dtype=np.uint8).tobytes(); b = rng.integers(0, 256, n2, dtype=np.uint8).tobytes()
        assert oracle.crc32_combine(zlib.crc32(a), zlib.crc32(b), n2) == zlib.crc32(a + b)
        assert oracle.crc32(b, oracle.crc32(a)) == zlib.crc32(a + b)       # chaining through startValue (crc.cpp:24-26)


def test_static_tables_match_reference(oracle, golden):
    t = oracle.tables()
    for k, v in t.items():
        assert [int(x) for x in golden[f"tables/{k}"]] == [int(x) & 0xFFFFFFFF if k.endswith("_f") else int(x) for x in v] or \
               [int(x) for x in golden[f"tables/{k}"]] == [int(x) for x in v], k
