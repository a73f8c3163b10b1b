"""Live comparison of the restatement with the compiled reference on seeded inputs (dev container only:
skipped where oracle/_ref/libzzref.so is absent)."""
import zlib

import numpy as np
import pytest

from oracle_lib import DEFLATE, GZIP, ZLIB, _padded
from zzflate_b200 import synth

S, D = 65536, 32768


def _inputs():
    rng = np.random.default_rng(11)
    yield "markov", synth.markov_text(400000, seg0=5).tobytes()
    yield "pattern", synth.repetitive(200000).tobytes()
    yield "zeros", bytes(200000)
    yield "lowent", rng.integers(0, 4, 150000, dtype=np.uint8).tobytes()
    yield "base64ish", (rng.integers(0, 64, 150000, dtype=np.uint8) + 48).tobytes()
    yield "runs", np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 600, 3000)).tobytes()[:300000]
    yield "period7", (bytes(range(7)) * 40000)[:250000]
    yield "textzeros", synth.markov_text(70000).tobytes() + bytes(70000) + synth.markov_text(70000, 9).tobytes()


@pytest.mark.parametrize("name,data", list(_inputs()), ids=[n for n, _ in _inputs()])
def test_chunks_identical_or_reference_defective(oracle, reference, name, data):
    buf = _padded(data)
    for level in (0, 1, 2):
        for off in range(0, len(data), S):
            ln = min(S, len(data) - off); d = min(D, off); final = off + ln == len(data)
            want = reference.chunk_encode(buf, off, ln, d, level, final, want_tokens=True)
            got = oracle.chunk_encode(buf, off, ln, d, level, final, want_tokens=True)
            if got["bytes"] != want["bytes"]:
                assert got["defects"] != 0, (name, level, off)
                z = zlib.decompressobj(-15, zdict=data[off - d: off]) if d else zlib.decompressobj(-15)
                assert z.decompress(got["bytes"]) == data[off: off + ln]
            elif level == 2 and got["defects"] == 0:
                assert np.array_equal(got["records"], want["records"]), (name, off)


@pytest.mark.parametrize("fmt", [ZLIB, GZIP, DEFLATE])
def test_whole_stream_identical(oracle, reference, fmt):
    data = synth.markov_text(1200000, seg0=2).tobytes()          # > 2 blocks of 500000: table carry-over
    for level in (0, 1, 2, 3):
        assert oracle.stream_reference(data, fmt, level)[0] == reference.encode(data, fmt, level)


def test_callback_api_equals_buffer_api(reference):
    data = synth.markov_text(300000).tobytes()
    for level in (0, 2):
        assert reference.encode(data, ZLIB, level, callback=True) == reference.encode(data, ZLIB, level)


def test_huffman_random_against_reference(oracle, reference):
    rng = np.random.default_rng(5)
    for t in range(300):
        n = (286, 30, 19)[t % 3]
        f = (rng.pareto(0.8, n) * 2).astype(np.int64).clip(0, 65000) * (rng.random(n) < 0.7)
        if t % 7 == 0:
            f = (2 ** rng.integers(0, 15, n)) * (rng.random(n) < 0.6)
        limit = 7 if n == 19 else 15
        assert oracle.calc_lengths(f.tolist(), limit) == reference.calc_lengths(f.tolist(), limit)


def test_adler_overflow_defect_R3_is_where_survey_says(oracle):
    # below the threshold the literal restatement of adler32x equals true Adler-32
    d = bytes([0xFF]) * (1 << 20)
    assert oracle.adler32x_literal(d, 1) == oracle.adler32(d, 1) == zlib.adler32(d)
