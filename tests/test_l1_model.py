"""The level-1 step logic of K-FIXED, as tools/model/l1_model.c states it in plain C, pinned on the CPU:

* the step-wise walks (l1m_warp: one match per turn; l1m_warp2: the part of a step that does not depend on the visited set
  settled in parallel, as the kernel does) give the tokens of the sequential walk (l1m_seq, a restatement of
  encoder.cpp:329-373 / oracle write_block_fixed_huff);
* those tokens, written out with the fixed Huffman code of RFC 1951 3.2.6, are the oracle's bytes for the chunk -- so the
  model is tied to the oracle (and through it to the reference), not only to itself.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle_lib import DEFLATE, PAD, _padded

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tools" / "model" / "l1_model.c"
LIB = ROOT / "tools" / "model" / "libl1model.so"
S, D = 65536, 32768

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
             8193, 12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


@pytest.fixture(scope="module")
def model():
    if not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(LIB), str(SRC)])
    lib = C.CDLL(str(LIB))
    for f in (lib.l1m_seq, lib.l1m_warp, lib.l1m_warp2):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return lib


class Bits:
    def __init__(self):
        self.acc = 0; self.n = 0; self.out = bytearray()

    def put(self, value, nbits):                     # LSB first
        self.acc |= value << self.n; self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 0xFF); self.acc >>= 8; self.n -= 8

    def code(self, code, nbits):                     # Huffman codes go in MSB first
        self.put(int(format(code, f"0{nbits}b")[::-1], 2), nbits)

    def symbol(self, v):                             # RFC 1951 3.2.6
        if v < 144: self.code(0x30 + v, 8)
        elif v < 256: self.code(0x190 + v - 144, 9)
        elif v < 280: self.code(v - 256, 7)
        else: self.code(0xC0 + v - 280, 8)

    def pad(self):
        if self.n: self.put(0, 8 - self.n)


def fixed_block(chunk: bytes, body: int, tokens, final: bool) -> bytes:
    """E(c) at level 1 from the tokens of the main block."""
    b = Bits()
    if body > 0:
        b.put((1 if final else 0) | (1 << 1), 3)
        pos = 0
        for start, length, dist in tokens:
            for i in range(pos, start): b.symbol(chunk[i])
            ls = max(k for k in range(29) if LEN_BASE[k] <= length)
            if length == 258: ls = 28
            b.symbol(257 + ls); b.put(length - LEN_BASE[ls], LEN_EXTRA[ls])
            ds = max(k for k in range(30) if DIST_BASE[k] <= dist)
            b.code(ds, 5); b.put(dist - DIST_BASE[ds], DIST_EXTRA[ds])
            pos = start + length
        for i in range(pos, body): b.symbol(chunk[i])
        b.symbol(256)
    if not final:                                    # the reference's byte-aligning 1-byte stored block (zzflate.cpp:116-120)
        b.put(0, 3); b.pad(); b.put(1, 16); b.put(0xFFFE, 16); b.put(chunk[len(chunk) - 1], 8)
    b.pad()
    return bytes(b.out)


def walk(lib, fn, buf, off, body, dict_size):
    tok = np.zeros(3 * 70000, dtype=np.uint32)
    k = fn(buf.ctypes.data + off, body, dict_size, tok.ctypes.data, 70000)
    return tok[: 3 * k].reshape(-1, 3)


def inputs(golden):
    from zzflate_b200 import synth
    yield "alice29", golden.input("alice29")[:150000]
    yield "kennedy", golden.input("kennedy")[:100000]
    yield "pattern", golden.input("pattern")[:140000]
    yield "zeros", bytes(70000)
    yield "mixed", golden.input("mixed")
    yield "text", synth.markov_text(2 * S + 999, threads=1).tobytes()
    yield "random", synth.random_bytes(S + 77, threads=1).tobytes()
    rng = np.random.default_rng(7)
    yield "few-symbols", rng.integers(0, 3, 90000, dtype=np.uint8).tobytes()          # same-hash lanes in nearly every step
    yield "short-period", (bytes(rng.integers(97, 123, 7, dtype=np.uint8)) * 20000)[:100000]


def test_step_walks_equal_the_sequential_walk_and_the_oracle(model, oracle, golden):
    for name, data in inputs(golden):
        buf = _padded(data); n = len(data)
        for chunk, dict_size in ((S, D), (8192, 2048), (4096, 0)):
            if chunk != S and name not in ("alice29", "few-symbols", "short-period"):
                continue
            for off in range(0, n, chunk):
                ln = min(chunk, n - off); final = off + ln == n
                body = ln if final else ln - 1
                d = min(dict_size, off)
                seq = walk(model, model.l1m_seq, buf, off, body, d)
                for fn in (model.l1m_warp, model.l1m_warp2):
                    got = walk(model, fn, buf, off, body, d)
                    assert got.shape == seq.shape and np.array_equal(got, seq), (name, chunk, off)
                if chunk == S or off < 3 * chunk:                                   # the bit writer is plain Python: a few chunks per geometry
                    want = oracle.chunk_encode(buf, off, ln, d, 1, final)["bytes"]
                    assert fixed_block(data[off: off + ln], body, seq.tolist(), final) == want, (name, chunk, off)


def test_step_walks_on_structured_random_inputs(model):
    """A short seeded run of the structured generator the model was fuzzed with (text, random, runs, short periods, few
    symbols, copies of earlier stretches at long distances), default and odd geometries."""
    from zzflate_b200 import synth
    rng = np.random.default_rng(20261018)
    text = synth.markov_text(1 << 19, seg0=11, threads=1).tobytes()

    def piece(kind, n):
        if kind == 0:
            s = int(rng.integers(0, len(text) - n)); return text[s: s + n]
        if kind == 1: return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        if kind == 2: return bytes([int(rng.integers(0, 256))]) * n
        if kind == 3:
            p = rng.integers(0, 256, int(rng.integers(1, 2000)), dtype=np.uint8).tobytes()
            return (p * (n // len(p) + 1))[:n]
        if kind == 4: return rng.integers(0, int(rng.integers(2, 6)), n, dtype=np.uint8).tobytes()
        if kind == 6: return (rng.integers(0, 64, n, dtype=np.uint8) + 48).astype(np.uint8).tobytes()
        p = rng.integers(97, 123, 7, dtype=np.uint8).tobytes()
        return (p * (n // 7 + 1))[:n]

    chunks = 0
    for _ in range(40):
        total = int(rng.choice([300, 5000, 65536, 65537, 70000, 131072]))
        buf = bytearray()
        while len(buf) < total:
            kind = int(rng.integers(0, 8))
            n = int(rng.choice([1, 3, 17, 258, 259, 300, 1000, 5000, 16384, 20000]))
            if kind == 5:
                if len(buf) > 600:
                    back = int(rng.integers(8, min(len(buf), 40000)))
                    ln = int(rng.integers(4, min(back + 1, 3000) + 1))
                    s = len(buf) - back
                    buf += buf[s: s + ln]
            else:
                buf += piece(kind, n)
        data = bytes(buf[:total])
        chunk, dict_size = (S, D) if rng.random() < 0.7 else (int(rng.choice([4096, 8192, 32768])), int(rng.choice([0, 2048, 32768])))
        b = _padded(data)
        for off in range(0, len(data), chunk):
            ln = min(chunk, len(data) - off); final = off + ln == len(data)
            body = ln if final else ln - 1
            d = min(dict_size, off)
            seq = walk(model, model.l1m_seq, b, off, body, d)
            for fn in (model.l1m_warp, model.l1m_warp2):
                got = walk(model, fn, b, off, body, d)
                assert got.shape == seq.shape and np.array_equal(got, seq), (chunk, dict_size, off, len(data))
            chunks += 1
    assert chunks > 60
