"""Generates tests/golden/vectors.npz from the UNMODIFIED reference (oracle/_ref/libzzref.so).

Run in the dev container only (needs /root/reference).  The GPU box has neither the reference sources
nor its corpus, so the inputs (slices of the reference's own test corpus plus seeded synthetic buffers)
travel inside the fixture together with the reference's outputs:

  <case>/input                     the input bytes
  <case>/L<level>/stream           ZzFlateEncode(Zlib, level, threaded=false) of the whole input
  <case>/L<level>/chunks           concatenated per-chunk E(c) streams (64 KiB chunks, 32 KiB dictionary,
                                   SURVEY A.7 recipe on the reference Encoder), sizes in .../sizes
  <case>/L<level>/wellformed       per chunk: the reference's own chunk stream inflates to the chunk
                                   (false only where the reference is defective)
  huff/*                           CalcLengths / FromLengths known answers
  cksum/*                          adler32x / crc32 / combine known answers
"""
import sys, zlib
from pathlib import Path
import numpy as np

root = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(root)); sys.path.insert(0, str(root / "tests"))
from oracle_lib import reference, _padded, ZLIB           # noqa: E402
from zzflate_b200 import synth                            # noqa: E402

ref = reference()
corpus = Path("/root/reference/zztest/corpus")
S, D = 65536, 32768
rng = np.random.default_rng(20261018)

cases = {}
for name, lo, n in [("alice29", 0, 140000), ("kennedy", 300000, 70000), ("ptt5", 0, 70000), ("lcet10", 100000, 66000),
                    ("plrabn12", 65000, 70000), ("asyoulik", 0, 65536), ("grammar", 0, None), ("xargs", 0, None), ("fields", 0, None),
                    ("cp", 0, None), ("sum", 0, None)]:
    fn = {"alice29": "alice29.txt", "kennedy": "kennedy.xls", "lcet10": "lcet10.txt", "plrabn12": "plrabn12.txt",
          "asyoulik": "asyoulik.txt", "grammar": "grammar.lsp", "xargs": "xargs.1", "fields": "fields.c", "cp": "cp.html"}.get(name, name)
    b = (corpus / fn).read_bytes()
    cases[name] = b[lo: lo + n] if n else b
exe = Path("/root/reference/zztest/ADInsight.exe").read_bytes()
cases["adinsight"] = exe[320000:320000 + 140000]
cases["markov"] = synth.markov_text(200000).tobytes()
cases["random"] = synth.random_bytes(70000).tobytes()
cases["zeros"] = bytes(140000)
cases["pattern"] = synth.repetitive(140000).tobytes()
cases["mixed"] = (synth.markov_text(40000).tobytes() + synth.random_bytes(40000).tobytes() + bytes(30000)
                  + synth.markov_text(30000, seg0=3).tobytes())
cases["hello"] = b"hello hello hello hello"
cases["abc"] = b"abcabcabcabcabcabcabcabcabcabc" * 20
cases["one"] = b"a"
cases["text300"] = synth.markov_text(300).tobytes()       # defect R1 territory for the whole-stream reference
for k in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):         # Test.cpp:330-338 SmallZerBouffer
    cases[f"zero{k}"] = bytes(k)

out = {}
summary = []
for name, data in cases.items():
    buf = _padded(data); n = len(data)
    out[f"{name}/input"] = np.frombuffer(data, dtype=np.uint8)
    for level in (0, 1, 2):
        stream = ref.encode(data, ZLIB, level)
        try:
            stream_ok = zlib.decompress(stream) == data
        except zlib.error:
            stream_ok = False
        out[f"{name}/L{level}/stream"] = np.frombuffer(stream, dtype=np.uint8)
        out[f"{name}/L{level}/stream_ok"] = np.array([stream_ok])
        chunks, sizes, well = [], [], []
        for off in range(0, n, S):
            ln = min(S, n - off); d = min(D, off); final = off + ln == n
            c = ref.chunk_encode(buf, off, ln, d, level, final)["bytes"]
            try:
                z = zlib.decompressobj(-15, zdict=data[off - d: off]) if d else zlib.decompressobj(-15)
                ok = z.decompress(c) == data[off: off + ln]
            except zlib.error:
                ok = False
            chunks.append(c); sizes.append(len(c)); well.append(ok)
        out[f"{name}/L{level}/chunks"] = np.frombuffer(b"".join(chunks), dtype=np.uint8)
        out[f"{name}/L{level}/sizes"] = np.array(sizes, dtype=np.int64)
        out[f"{name}/L{level}/wellformed"] = np.array(well)
        summary.append((name, level, n, len(stream), stream_ok, sum(sizes), all(well)))

# Huffman known answers
freqs, lens, limits = [], [], []
for t in range(240):
    n = (286, 30, 19)[t % 3]
    style = t % 4
    f = np.zeros(286, dtype=np.int64)
    if style == 0:
        f[:n] = rng.integers(0, 300, n) * (rng.random(n) < 0.6)
    elif style == 1:
        f[:n] = np.floor(rng.pareto(0.7, n) * 3).clip(0, 60000) * (rng.random(n) < 0.8)
    elif style == 2:
        f[:n] = rng.integers(0, 4, n)
    else:
        f[:n] = (2 ** rng.integers(0, 16, n)) * (rng.random(n) < 0.5)     # forces the length limiter
    limit = 7 if n == 19 else 15
    l = np.zeros(286, dtype=np.int64); l[:n] = ref.calc_lengths(f[:n].tolist(), limit)
    freqs.append(f); lens.append(l); limits.append((n, limit))
out["huff/freqs"] = np.array(freqs); out["huff/lengths"] = np.array(lens); out["huff/shape"] = np.array(limits)
rle_in, rle_out, rle_freq = [], [], []
for t in range(60):
    n = (286, 30)[t % 2]
    l = np.zeros(286, dtype=np.int64)
    l[:n] = rng.integers(0, 16, n) * (rng.random(n) < (0.2, 0.5, 0.9)[t % 3])
    if t % 5 == 0:
        l[:n] = np.repeat(rng.integers(0, 16, 8), 40)[:n]
    recs, f19 = ref.from_lengths(l[:n].tolist())
    r = np.zeros((300, 2), dtype=np.int64); r[: len(recs)] = recs
    rle_in.append(l); rle_out.append(r); rle_freq.append(f19 + [len(recs), n])
out["rle/lengths"] = np.array(rle_in); out["rle/records"] = np.array(rle_out); out["rle/freqs"] = np.array(rle_freq)

# checksum known answers (Test.cpp:301-313 plus random buffers)
kat = bytes([0, 1, 23, 30, 4, 69, 145, 32, 216])
out["cksum/kat"] = np.array([ref.adler32x(kat, 1), ref.adler32x(kat[:5], 1), ref.adler32x(kat[5:], 0),
                             ref.combine(ref.adler32x(kat[:5], 1), ref.adler32x(kat[5:], 0), 4), ref.crc32(kat)], dtype=np.uint64)
ck = []
for name in ("alice29", "kennedy", "random", "zeros", "hello", "one"):
    d = cases[name]
    ck.append([ref.adler32x(d, 1), ref.crc32(d), ref.adler32x(d, 0x12345678 % 65521 | (77 << 16)), ref.crc32(d, 0xDEADBEEF)])
out["cksum/cases"] = np.array(ck, dtype=np.uint64)
out["cksum/names"] = np.array(["alice29", "kennedy", "random", "zeros", "hello", "one"])
t = ref.tables()
for k, v in t.items():
    out[f"tables/{k}"] = np.array(v, dtype=np.int64)

dest = root / "tests" / "golden" / "vectors.npz"
np.savez_compressed(dest, **out)
for s in summary:
    print("%-10s L%d n=%-7d stream=%-7d ok=%-5s chunks=%-7d wellformed=%s" % s)
print("wrote", dest, dest.stat().st_size, "bytes")
