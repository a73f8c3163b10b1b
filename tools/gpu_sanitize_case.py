import sys, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import zzflate_b200 as zz
from zzflate_b200 import synth
data = synth.markov_text(3 * 65536 + 4321, threads=1).tobytes() + bytes(70000) + synth.random_bytes(20000, threads=1).tobytes()
for level in (2, 1, 0):
    for fmt, wb in ((zz.Format.Zlib, 15), (zz.Format.Gzip, 31)):
        out = zz.ZzFlateEncode(data, zz.Config(fmt, level, False))
        assert zlib.decompress(out, wb) == data
out, *_ = zz.deflate_raw(data[:100000], level=2, chunk=4096, dict_size=2048)
assert zlib.decompress(out, -15) == data[:100000]
print("sanitize case ok")
