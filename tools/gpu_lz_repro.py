"""Re-runs a saved fuzz failure (<name>_c<chunk>_d<dict>.bin) under one K-LZ option combination and compares the
stream and every chunk's tokens with the oracle.
   python tools/gpu_lz_repro.py file name=v,name=v"""
import re, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import zzflate_b200 as zz
from zzflate_b200 import _lib
from oracle_lib import oracle, _padded, DEFLATE
o = oracle(); lib = _lib.load()
fn = sys.argv[1]
for kv in sys.argv[2].split(','):
    k, v = kv.split('='); assert lib.zzgpu_set_option(k.encode(), int(v)) == 0
m = re.search(r'_c(\d+)_d(\d+)\.bin', fn); chunk, dict_size = int(m.group(1)), int(m.group(2))
data = open(fn, 'rb').read()
want, _ = o.stream_chunked(data, DEFLATE, 2, chunk, dict_size)
res = []
for rep in range(3):
    got, *_ = zz.deflate_raw(data, level=2, chunk=chunk, dict_size=dict_size)
    res.append(got == want)
print("stream", res, flush=True)
buf = _padded(data); tok = []
for ci, off in enumerate(range(0, len(data), chunk)):
    ln = min(chunk, len(data) - off)
    tap = zz.debug_chunk(data, ci, chunk=chunk, dict_size=dict_size)
    w = o.chunk_encode(buf, off, ln, min(dict_size, off), 2, off + ln == len(data), want_tokens=True)
    tok.append(bool(np.array_equal(tap['matches'], w['matches'])))
print(fn.split('/')[-1], sys.argv[2], 'stream', res, 'tokens', tok, flush=True)
