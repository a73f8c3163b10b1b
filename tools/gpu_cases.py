import sys, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import zzflate_b200 as zz
from oracle_lib import oracle, _padded, DEFLATE
from conftest import Golden
o = oracle(); g = Golden()
S, D = 65536, 32768
for case in g.cases:
    data = g.input(case)
    out, *_ = zz.deflate_raw(data, level=2)
    want, _ = o.stream_chunked(data, DEFLATE, 2)
    if out == want: continue
    print('MISMATCH', case, len(data), len(out), len(want))
    buf = _padded(data)
    for ci, off in enumerate(range(0, len(data), S)):
        ln = min(S, len(data) - off)
        tap = zz.debug_chunk(data, ci)
        w = o.chunk_encode(buf, off, ln, min(D, off), 2, off + ln == len(data), want_tokens=True)
        m1, m2 = tap['matches'], w['matches']
        if not np.array_equal(m1, m2):
            k = 0
            while k < min(len(m1), len(m2)) and (m1[k] == m2[k]).all(): k += 1
            print('  chunk', ci, 'ntok', len(m1), len(m2), 'first diff', k, m1[max(0,k-2):k+3].tolist(), m2[max(0,k-2):k+3].tolist())
