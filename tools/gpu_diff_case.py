import sys, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import zzflate_b200 as zz
from oracle_lib import oracle, _padded, DEFLATE
o = oracle()
fn, level, chunk, dict_size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
data = open(fn, 'rb').read()
got, *_ = zz.deflate_raw(data, level=level, chunk=chunk, dict_size=dict_size)
want, defects = o.stream_chunked(data, DEFLATE, level, chunk, dict_size)
print('equal', got == want, len(got), len(want), 'defects', defects)
buf = _padded(data)
for ci, off in enumerate(range(0, len(data), chunk)):
    ln = min(chunk, len(data) - off)
    tap = zz.debug_chunk(data, ci, level=level, chunk=chunk, dict_size=dict_size)
    w = o.chunk_encode(buf, off, ln, min(dict_size, off), level, off + ln == len(data), want_tokens=True)
    cand = o.chunk_candidates(buf, off, ln, min(dict_size, off))
    m1, m2 = tap['matches'], w['matches']
    ceq = np.array_equal(tap['cand'][:ln], cand)
    if not np.array_equal(m1, m2) or not ceq or not np.array_equal(tap['hist'][:286], w['lit_freq']):
        k = 0
        while k < min(len(m1), len(m2)) and (m1[k] == m2[k]).all(): k += 1
        print('chunk', ci, 'off', off, 'cand eq', ceq, 'ntok', len(m1), len(m2), 'first diff', k, m1[max(0,k-2):k+3].tolist(), m2[max(0,k-2):k+3].tolist(), 'defects', w['defects'])
        if not ceq:
            bad = np.nonzero(tap['cand'][:ln] != cand)[0]; print('  cand diffs', len(bad), bad[:8], cand[bad[:8]], tap['cand'][bad[:8]])
        break
