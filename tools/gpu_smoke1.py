import sys, zlib, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import zzflate_b200 as zz
from oracle_lib import oracle, _padded
o = oracle()
rng = np.random.default_rng(1)
words = [bytes(rng.integers(97, 123, rng.integers(2, 9)).astype(np.uint8)) for _ in range(2000)]
def text(n, seed):
    r = np.random.default_rng(seed); out = bytearray()
    while len(out) < n:
        out += words[int(r.zipf(1.3)) % len(words)] + b' '
    return bytes(out[:n])
cases = {'text200k': text(200000, 3), 'zeros100k': bytes(100000), 'rand70k': rng.integers(0,256,70000,dtype=np.uint8).tobytes(),
         'small': b'hello hello hello hello', 'one': b'a', 'text64k': text(65536, 5), 'text64k+1': text(65537, 6)}
for name, data in cases.items():
    for level in (2, 1, 0):
        try:
            out, a0, crc, st = zz.deflate_raw(data, level=level)
        except Exception as e:
            print(name, level, 'ERROR', e); continue
        ref, defects = o.stream_chunked(data, 2, level)   # DEFLATE=2
        try:
            ok = zlib.decompress(out, -15) == data
        except Exception as e:
            ok = f'inflate fail {e}'
        print(name, 'L', level, 'len', len(out), 'ref', len(ref), 'identical', out == ref, 'roundtrip', ok,
              'adler', zz.combine(1, a0, len(data)) == zlib.adler32(data), 'crc', crc == zlib.crc32(data),
              'ms', round(st.device_ms, 3), 'matches', st.matches, 'stored', st.stored_chunks)
        if out != ref and level == 2:
            buf = _padded(data)
            for ci in range((len(data)+65535)//65536):
                off = ci*65536; ln = min(65536, len(data)-off)
                d = zz.debug_chunk(data, ci)
                r = o.chunk_encode(buf, off, ln, min(32768, off), 2, off+ln==len(data), want_tokens=True)
                c = o.chunk_candidates(buf, off, ln, min(32768, off))
                print('  chunk', ci, 'cand eq', np.array_equal(c, d['cand'][:ln]), 'ntok', len(d['matches']), len(r['matches']),
                      'tok eq', np.array_equal(d['matches'], r['matches']),
                      'hist eq', np.array_equal(d['hist'][:286], r['lit_freq']) and np.array_equal(d['hist'][286:], r['dist_freq']),
                      'litlen eq', np.array_equal(d['lit_len'], r['lit_len']), 'dist eq', np.array_equal(d['dist_len'], r['dist_len']),
                      'meta eq', np.array_equal(d['meta_len'][:19], r['meta_len']), 'bt', d['block_type'], r['block_type'], 'bits', d['total_bits'], r['block_bits'])
                if not np.array_equal(d['matches'], r['matches']):
                    m1, m2 = d['matches'], r['matches']
                    k = 0
                    while k < min(len(m1), len(m2)) and (m1[k] == m2[k]).all(): k += 1
                    print('   first diff at', k, m1[max(0,k-1):k+2].tolist(), m2[max(0,k-1):k+2].tolist())
                if not np.array_equal(c, d['cand'][:ln]):
                    bad = np.nonzero(c != d['cand'][:ln])[0]
                    print('   cand diffs', len(bad), bad[:5], c[bad[:5]], d['cand'][bad[:5]])
