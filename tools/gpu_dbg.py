import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import zzflate_b200 as zz
from oracle_lib import oracle, _padded
o = oracle()
data = b'hello hello hello hello, this is a test of the emergency broadcast system' * 30
d = zz.debug_chunk(data, 0)
buf = _padded(data)
r = o.chunk_encode(buf, 0, len(data), 0, 2, True, want_tokens=True)
print('gpu meta', d['meta_len'].tolist())
print('ora meta', r['meta_len'].tolist())
print('gpu lit ', d['lit_len'].tolist()[90:130], d['lit_len'].tolist()[250:])
print('ora lit ', r['lit_len'].tolist()[90:130], r['lit_len'].tolist()[250:])
print('gpu dist', d['dist_len'].tolist()); print('ora dist', r['dist_len'].tolist())
recs, f19 = o.from_lengths(r['lit_len'].tolist())
recs2, f19 = o.from_lengths(r['dist_len'].tolist(), f19)
print('ora metaF', f19, 'lens', o.calc_lengths(f19, 7))
print(d['hdr_bits'], d['total_bits'], r['block_bits'])
