"""Randomised parity fuzz: GPU stream vs oracle stream on structured random inputs (run on the GPU box).
   python tools/gpu_fuzz.py [seconds] [seed]"""
import sys, time, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import zzflate_b200 as zz
from zzflate_b200 import synth
from oracle_lib import oracle, DEFLATE
o = oracle()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
text = synth.markov_text(1 << 20, seg0=11).tobytes()

def piece(kind, n):
    if kind == 0: return text[(o_ := int(rng.integers(0, len(text) - n))): o_ + n]
    if kind == 1: return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == 2: return bytes([int(rng.integers(0, 256))]) * n
    if kind == 3:
        p = rng.integers(0, 256, int(rng.integers(1, 2000)), dtype=np.uint8).tobytes()
        return (p * (n // len(p) + 1))[:n]
    if kind == 4: return rng.integers(0, int(rng.integers(2, 6)), n, dtype=np.uint8).tobytes()
    if kind == 5:      # long matches at long distances: copy of an earlier stretch
        return b''
    if kind == 6: return (rng.integers(0, 64, n, dtype=np.uint8) + 48).astype(np.uint8).tobytes()
    p = rng.integers(97, 123, 7, dtype=np.uint8).tobytes()
    return (p * (n // 7 + 1))[:n]

cases = fails = 0
t0 = time.time()
while time.time() - t0 < budget:
    total = int(rng.choice([300, 5000, 65536, 65537, 70000, 131072, 200000, 330000]))
    buf = bytearray()
    while len(buf) < total:
        kind = int(rng.integers(0, 8))
        n = int(rng.choice([1, 3, 17, 258, 259, 300, 1000, 5000, 16384, 20000, 40000]))
        if kind == 5 and len(buf) > 600:
            back = int(rng.integers(8, min(len(buf), 40000)))
            ln = int(rng.integers(4, min(back + 1, 3000) + 1))
            s = len(buf) - back
            buf += buf[s: s + ln]
        else:
            buf += piece(kind, n)
    data = bytes(buf[:total])
    for level in (2, 1):
        chunk, dict_size = (65536, 32768) if rng.random() < 0.8 else (int(rng.choice([4096, 8192, 32768])), int(rng.choice([0, 2048, 32768])))
        got, a0, crc, st = zz.deflate_raw(data, level=level, chunk=chunk, dict_size=dict_size)
        want, defects = o.stream_chunked(data, DEFLATE, level, chunk, dict_size)
        cases += 1
        if got != want or zlib.decompress(got, -15) != data:
            fails += 1
            fn = f'gpurun_out/fuzz_fail_{seed}_{cases}.bin'
            open(fn, 'wb').write(data)
            print('MISMATCH level', level, 'chunk', chunk, dict_size, 'len', len(data), 'saved', fn, flush=True)
print(f'fuzz: {cases} cases, {fails} failures, seed {seed}')
sys.exit(1 if fails else 0)
