"""Bring-up check of kernel variants on the GPU box: for every option combination, golden cases + synthetic
workloads + a short structured fuzz against the oracle; the first mismatches are localised (chunk, first differing token).
   python tools/gpu_lz_check.py <fuzz seconds per combo> name=v,name=v  [name=v,... more combos]"""
import sys, time, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools/model')
import numpy as np
import zzflate_b200 as zz
from zzflate_b200 import synth, _lib
from oracle_lib import oracle, _padded, DEFLATE
from conftest import Golden
o = oracle(); g = Golden(); lib = _lib.load()
S, D = 65536, 32768

def localise(data, chunk=S, dict_size=D, maxshow=2):
    buf = _padded(data); shown = 0
    for ci, off in enumerate(range(0, len(data), chunk)):
        ln = min(chunk, len(data) - off)
        try:
            tap = zz.debug_chunk(data, ci, chunk=chunk, dict_size=dict_size)
        except Exception as e:
            print('   debug_chunk failed', e); return
        w = o.chunk_encode(buf, off, ln, min(dict_size, off), 2, off + ln == len(data), want_tokens=True)
        m1, m2 = tap['matches'], w['matches']
        if not np.array_equal(m1, m2):
            k = 0
            while k < min(len(m1), len(m2)) and (m1[k] == m2[k]).all(): k += 1
            print('   chunk', ci, 'ntok', len(m1), len(m2), 'first diff', k, 'gpu', m1[max(0, k - 2):k + 3].tolist(), 'want', m2[max(0, k - 2):k + 3].tolist(), flush=True)
            shown += 1
            if shown >= maxshow: return
        elif not np.array_equal(tap['hist'][:286], w['lit_freq']):
            print('   chunk', ci, 'tokens equal but literal histogram differs'); shown += 1
            if shown >= maxshow: return

def one(name, data, level=2, chunk=S, dict_size=D):
    print('  case', name, len(data), chunk, dict_size, flush=True) if VERBOSE else None
    try:
        got, *_ = zz.deflate_raw(data, level=level, chunk=chunk, dict_size=dict_size)
    except Exception as e:
        print('  ERROR', name, e, flush=True); return 1
    want, _ = o.stream_chunked(data, DEFLATE, level, chunk, dict_size, threads=8 if len(data) > (4 << 20) else 1)
    if got == want: return 0
    print('  MISMATCH', name, 'len', len(data), 'out', len(got), len(want), 'chunk', chunk, dict_size, flush=True)
    if level >= 2: localise(data, chunk, dict_size)
    return 1

from fuzz_model import gen, rng     # the same structured generator the CPU model was fuzzed with

import os
VERBOSE = bool(os.environ.get('LZ_CHECK_VERBOSE'))
budget = float(sys.argv[1])
total_bad = 0
for combo in sys.argv[2:]:
    for kv in combo.split(','):
        k, v = kv.split('='); assert lib.zzgpu_set_option(k.encode(), int(v)) == 0, kv
    bad = 0; t0 = time.time()
    for case in g.cases:
        bad += one(case, g.input(case))
    for name in ('text', 'zeros', 'pattern', 'random'):
        bad += one(name, synth.workload(name, 5 * S + 4321).tobytes())
    big = synth.markov_text(300 * S + 99, seg0=4); big[7 * S: 8 * S] = 0; big[20 * S: 20 * S + 40000] = synth.random_bytes(40000)
    bad += one('text300', big.tobytes())
    for geom in ((4096, 2048), (8192, 0), (32768, 32768), (65536, 1000), (1024, 32768)):
        bad += one('alice-geom', g.input('alice29')[:90000], 2, *geom)
        bad += one('pattern-geom', g.input('pattern')[:90000], 2, *geom)
    cases = 0
    while time.time() - t0 < budget and bad < 6:
        data = gen()
        geom = (S, D) if rng.random() < 0.8 else (int(rng.choice([4096, 8192, 32768])), int(rng.choice([0, 2048, 32768])))
        b = one('fuzz', data, 2, *geom); cases += 1
        if b:
            bad += b
            open(f'gpurun_out/lzfail_{combo.replace(",", "_").replace("=", "")}_{cases}_c{geom[0]}_d{geom[1]}.bin', 'wb').write(data)
    print(f'combo {combo}: {bad} mismatches, {cases} fuzz cases, {time.time() - t0:.0f} s', flush=True)
    total_bad += bad
sys.exit(1 if total_bad else 0)
