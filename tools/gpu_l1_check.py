"""Bring-up check of the level-1 kernel (K-FIXED) on the GPU box: golden cases, synthetic workloads, odd geometries and a
short structured fuzz against the oracle at level 1; then the 1 GiB timings.  Option combinations (zzgpu_set_option) can be
given as further arguments; "-" runs the defaults.
   python tools/gpu_l1_check.py <fuzz seconds per combo> <time: 0/1> [name=v,name=v | -] ..."""
import sys, time, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools/model')
import numpy as np
import zzflate_b200 as zz
from zzflate_b200 import synth, _lib
from oracle_lib import oracle, DEFLATE
from conftest import Golden
o = oracle(); g = Golden(); lib = _lib.load()
S, D = 65536, 32768

def one(name, data, chunk=S, dict_size=D):
    try:
        got, *_ = zz.deflate_raw(data, level=1, chunk=chunk, dict_size=dict_size)
    except Exception as e:
        print('  ERROR', name, e, flush=True); return 1
    want, _ = o.stream_chunked(data, DEFLATE, 1, chunk, dict_size, threads=8 if len(data) > (4 << 20) else 1)
    if got == want: return 0
    k = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), -1)
    print('  MISMATCH', name, 'len', len(data), 'out', len(got), len(want), 'chunk', chunk, dict_size, 'first differing byte', k, flush=True)
    return 1

import fuzz_model
from fuzz_model import gen     # the same structured generator the CPU models were fuzzed with
import os
fuzz_model.rng = rng = np.random.default_rng(int(os.environ.get('L1_SEED', '1')))

budget = float(sys.argv[1]); timing = int(sys.argv[2])
total_bad = 0
for combo in (sys.argv[3:] or ['-']):
    for kv in (combo.split(',') if combo != '-' else []):
        k, v = kv.split('='); assert lib.zzgpu_set_option(k.encode(), int(v)) == 0, kv
    bad = 0; t0 = time.time()
    for case in g.cases:
        bad += one(case, g.input(case))
    for name in ('text', 'zeros', 'pattern', 'random'):
        bad += one(name, synth.workload(name, 5 * S + 4321).tobytes())
    big = synth.markov_text(300 * S + 99, seg0=4); big[7 * S: 8 * S] = 0; big[20 * S: 20 * S + 40000] = synth.random_bytes(40000)
    bad += one('text300', big.tobytes())
    for geom in ((4096, 2048), (8192, 0), (32768, 32768), (65536, 1000), (1024, 32768)):
        bad += one('alice-geom', g.input('alice29')[:90000], *geom)
        bad += one('pattern-geom', g.input('pattern')[:90000], *geom)
    cases = 0
    while time.time() - t0 < budget and bad < 6:
        data = gen()
        geom = (S, D) if rng.random() < 0.8 else (int(rng.choice([4096, 8192, 32768])), int(rng.choice([0, 2048, 32768])))
        b = one('fuzz', data, *geom); cases += 1
        if b:
            bad += b
            open(f'gpurun_out/l1fail_{combo.replace(",", "_").replace("=", "")}_{cases}_c{geom[0]}_d{geom[1]}.bin', 'wb').write(data)
    print(f'combo {combo}: {bad} mismatches, {cases} fuzz cases, {time.time() - t0:.0f} s', flush=True)
    total_bad += bad
    if timing:
        import torch
        n = 1 << 30
        for wl in ('text', 'random'):
            src = torch.from_numpy(synth.workload(wl, n)).cuda()
            dst = torch.empty(zz.bound(n, 1), dtype=torch.uint8, device='cuda')
            best = None
            for rep in range(3):
                out_len, a0, crc, st = zz.deflate_device(src.data_ptr(), n, dst.data_ptr(), dst.numel(), level=1, checksums=0)
                if best is None or st.device_ms < best.device_ms: best = st
            print(f'  timing {combo} {wl}: {best.device_ms:.2f} ms per GiB = {n / best.device_ms / 1e6:.1f} GB/s; stages', [round(x, 2) for x in best.stage_ms], 'out', out_len, flush=True)
            del src, dst
sys.exit(1 if total_bad else 0)
