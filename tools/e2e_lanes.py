"""Host-buffer call with one and with two kernel lanes: wall time, stage sums, and the bytes against each other."""
import sys, ctypes as C, time, zlib
sys.path.insert(0, '.')
import numpy as np, torch
import zzflate_b200 as zz
from zzflate_b200 import _lib, synth
lib = _lib.load()
n = 1 << 30
src = torch.from_numpy(synth.markov_text(n)).pin_memory()
cap = zz.bound(n)
dst = torch.empty(cap, dtype=torch.uint8).pin_memory()
out_len = C.c_size_t(0); a0 = C.c_uint32(0); crc = C.c_uint32(0); st = _lib.Stats()
ref = None
for lanes in (0, 1, 0, 1):
    lib.zzgpu_set_option(b"lanes", lanes)
    for level in (2, 1):
        best = 1e9
        for it in range(4):
            t = time.perf_counter()
            rc = lib.zzgpu_deflate_ex(src.data_ptr(), n, 0, 1, 0, dst.data_ptr(), cap, 0, level, 65536, 32768, 3, C.byref(out_len), C.byref(a0), C.byref(crc), C.byref(st))
            wall = (time.perf_counter() - t) * 1e3
            if rc: print('ERR', lib.zzgpu_last_error()); break
            best = min(best, wall)
        h = (out_len.value, zlib.crc32(dst[:out_len.value].numpy()), a0.value, crc.value)
        if ref is None or level not in ref: ref = dict(ref or {}); ref[level] = h
        print('lanes', lanes, 'level', level, 'best wall ms', round(best, 2), 'GB/s', round(n / best / 1e6, 2), 'kernel sum', round(st.device_ms, 2),
              {k: round(st.stage_ms[i], 2) for i, k in enumerate(_lib.STAGES) if st.stage_launches[i]}, 'same', h == ref[level], flush=True)
