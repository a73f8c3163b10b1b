"""Sweep of the host-buffer piece schedule (options piece_first / piece_mid / piece_last / lanes): best wall time of 4 calls."""
import sys, ctypes as C, time, zlib, itertools
sys.path.insert(0, '.')
import torch
import zzflate_b200 as zz
from zzflate_b200 import _lib, synth
lib = _lib.load()
n = 1 << 30
src = torch.from_numpy(synth.markov_text(n)).pin_memory()
cap = zz.bound(n)
dst = torch.empty(cap, dtype=torch.uint8).pin_memory()
out_len = C.c_size_t(0); a0 = C.c_uint32(0); crc = C.c_uint32(0); st = _lib.Stats()
ref = None; res = []
for lanes, first, mid, last in itertools.product((1, 2, 3, 4), (888,), (666, 888, 1110, 1332), (888,)):
    for k, v in (("lanes", lanes), ("piece_first", first), ("piece_mid", mid), ("piece_last", last)): assert lib.zzgpu_set_option(k.encode(), v) == 0
    best = 1e9
    for it in range(4):
        t = time.perf_counter()
        rc = lib.zzgpu_deflate_ex(src.data_ptr(), n, 0, 1, 0, dst.data_ptr(), cap, 0, 2, 65536, 32768, 3, C.byref(out_len), C.byref(a0), C.byref(crc), C.byref(st))
        best = min(best, (time.perf_counter() - t) * 1e3)
        assert rc == 0, lib.zzgpu_last_error()
    h = (out_len.value, a0.value, crc.value)
    ref = ref or h
    res.append((best, lanes, first, mid, last, h == ref))
for r in sorted(res)[:12]: print(r)
print('worst', sorted(res)[-1], 'all same', all(r[-1] for r in res))
