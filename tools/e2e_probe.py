import sys, ctypes as C, time
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
for it in range(4):
    t = time.perf_counter()
    rc = lib.zzgpu_deflate_ex(src.data_ptr(), n, 0, 1, 0, dst.data_ptr(), cap, 0, 2, 65536, 32768, 0, C.byref(out_len), C.byref(a0), C.byref(crc), C.byref(st))
    wall = (time.perf_counter() - t) * 1e3
    if rc: print('ERR', lib.zzgpu_last_error())
    print('rc', rc, 'wall ms', round(wall, 2), 'total_ms', round(st.total_ms, 2), 'kernel sum', round(st.device_ms, 2),
          {k: (round(st.stage_ms[i], 2), st.stage_launches[i]) for i, k in enumerate(_lib.STAGES) if st.stage_launches[i]})
# raw copy speeds
d = torch.empty(n, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); print('H2D GB/s', n / (time.perf_counter() - t) / 1e9)
t = time.perf_counter(); dst[:n//2].copy_(d[:n//2], non_blocking=True); torch.cuda.synchronize(); print('D2H GB/s', n / 2 / (time.perf_counter() - t) / 1e9)
# both directions at once (two streams), as the pipelined call uses the link
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
m = int(n * 0.55)
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s1): d.copy_(src, non_blocking=True)
    with torch.cuda.stream(s2): dst[:m].copy_(d[:m], non_blocking=True)
    torch.cuda.synchronize(); el = time.perf_counter() - t
    print('H2D 1 GiB + D2H 0.55 GiB concurrently: ms', round(el * 1e3, 2), 'H2D-equivalent GB/s', round(n / el / 1e9, 2))
