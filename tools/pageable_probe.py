import sys, time, zlib
sys.path.insert(0, '.')
import numpy as np
import zzflate_b200 as zz
from zzflate_b200 import synth
n = 1 << 30
data = synth.markov_text(n)              # ordinary (pageable) numpy memory
cfg = zz.Config(zz.Format.Deflate, 2, False)
dest = np.empty(zz.bound(n) + 32, dtype=np.uint8)
dest[:] = 0
for it in range(4):
    t = time.perf_counter()
    w = zz.encode_ptr(dest.ctypes.data, dest.size, data.ctypes.data, n, cfg)
    dt = time.perf_counter() - t
    print('pageable host buffers: %.1f ms  %.2f GB/s  out %d' % (dt * 1e3, n / dt / 1e9, w))
