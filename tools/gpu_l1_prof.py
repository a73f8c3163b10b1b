"""One level-1 encode of a host buffer (piece-wise path: k_fixed launches of 888 / 1332 chunks) -- the target of the
ncu captures of K-FIXED.   python tools/gpu_l1_prof.py <workload> <MiB>"""
import sys
sys.path.insert(0, '.')
import zzflate_b200 as zz
from zzflate_b200 import synth
wl, mib = sys.argv[1], int(sys.argv[2])
data = synth.workload(wl, mib << 20)
out, a0, crc, st = zz.deflate_raw(data, level=1)
print(wl, mib, 'MiB ->', len(out), 'bytes; device', round(st.device_ms, 2), 'ms; stages', [round(x, 2) for x in st.stage_ms])
