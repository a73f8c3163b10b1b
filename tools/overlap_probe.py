import sys
sys.path.insert(0, '.')
import torch, zzflate_b200 as zz
from zzflate_b200 import synth, _lib
lib = _lib.load()
n = 1 << 30
src = torch.from_numpy(synth.markov_text(n)).cuda()
dst = torch.empty(zz.bound(n), dtype=torch.uint8, device='cuda')
ref = None
for mode in (0, 1, 0, 1):
    lib.zzgpu_set_option(b"overlap", mode)
    ms = []
    for _ in range(4):
        out_len, a0, crc, st = zz.deflate_device(src.data_ptr(), n, dst.data_ptr(), dst.numel(), checksums=1)
        ms.append(st.device_ms)
    h = hash(dst[:out_len].cpu().numpy().tobytes())
    ref = ref or h
    print('overlap', mode, 'ms', [round(x, 2) for x in ms], 'GB/s', round(n / min(ms) / 1e6, 2), 'same bytes', h == ref)
