"""CPU fuzz of tools/model/l1_model.c: the step-wise level-1 walk against the sequential one (development aid; no GPU).
   python tools/model/fuzz_l1_model.py [seconds] [seed]"""
import ctypes as C, subprocess, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools/model')
import numpy as np
from zzflate_b200 import synth
subprocess.check_call("gcc -O2 -shared -fPIC -o tools/model/libl1model.so tools/model/l1_model.c".split())
lib = C.CDLL('tools/model/libl1model.so')
for f in (lib.l1m_seq, lib.l1m_warp, lib.l1m_warp2):
    f.restype = C.c_int; f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sys.argv = sys.argv[:1]
import fuzz_model as fm
fm.rng = np.random.default_rng(seed)

def check(data, chunk=65536, dict_size=32768):
    n = len(data)
    buf = np.zeros(n + 64, dtype=np.uint8); buf[:n] = np.frombuffer(data, dtype=np.uint8)
    bad = 0
    for off in range(0, n, chunk):
        ln = min(chunk, n - off); final = off + ln == n
        body = ln if final else ln - 1
        d = min(dict_size, off)
        a = np.zeros(3 * 70000, dtype=np.uint32); b = np.zeros(3 * 70000, dtype=np.uint32)
        ka = lib.l1m_seq(buf.ctypes.data + off, body, d, a.ctypes.data, 70000)
        kb = lib.l1m_warp(buf.ctypes.data + off, body, d, b.ctypes.data, 70000)
        if ka != kb or not np.array_equal(a[:3 * ka], b[:3 * kb]):
            bad += 1
            print("MISMATCH chunk at", off, ka, kb, flush=True)
        kb = lib.l1m_warp2(buf.ctypes.data + off, body, d, b.ctypes.data, 70000)
        if ka != kb or not np.array_equal(a[:3 * ka], b[:3 * kb]):
            bad += 1
            m = min(ka, kb); x = a[:3 * m].reshape(-1, 3); y = b[:3 * m].reshape(-1, 3)
            df = np.nonzero((x != y).any(axis=1))[0]
            print("MISMATCH (warp2) chunk at", off, ka, kb, (df[0], x[df[0]], y[df[0]]) if len(df) else None, flush=True)
    return bad

cases = fails = 0
st = (C.c_long * 3)()
for name in ("text", "random", "zeros", "pattern"):
    data = synth.workload(name, 16 * 65536 + 1234).tobytes()
    lib.l1m_stats(st)
    fails += check(data); cases += 1
    lib.l1m_stats(st)
    print(f"{name}: {len(data)} bytes, steps {st[0]} ({len(data) / st[0]:.1f} positions per step), ambiguous ends {st[1]}, long matches {st[2]}")
    s2 = (C.c_long * 3)(); lib.l1m_stats2(s2)
    print(f"   settled in parallel: {s2[0]} steps, {s2[1]} doubling rounds; sequential iterations {s2[2]}")
t0 = time.time()
while time.time() - t0 < budget:
    data = fm.gen()
    geom = (65536, 32768) if fm.rng.random() < 0.8 else (int(fm.rng.choice([4096, 8192, 32768])), int(fm.rng.choice([0, 2048, 32768])))
    b = check(data, *geom); cases += 1; fails += b
    if b: open(f"/tmp/l1_model_fail_{seed}_{cases}.bin", "wb").write(data)
print(f"l1 model fuzz: {cases} cases, {fails} failing chunks")
sys.exit(1 if fails else 0)
