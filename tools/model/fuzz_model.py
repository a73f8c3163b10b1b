"""CPU fuzz of tools/model/lz_model.c against the oracle's tokens (development aid; no GPU).
   python tools/model/fuzz_model.py [seconds] [seed]"""
import ctypes as C, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from zzflate_b200 import synth
from oracle_lib import oracle, _padded, PAD
lib = C.CDLL('tools/model/liblzmodel.so')
lib.lzm_chunk.restype = C.c_int
lib.lzm_chunk.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
o = oracle()
_main = __name__ == "__main__"
budget = float(sys.argv[1]) if _main and len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if _main and len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
text = synth.markov_text(1 << 20, seg0=11, threads=1).tobytes()

def piece(kind, n):
    if kind == 0:
        s = int(rng.integers(0, len(text) - n)); return text[s: s + n]
    if kind == 1: return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == 2: return bytes([int(rng.integers(0, 256))]) * n
    if kind == 3:
        p = rng.integers(0, 256, int(rng.integers(1, 2000)), dtype=np.uint8).tobytes()
        return (p * (n // len(p) + 1))[:n]
    if kind == 4: return rng.integers(0, int(rng.integers(2, 6)), n, dtype=np.uint8).tobytes()
    if kind == 6: return (rng.integers(0, 64, n, dtype=np.uint8) + 48).astype(np.uint8).tobytes()
    p = rng.integers(97, 123, 7, dtype=np.uint8).tobytes()
    return (p * (n // 7 + 1))[:n]

def gen():
    total = int(rng.choice([300, 5000, 65536, 65537, 70000, 131072, 200000]))
    buf = bytearray()
    while len(buf) < total:
        kind = int(rng.integers(0, 8))
        n = int(rng.choice([1, 3, 17, 258, 259, 300, 1000, 5000, 16384, 20000, 40000]))
        if kind == 5:
            if len(buf) > 600:
                back = int(rng.integers(8, min(len(buf), 40000)))
                ln = int(rng.integers(4, min(back + 1, 3000) + 1))
                s = len(buf) - back
                buf += buf[s: s + ln]
        else:
            buf += piece(kind, n)
    return bytes(buf[:total])

def check(data, chunk=65536, dict_size=32768):
    buf = _padded(data); n = len(data)
    bad = 0
    for off in range(0, n, chunk):
        ln = min(chunk, n - off); final = off + ln == n
        d = min(dict_size, off)
        cand = o.chunk_candidates(buf, off, ln, d)
        want = o.chunk_encode(buf, off, ln, d, 2, final, want_tokens=True)["matches"]
        tok = np.zeros(3 * 20000, dtype=np.uint32)
        pre = min(off, d + 288)
        cand = np.ascontiguousarray(cand)
        k = lib.lzm_chunk(buf.ctypes.data + off, ln, ln if final else ln - 1, pre, cand.ctypes.data, tok.ctypes.data, 20000)
        got = tok[: 3 * k].reshape(-1, 3)
        if got.shape != want.shape or not np.array_equal(got, want):
            bad += 1
            m = min(len(got), len(want))
            diff = np.nonzero((got[:m] != want[:m]).any(axis=1))[0]
            print("MISMATCH chunk at", off, "tokens", len(got), len(want), "first diff", (diff[0], got[diff[0]], want[diff[0]]) if len(diff) else None, flush=True)
    return bad

if __name__ == "__main__":
    cases = fails = 0
    t0 = time.time()
    for spec in (1, 0):
        lib.lzm_set_spec(spec)
        for name in ("text", "zeros", "pattern"):
            data = synth.workload(name, 3 * 65536 + 1234).tobytes() if name != "text" else text[: 3 * 65536 + 1234]
            fails += check(data); cases += 1
    lib.lzm_set_spec(1)
    while time.time() - t0 < budget:
        data = gen()
        geom = (65536, 32768) if rng.random() < 0.8 else (int(rng.choice([4096, 8192, 32768])), int(rng.choice([0, 2048, 32768])))
        b = check(data, *geom); cases += 1; fails += b
        if b:
            open(f"/tmp/model_fail_{seed}_{cases}.bin", "wb").write(data)
    st = (C.c_long * 3)(); lib.lzm_stats(st)
    print(f"model fuzz: {cases} cases, {fails} failing chunks; true-walk steps {st[0]}, merges {st[1]}, spec steps {st[2]}")
    sys.exit(1 if fails else 0)
