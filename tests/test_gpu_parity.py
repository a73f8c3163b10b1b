"""Parity tests proper: the CUDA path, called through the C-ABI (libzzflate_b200.so), against the oracle on
the same inputs, against the reference's golden vectors, and -- at BASELINE sizes -- through
size-independent properties (zlib inflate round trip, checksum agreement).  Bit-exact everywhere."""
import zlib

import numpy as np
import pytest

from oracle_lib import DEFLATE, GZIP, ZLIB, _padded

pytestmark = pytest.mark.gpu

S, D = 65536, 32768


@pytest.fixture(scope="module")
def zz():
    import zzflate_b200
    from zzflate_b200 import build
    build.build()
    assert zzflate_b200.device_count() > 0, "no CUDA device: the GPU tests must not fall back to anything"
    return zzflate_b200


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_golden_cases_identical_to_oracle_and_reference(zz, oracle, golden, level):
    ref_level = min(level, 2)
    for case in golden.cases:
        data = golden.input(case)
        out, a0, crc, st = zz.deflate_raw(data, level=level)
        want, _ = oracle.stream_chunked(data, DEFLATE, level)
        assert out == want, (case, level)
        assert zlib.decompress(out, -15) == data
        assert zz.ZzFlateDecode(out, zz.Format.Deflate, max_len=len(data) + 16) == data          # include/decoder.h, the second judge
        assert zz.combine(1, a0, len(data)) == zlib.adler32(data) and crc == zlib.crc32(data)
        # directly against the reference's own per-chunk bytes, no oracle in between
        chunks, well = golden.chunks(case, ref_level)
        pos = 0
        for c, ok in zip(chunks, well):
            if ok:
                assert out[pos: pos + len(c)] == c, (case, level)
            pos += len(c) if ok else 0
            if not ok:
                break


def test_stage_taps_match_oracle(zz, oracle, golden):
    for case in ("alice29", "kennedy", "ptt5", "markov", "pattern", "zeros", "mixed", "adinsight", "abc"):
        data = golden.input(case); buf = _padded(data)
        for ci, off in enumerate(range(0, len(data), S)):
            ln = min(S, len(data) - off)
            tap = zz.debug_chunk(data, ci)
            want = oracle.chunk_encode(buf, off, ln, min(D, off), 2, off + ln == len(data), want_tokens=True)
            cand = oracle.chunk_candidates(buf, off, ln, min(D, off))
            assert np.array_equal(tap["cand"][:ln], cand), (case, ci)
            assert np.array_equal(tap["matches"], want["matches"]), (case, ci)
            assert np.array_equal(tap["hist"][:286], want["lit_freq"]) and np.array_equal(tap["hist"][286:], want["dist_freq"])
            assert np.array_equal(tap["lit_len"], want["lit_len"]) and np.array_equal(tap["dist_len"], want["dist_len"])
            assert np.array_equal(tap["meta_len"], want["meta_len"])
            assert tap["block_type"] == want["block_type"] and tap["total_bits"] == want["block_bits"]


@pytest.mark.parametrize("fmt,ofmt,wbits", [("Zlib", ZLIB, 15), ("Gzip", GZIP, 31), ("Deflate", DEFLATE, -15)])
def test_public_api_formats(zz, oracle, golden, fmt, ofmt, wbits):
    for case in ("alice29", "kennedy", "random", "zeros", "hello", "one", "zero512", "mixed"):
        data = golden.input(case)
        for level in (0, 1, 2, 3):
            for threaded in (False, True):
                cfg = zz.Config(zz.Format[fmt], level, threaded)
                out = zz.ZzFlateEncode(data, cfg)
                assert out is not None
                assert out == oracle.stream_chunked(data, ofmt, level)[0], (case, level)
                assert zlib.decompress(out, wbits) == data


def test_callback_api(zz, golden):
    # Test.cpp:207-222: header, data buffers, trailer arrive in order; concatenation is the stream
    data = golden.input("markov") * 12                      # > 1 000 000 bytes of output pieces
    for level in (0, 2):
        cfg = zz.Config(zz.Format.Zlib, level, False)
        pieces = []
        zz.ZzFlateEncodeToCallback(data, cfg, lambda b: pieces.append(b) or False)
        assert b"".join(pieces) == zz.ZzFlateEncode(data, cfg)
        assert pieces[0] == b"\x78\x01" and len(pieces[-1]) == 4
        assert all(len(p) <= 1000000 for p in pieces)
        if level == 0:
            assert len(pieces) >= 4


def test_callback_api_streams(zz, oracle):
    """N2 (zzflate.cpp:197-222 on the GPU): the callback receives the first data while later pieces of the input have
    not even been copied to the device, and nothing of the size of the stream is buffered: 600 MiB of pageable input
    are cut into pieces (H2D -> kernels -> D2H), the first slice must arrive before the last piece's H2D is done."""
    from zzflate_b200 import synth, _lib
    n = 600 << 20
    data = synth.markov_text(n, seg0=5)
    lib = _lib.load()
    state = {"first_at": None, "bytes": 0, "adler": 1, "crc": 0}
    d = zlib.decompressobj(15)
    pos = [0]
    ok = [True]

    def sink(b):
        if len(b) > 4 and state["first_at"] is None and pos[0] == 0 and state["bytes"] >= 2:
            state["first_at"] = (lib.zzgpu_get_counter(b"sink_first_h2d_done"), lib.zzgpu_get_counter(b"sink_pieces"))
        state["bytes"] += len(b)
        piece = d.decompress(b)
        ok[0] = ok[0] and piece == data[pos[0]: pos[0] + len(piece)].tobytes()
        pos[0] += len(piece)
        assert len(b) <= 1000000
        return False

    zz.ZzFlateEncodeToCallback(data, zz.Config(zz.Format.Zlib, 2, False), sink)
    assert ok[0] and pos[0] == n and d.eof
    done, pieces = lib.zzgpu_get_counter(b"sink_first_h2d_done"), lib.zzgpu_get_counter(b"sink_pieces")
    assert pieces >= 4 and 1 <= done < pieces, (done, pieces)


def test_gzip_members_for_large_inputs(zz, oracle, golden):
    """N4: a gzip member stores its input length modulo 2^32, so inputs of 4 GiB and more are written as a series of members
    (RFC 1952 2.2), each a complete stream with its own CRC-32 and ISIZE.  With the member size lowered to 128 KiB a
    300 KiB input becomes three members; gzip / zlib / the library's own inflater read them back as one stream."""
    import ctypes as C, gzip
    from zzflate_b200 import _lib
    lib = _lib.load()
    lib.zz_c_set_gzip_member_bytes.argtypes = [C.c_size_t]; lib.zz_c_set_gzip_member_bytes.restype = None
    data = (golden.input("alice29") + golden.input("kennedy") + golden.input("markov"))[:307200]
    lib.zz_c_set_gzip_member_bytes(2 * S)
    try:
        for threaded in (False, True):
            out = zz.ZzFlateEncode(data, zz.Config(zz.Format.Gzip, 2, threaded))
            assert out is not None and len(data) > 4 * S and out.count(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff") >= 3
            assert gzip.decompress(out) == data
            assert zz.ZzFlateDecode(out, zz.Format.Gzip, max_len=len(data) + 16) == data
            members = [oracle.stream_chunked(data[o: o + 2 * S], GZIP, 2)[0] for o in range(0, len(data), 2 * S)]
            assert out == b"".join(members)
        pieces = []
        zz.ZzFlateEncodeToCallback(data, zz.Config(zz.Format.Gzip, 2, False), lambda b: pieces.append(b) or False)
        assert b"".join(pieces) == out
        z = zz.ZzFlateEncode(data, zz.Config(zz.Format.Zlib, 2, False))          # zlib has no length field: one stream
        assert zlib.decompress(z) == data and z == oracle.stream_chunked(data, ZLIB, 2)[0]
    finally:
        lib.zz_c_set_gzip_member_bytes(0)


def test_free_mode(zz, oracle, golden):
    """N1 (first step): mode 1 of zzgpu_deflate_mode keeps the tokens and code lengths of the reference-equivalent parse but
    trims HLIT / HDIST / HCLEN in the dynamic block header to the codes in use.  Every stream must inflate (zlib and
    include/decoder.h) and never be larger than E-mode's."""
    from zzflate_b200 import synth
    total_e = total_f = 0
    for case in golden.cases:
        data = golden.input(case)
        for level in (2, 3):
            e, *_ = zz.deflate_raw(data, level=level)
            f, a0, crc, st = zz.deflate_raw(data, level=level, mode=1)
            assert zlib.decompress(f, -15) == data, case
            assert zz.ZzFlateDecode(f, zz.Format.Deflate, max_len=len(data) + 16) == data, case
            assert len(f) <= len(e), (case, len(f), len(e))
            assert zz.combine(1, a0, len(data)) == zlib.adler32(data) and crc == zlib.crc32(data)
        total_e += len(e); total_f += len(f)
    assert total_f < total_e
    for n in (40, 80, 150, 260, 300, 400, 600, 1000, 1500):                             # small inputs: stored or dynamic by size
        small = golden.input("alice29")[:n]
        f, *_ = zz.deflate_raw(small, level=2, mode=1)
        assert zlib.decompress(f, -15) == small and len(f) <= len(zz.deflate_raw(small, level=2)[0])
    for lvl in (0, 1):                                                                  # other levels ignore the mode
        assert zz.deflate_raw(golden.input("alice29"), level=lvl, mode=1)[0] == zz.deflate_raw(golden.input("alice29"), level=lvl)[0]
    text = synth.markov_text(64 << 20, seg0=2)
    e, *_ = zz.deflate_raw(text, level=2)
    f, *_ = zz.deflate_raw(text, level=2, mode=1)
    assert zlib.decompress(f, -15) == text.tobytes() and len(f) < len(e)
    print("free mode vs E-mode on 64 MiB of Markov text: %d vs %d bytes (%.4f %%); golden inputs: %d vs %d (%.4f %%)"
          % (len(f), len(e), 100.0 * (len(f) - len(e)) / len(e), total_f, total_e, 100.0 * (total_f - total_e) / total_e))


def test_hold_and_fetch(zz, oracle, golden):
    """zzgpu_deflate_hold / zzgpu_fetch (the stitch of zzflate.cpp:136-154 without a temporary): same bytes as the
    one-call path, to memory and to a sink; a second call of the holding thread is refused until the fetch."""
    import ctypes as C
    from zzflate_b200 import _lib
    lib = _lib.load()
    data = np.frombuffer(golden.input("markov"), dtype=np.uint8)
    want, a0w, crcw, _ = zz.deflate_raw(data, level=2)
    out_len = C.c_size_t(0); a0 = C.c_uint32(0); crc = C.c_uint32(0)
    _lib.check(lib.zzgpu_deflate_hold(data.ctypes.data, data.size, 0, 1, 2, 0, 32768, 3, C.byref(out_len), C.byref(a0), C.byref(crc), None))
    assert out_len.value == len(want) and a0.value == a0w and crc.value == crcw
    dummy = C.c_size_t(0)
    assert lib.zzgpu_deflate_ex(data.ctypes.data, 100, 0, 1, 0, None, 0, 0, 2, 0, 32768, 0, C.byref(dummy), None, None, None) == _lib.E_ARG
    dst = np.zeros(out_len.value + 7, dtype=np.uint8)
    _lib.check(lib.zzgpu_fetch(dst.ctypes.data + 3, out_len.value, _lib.SINK_FN(0), None, 0))
    assert dst[3: 3 + out_len.value].tobytes() == want
    assert lib.zzgpu_fetch(dst.ctypes.data, dst.size, _lib.SINK_FN(0), None, 0) == _lib.E_ARG      # nothing held any more
    got = []
    cb = _lib.SINK_FN(lambda p, n_, u: got.append(C.string_at(p, n_)) or 0)
    _lib.check(lib.zzgpu_deflate_hold(data.ctypes.data, data.size, 0, 1, 2, 0, 32768, 0, C.byref(out_len), None, None, None))
    _lib.check(lib.zzgpu_fetch(None, 0, cb, None, 50000))
    assert b"".join(got) == want and max(map(len, got)) <= 50000
    _lib.check(lib.zzgpu_deflate_hold(data.ctypes.data, data.size, 0, 1, 2, 0, 32768, 0, C.byref(out_len), None, None, None))
    lib.zzgpu_release()
    assert zz.deflate_raw(data, level=2)[0] == want


def test_host_input_in_segments(zz, oracle):
    """Host-buffer calls stage at most one segment (2 GiB by default) on the device; with 64 MiB segments a 200 MiB call
    runs as four, each primed with the 32 KiB before it: same bytes and checksums as one segment."""
    from zzflate_b200 import synth, _lib
    lib = _lib.load()
    n = (200 << 20) + 4321
    data = synth.markov_text(n, seg0=9)
    want, _ = oracle.stream_chunked(data, GZIP, 2, threads=8)
    assert lib.zzgpu_set_option(b"segment_mib", 64) == 0
    try:
        assert zz.ZzFlateEncode(data, zz.Config(zz.Format.Gzip, 2, False)) == want
        pieces = []
        zz.ZzFlateEncodeToCallback(data, zz.Config(zz.Format.Gzip, 2, False), lambda b: pieces.append(b) or False)
        assert b"".join(pieces) == want
        assert zz.adler32x(1, data) == zlib.adler32(data)
    finally:
        assert lib.zzgpu_set_option(b"segment_mib", 2048) == 0


@pytest.mark.parametrize("fmt,ofmt,wbits", [("Zlib", ZLIB, 15), ("Gzip", GZIP, 31)])
def test_threaded_on_several_gpus(zz, oracle, fmt, ofmt, wbits):
    """Config.threaded on a multi-GPU host (zzflate.cpp:97-154): contiguous chunk ranges per device, only the last shard
    with data final, stitched in place, checksums folded -- byte-identical to the single-stream oracle.  Chunk counts
    9, 12 and 17 used to leave trailing empty shards on 4 / 8 GPUs."""
    ndev = zz.device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    from zzflate_b200 import synth
    for nchunks, tail in ((9, 0), (12, 5), (17, 60000), (8 * ndev, 0), (700, 123)):
        data = synth.markov_text(nchunks * S - tail, seg0=3)
        want, _ = oracle.stream_chunked(data, ofmt, 2, threads=8)
        for level in (2, 1, 0):
            w = want if level == 2 else oracle.stream_chunked(data, ofmt, level, threads=8)[0]
            got = zz.ZzFlateEncode(data, zz.Config(zz.Format[fmt], level, True))
            assert got == w, (nchunks, level)
            assert zlib.decompress(got, wbits) == data.tobytes()
        pieces = []
        zz.ZzFlateEncodeToCallback(data, zz.Config(zz.Format[fmt], 2, True), lambda b: pieces.append(b) or False)
        assert b"".join(pieces) == want, nchunks


def test_error_behaviour(zz, golden):
    data = golden.input("alice29")
    assert zz.ZzFlateEncode(data, zz.Config(zz.Format.Zlib, 4, False)) is None            # zzflate.cpp:230
    assert zz.ZzFlateEncode(data, zz.Config(zz.Format.Gzip, 2, False), dest_len=5) is None   # header does not fit
    assert zz.ZzFlateEncode(data, zz.Config(zz.Format.Zlib, 2, False), dest_len=1000) is None   # explicit, not truncated
    got = []
    zz.ZzFlateEncodeToCallback(data, zz.Config(zz.Format.Zlib, 9, False), lambda b: got.append(b) or False)
    assert got == []                                                                       # zzflate.cpp:201-202
    with pytest.raises(zz.ZzGpuError):
        zz.deflate_raw(data, level=2, chunk=1000)
    empty = zz.ZzFlateEncode(b"", zz.Config(zz.Format.Zlib, 2, False))                     # R7 policy
    assert zlib.decompress(empty) == b""


def test_reference_test_sizing_conventions(zz, golden):
    # Test.cpp:147 (1.01x), :209 (max(200,n) at level 1), :254 (n, gzip) -- on compressible input these fit
    data = golden.input("alice29")
    for cfg, cap in [(zz.Config(zz.Format.Zlib, 2, False), int(len(data) * 1.01)),
                     (zz.Config(zz.Format.Zlib, 1, False), max(200, len(data))),
                     (zz.Config(zz.Format.Gzip, 1, False), len(data))]:
        out = zz.ZzFlateEncode(data, cfg, dest_len=cap)
        assert out is not None and zlib.decompress(out, 47) == data


def test_small_zero_buffers(zz):                          # Test.cpp:330-338
    for k in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
        for level in (1, 2):
            out = zz.ZzFlateEncode(bytes(k), zz.Config(zz.Format.Zlib, level, False))
            assert zlib.decompress(out) == bytes(k)


def test_checksums_on_gpu(zz, oracle):
    rng = np.random.default_rng(9)
    for n in (0, 1, 255, 256, 257, 65535, 65536, 65537, 1000003):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert zz.adler32x(1, d) == zlib.adler32(d)
        assert zz.crc32(d) == zlib.crc32(d)
        assert zz.adler32x(0, d) == oracle.adler32(d, 0)
        assert zz.crc32(d, 0x1234) == zlib.crc32(d, 0x1234)
    kat = bytes([0, 1, 23, 30, 4, 69, 145, 32, 216])      # Test.cpp:301-313
    assert zz.adler32x(1, kat) == zz.combine(zz.adler32x(1, kat[:5]), zz.adler32x(0, kat[5:]), 4)
    ff = bytes([0xFF]) * (8 << 20)
    assert zz.adler32x(1, ff) == zlib.adler32(ff)


@pytest.mark.parametrize("chunk,dict_size", [(4096, 2048), (1024, 32768), (8192, 0), (32768, 32768), (65536, 1000)])
def test_other_chunk_and_dictionary_sizes(zz, oracle, golden, chunk, dict_size):
    for case in ("alice29", "pattern", "mixed"):
        data = golden.input(case)[:90000]
        for level in (0, 1, 2):
            out, *_ = zz.deflate_raw(data, level=level, chunk=chunk, dict_size=dict_size)
            assert out == oracle.stream_chunked(data, DEFLATE, level, chunk, dict_size)[0], (case, level)


def test_shards_with_history_concatenate(zz, oracle, golden):
    data = golden.input("markov")
    whole, _ = oracle.stream_chunked(data, DEFLATE, 2)
    for cut in (S, 2 * S):
        first, a1, c1, _ = zz.deflate_raw(data[:cut], level=2, final=False)
        second, a2, c2, _ = zz.deflate_raw(data, level=2, history=cut, final=True)
        assert first + second == whole
        assert zz.combine(zz.combine(1, a1, cut), a2, len(data) - cut) == zlib.adler32(data)
        assert zz.crc32_combine(c1, c2, len(data) - cut) == zlib.crc32(data)


def test_encoder_object(zz, golden):
    import ctypes as C
    from zzflate_b200 import _lib
    lib = _lib.load()
    lib.zz_c_encoder_new.restype = C.c_void_p; lib.zz_c_encoder_new.argtypes = [C.c_int, C.c_void_p, C.c_int64]
    lib.zz_c_encoder_add_data.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.zz_c_encoder_bytes.restype = C.c_size_t; lib.zz_c_encoder_bytes.argtypes = [C.c_void_p]
    lib.zz_c_encoder_data.restype = C.c_void_p; lib.zz_c_encoder_data.argtypes = [C.c_void_p]
    lib.zz_c_encoder_free.argtypes = [C.c_void_p]; lib.zz_c_encoder_set_level.argtypes = [C.c_void_p, C.c_int]
    data = np.frombuffer(golden.input("alice29"), dtype=np.uint8)
    e = lib.zz_c_encoder_new(2, None, 0)
    cut = 70001
    assert lib.zz_c_encoder_add_data(e, data.ctypes.data, cut, 0) == 1
    lib.zz_c_encoder_set_level(e, 1)
    assert lib.zz_c_encoder_add_data(e, data.ctypes.data + cut, data.size - cut, 1) == 1
    n = lib.zz_c_encoder_bytes(e)
    out = C.string_at(lib.zz_c_encoder_data(e), n)
    lib.zz_c_encoder_free(e)
    assert zlib.decompress(out, -15) == data.tobytes()
    # the dictionary is the Encoder's private copy of what it was fed: a second buffer anywhere in memory continues the
    # stream exactly as if the data had been contiguous
    e = lib.zz_c_encoder_new(2, None, 0)
    cut = 2 * S
    first = data[:cut].copy(); second = data[cut:].copy()
    assert lib.zz_c_encoder_add_data(e, first.ctypes.data, cut, 0) == 1
    first[:] = 0                                                     # the caller's earlier buffer is no longer needed
    assert lib.zz_c_encoder_add_data(e, second.ctypes.data, second.size, 1) == 1
    n = lib.zz_c_encoder_bytes(e)
    out = C.string_at(lib.zz_c_encoder_data(e), n)
    lib.zz_c_encoder_free(e)
    from oracle_lib import oracle as get_oracle
    assert out == get_oracle().stream_chunked(data, DEFLATE, 2)[0]


def test_cpp_caller_against_the_headers(zz, golden, tmp_path):
    """tests/cpp/boundary_test.cpp: a C++14 translation unit that includes include/zzflate.h, encoder.h, crc.h,
    huffman.h, outputbitstream.h and links libzzflate_b200.so -- the reference's own tests (Test.cpp:202-338,
    TestHuffman.cpp, TestBitOutput.cpp) re-hosted with plain checks."""
    import subprocess
    from zzflate_b200 import build
    exe = build.build_boundary_test()
    f = tmp_path / "adinsight.bin"
    f.write_bytes(golden.input("adinsight"))
    r = subprocess.run([str(exe), str(f)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "boundary ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_device_resident_buffers(zz, oracle):
    import torch
    from zzflate_b200 import synth
    data = synth.markov_text(3 * S + 777, seg0=1)
    src = torch.from_numpy(data.copy()).cuda()
    dst = torch.empty(zz.bound(data.size), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    n, a0, crc, st = zz.deflate_device(src.data_ptr(), data.size, dst.data_ptr(), dst.numel(), level=2, checksums=3)
    out = dst[:n].cpu().numpy().tobytes()
    assert out == oracle.stream_chunked(data, DEFLATE, 2)[0]
    assert st.kernel_launches >= 5 and st.device_ms > 0
    # unaligned device pointers take the byte-wise window path
    src2 = torch.empty(data.size + 64, dtype=torch.uint8, device="cuda")
    for shift in (1, 7, 13):
        src2[shift: shift + data.size] = src
        torch.cuda.synchronize()
        n2, *_ = zz.deflate_device(src2.data_ptr() + shift, data.size, dst.data_ptr() + 3, dst.numel() - 3, level=2)
        assert dst[3: 3 + n2].cpu().numpy().tobytes() == out


def test_regressions_found_by_fuzzing(zz, oracle):
    """tests/golden/regress/*.bin: inputs on which tools/gpu_fuzz.py once found a mismatch (file name carries level,
    chunk and dictionary size).  fuzz_1_3613: a match of >= 32 bytes whose backward extension is so long that it ends
    inside the tile of the state it was taken from (R6 territory, lb = 258).  lzwalk_*: inputs on which the lanes of K-LZ's walking warp ran as
    two groups and the slower one missed a join the faster one had recorded (found by tools/gpu_lz_check.py)."""
    import re
    from pathlib import Path
    files = sorted((Path(__file__).parent / "golden" / "regress").glob("*.bin"))
    assert files
    for f in files:
        m = re.search(r"_l(\d)_c(\d+)_d(\d+)", f.name)
        level, chunk, dict_size = int(m.group(1)), int(m.group(2)), int(m.group(3))
        data = f.read_bytes()
        got, *_ = zz.deflate_raw(data, level=level, chunk=chunk, dict_size=dict_size)
        assert got == oracle.stream_chunked(data, DEFLATE, level, chunk, dict_size)[0], f.name
        assert zlib.decompress(got, -15) == data
        got2, *_ = zz.deflate_raw(data, level=level)
        assert got2 == oracle.stream_chunked(data, DEFLATE, level)[0], f.name


def test_short_fuzz_run(zz):
    """20 s of tools/gpu_fuzz.py (structured random inputs, both levels, default and odd geometries) against the oracle;
    longer runs (3 x 240 s, 118 000 cases) were clean after the fix recorded in tests/golden/regress."""
    import subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "tools" / "gpu_fuzz.py"), "20", "12345"], capture_output=True, text=True, cwd=str(root))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout


def test_host_buffers_pipelined_path(zz, oracle):
    """Inputs >= 32 MiB on host buffers go through the piece-wise path (H2D / kernels / D2H overlapped on three
    streams); the bytes must not depend on how the call was cut into pieces."""
    import ctypes as C
    from zzflate_b200 import synth, _lib
    n = (96 << 20) + 12345
    data = synth.markov_text(n, seg0=7)
    data[40 << 20: 41 << 20] = 0                                 # a stretch of zeros and of noise inside the text
    data[41 << 20: 42 << 20] = synth.random_bytes(1 << 20)
    want, _ = oracle.stream_chunked(data, ZLIB, 2, threads=8)
    for threaded in (False, True):
        got = zz.ZzFlateEncode(data, zz.Config(zz.Format.Zlib, 2, threaded))
        assert got == want
    lib = _lib.load()
    out, a0, crc, st = zz.deflate_raw(data, level=2)
    assert out == want[2:-4]
    assert zz.combine(1, a0, n) == zlib.adler32(data) and crc == zlib.crc32(data)
    g1 = zz.ZzFlateEncode(data[: 40 << 20], zz.Config(zz.Format.Gzip, 1, False))
    assert zlib.decompress(g1, 31) == data[: 40 << 20].tobytes()


def test_kernel_variants_give_the_same_bytes(zz, oracle, golden):
    """K-HUFF has two variants (warp per chunk for launches of one wave, thread per chunk for full batches, picked by
    launch size) and K-LZ two A/B switches (window by TMA bulk copy or by LDG/STS; speculative chains or the true walk
    alone): every combination must give the oracle's bytes."""
    from zzflate_b200 import synth, _lib
    lib = _lib.load()
    small = golden.input("mixed")                                # a few chunks: warp-per-chunk K-HUFF
    n = 8000 * S + 777                                           # > 148 * 52 chunks: thread-per-chunk K-HUFF
    big = synth.markov_text(n, seg0=3)
    big[5 * S: 6 * S] = synth.random_bytes(S)                    # a stored chunk, a run of zeros and few-symbol data in between
    big[9 * S: 9 * S + 3000] = 0
    big[11 * S: 12 * S] = synth.random_bytes(S) % 7 + 48
    want_small = oracle.stream_chunked(small, DEFLATE, 2)[0]
    want_big = oracle.stream_chunked(big, DEFLATE, 2, threads=8)[0]
    try:
        for tma, spec in ((1, 1), (0, 1), (1, 0), (0, 0)):
            assert lib.zzgpu_set_option(b"tma", tma) == 0 and lib.zzgpu_set_option(b"spec", spec) == 0
            assert zz.deflate_raw(small, level=2)[0] == want_small, (tma, spec)
            for case in ("pattern", "zeros", "kennedy", "ptt5"):
                data = golden.input(case)
                assert zz.deflate_raw(data, level=2)[0] == oracle.stream_chunked(data, DEFLATE, 2)[0], (case, tma, spec)
            if spec:                                             # (the true walk alone is slow on a large input)
                assert zz.deflate_raw(big, level=2)[0] == want_big, (tma, spec)
    finally:
        lib.zzgpu_set_option(b"tma", 1); lib.zzgpu_set_option(b"spec", 1)
    assert lib.zzgpu_set_option(b"no-such-option", 1) == _lib.E_ARG


@pytest.mark.parametrize("workload,size_mib,level", [("text", 1024, 2), ("text", 1088, 2), ("random", 1024, 2), ("zeros", 1024, 2),
                                                     ("pattern", 1024, 2), ("random", 1024, 1), ("text", 1024, 1), ("random", 1024, 0)])
def test_baseline_sizes_bit_exact(zz, workload, size_mib, level):
    """BASELINE.json configs 2-4 at full size (1 GiB; 1088 MiB = 17 408 chunks crosses the 16 384-chunk scratch batch):
    the whole stream equals the (multi-threaded) oracle's byte for byte, inflates through zlib to the input, and the
    checksums agree with zlib's."""
    import os
    import torch
    from oracle_lib import oracle as get_oracle
    from zzflate_b200 import synth
    n = size_mib << 20
    data = synth.workload(workload, n)
    src = torch.from_numpy(data).cuda()
    dst = torch.empty(zz.bound(n, level), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    out_len, a0, crc, st = zz.deflate_device(src.data_ptr(), n, dst.data_ptr(), dst.numel(), level=level, checksums=3)
    out = dst[:out_len].cpu().numpy()
    del src, dst
    assert zz.combine(1, a0, n) == zlib.adler32(data)
    assert crc == zlib.crc32(data)
    want, _ = get_oracle().stream_chunked(data, DEFLATE, level, threads=os.cpu_count() or 8)
    assert out_len == len(want)
    assert out.tobytes() == want
    d = zlib.decompressobj(-15)
    pos = 0
    step = 64 << 20
    for lo in range(0, out_len, step):
        piece = d.decompress(out[lo: lo + step].tobytes())
        assert piece == data[pos: pos + len(piece)].tobytes()
        pos += len(piece)
    assert pos == n and d.eof
