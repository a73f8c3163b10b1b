"""N>1 host logic on CPU: two gloo ranks each encode their shard (with the oracle standing in for the
kernels), rank 0 stitches and folds the checksums; the result must equal the single-process stream."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zzflate_b200 import shard, synth

S, D = 65536, 32768


def test_shard_ranges_cover_input():
    for n in (0, 1, 65535, 65536, 65537, 1 << 20, (1 << 20) + 7):
        for world in (1, 2, 3, 8):
            r = shard.shard_ranges(n, world)
            assert len(r) == world and sum(l for _, l in r) == n
            pos = 0
            for off, ln in r:
                assert off == pos
                assert off % S == 0 or ln == 0
                pos += ln


def _worker(rank, world, port, n, result):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
    from oracle_lib import oracle, _padded
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = synth.markov_text(n, threads=1)
    buf = _padded(data)
    o = oracle()
    ranges = shard.shard_ranges(n, world)
    off, ln = ranges[rank]
    last = shard.last_rank_with_data(ranges)
    parts = []
    for c in range(off, off + ln, S):
        cl = min(S, off + ln - c)
        parts.append(o.chunk_encode(buf, c, cl, min(D, c), 2, rank == last and c + cl == off + ln)["bytes"])
    mine = b"".join(parts)
    piece = data[off: off + ln].tobytes()
    meta = torch.tensor([len(mine), o.adler32(piece, 0), zlib.crc32(piece), ln], dtype=torch.int64)
    metas = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(metas, meta)
    maxlen = max(int(m[0]) for m in metas)
    padded = torch.zeros(maxlen, dtype=torch.uint8); padded[: len(mine)] = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    bufs = [torch.zeros(maxlen, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0)
    if rank == 0:
        stream = shard.stitch([bytes(b[: int(m[0])].numpy()) for b, m in zip(bufs, metas)])
        adler, crc = 1, 0
        for m in metas:
            adler = o.combine(adler, int(m[1]), int(m[3])); crc = o.crc32_combine(crc, int(m[2]), int(m[3]))
        single, _ = o.stream_chunked(data, 2, 2)
        result.put((stream == single, zlib.decompress(stream, -15) == data.tobytes(),
                    adler == zlib.adler32(data.tobytes()), crc == zlib.crc32(data.tobytes())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shards_stitch_to_single_stream():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5 * S + 1234, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
    assert res == (True, True, True, True)
