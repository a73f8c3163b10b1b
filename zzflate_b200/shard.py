"""Host-side sharding of one stream over ranks / GPUs: contiguous ranges of chunks, no
data-path collective.  Every chunk ends byte-aligned, so the stream is the concatenation of the ranks'
outputs in rank order; only the globally last chunk carries BFINAL; checksums fold in O(ranks).
Mirrors the partitioning of zzflate/zzflate.cpp:67-78,97-99 (divideInRanges) at chunk granularity."""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import api


def shard_ranges(n: int, world: int, chunk: int = api.DEFAULT_CHUNK) -> List[Tuple[int, int]]:
    """(offset, length) per rank; lengths are whole chunks except for the stream tail; may be (x, 0)."""
    nchunks = (n + chunk - 1) // chunk
    per = (nchunks + world - 1) // world if world else 0
    out = []
    for r in range(world):
        c0, c1 = min(nchunks, per * r), min(nchunks, per * (r + 1))
        off = min(n, c0 * chunk)
        out.append((off, min(n, c1 * chunk) - off))
    return out


def last_rank_with_data(ranges: Sequence[Tuple[int, int]]) -> int:
    idx = [i for i, (_, ln) in enumerate(ranges) if ln > 0]
    return idx[-1] if idx else 0


def fold_checksums(parts: Sequence[Tuple[int, int, int]]) -> Tuple[int, int]:
    """parts: (adler32 with start value 0, crc32, length) per rank in order -> (Adler-32, CRC-32) of the stream."""
    adler, crc = 1, 0
    for a0, c, ln in parts:
        adler = api.combine(adler, a0, ln)
        crc = api.crc32_combine(crc, c, ln)
    return adler, crc


def stitch(parts: Sequence[bytes]) -> bytes:
    return b"".join(parts)
