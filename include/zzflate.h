/* zzflate.h -- public API of the B200-native encoder.
 *
 * Same entry points, argument meaning and error convention as the reference's zzflate/zzflate.h:8-19, so
 * existing callers link unchanged; the work behind them runs in sm_100a kernels reached through the
 * C-ABI of zzgpu.h.  There is no CPU fallback: without a usable CUDA device both functions fail
 * (*destLen = ~0 / no callback invocations).
 *
 *   Config.format    Zlib (78 01 + Adler-32 BE), Gzip (10-byte header + CRC-32 LE + ISIZE LE), Deflate (raw)
 *   Config.level     0 stored, 1 fixed Huffman, 2 and 3 dynamic Huffman (identical, encoder.cpp:508-524);
 *                    anything above 3 is an error (zzflate.cpp:230)
 *   Config.threaded  reference: fan out over hardware_concurrency() host threads (zzflate.cpp:97-132);
 *                    here: fan out over every visible GPU, one contiguous range of chunks per device
 *
 * Differences from the reference, all on inputs where the reference itself misbehaves:
 *   - a destination that is too small yields *destLen = ~0 instead of a silently truncated stream;
 *   - empty input yields one final empty stored block instead of no block at all (R7);
 *   - the Zlib trailer is the true Adler-32 also above ~362 MiB (R3).
 */
#ifndef ZZFLATE_B200_ZZFLATE_H
#define ZZFLATE_B200_ZZFLATE_H

#include <stddef.h>
#include <stdint.h>
#include <functional>

enum Format { Zlib, Gzip, Deflate };

struct Config
{
    Format format;
    uint8_t level;
    bool threaded;
};

/* dest/destLen: caller-owned buffer, *destLen is capacity on entry and bytes written on return
 * (~0 on error).  Replaces zzflate/zzflate.cpp:225-242. */
void ZzFlateEncode(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, const Config* config);

/* Header, then the compressed stream in pieces of at most 1 000 000 bytes (the reference's buffer size,
 * outputbitstream.h:183), then the trailer, in order; pointers are valid only during the call and the
 * callback's return value is ignored, as in zzflate/zzflate.cpp:197-222. */
void ZzFlateEncodeToCallback(const uint8_t* source, size_t sourceLen, const Config* config,
                             std::function<bool(const uint8_t*, size_t)> callback);

#endif
