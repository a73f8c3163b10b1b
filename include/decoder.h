/* decoder.h -- a working inflater in the place the reference reserved for one (zzflate/decoder.h:4-33 is a 32-line stub:
 * private constructor, Peek16Bits() returns 0, ReadCode() returns nothing).  It keeps the stub's design -- a bit
 * stream that is peeked 16 bits at a time and two 65 536-entry tables (literal/length codes, distance codes) indexed by
 * the peeked bits -- and completes it: stored, fixed and dynamic blocks (RFC 1951), zlib / gzip framing with
 * Adler-32 / CRC-32 / ISIZE checks (RFC 1950 / 1952), concatenated gzip members.  Host code: it is the second,
 * zlib-independent judge of the encoder's streams (tests and bench.py inflate every output through zlib AND through
 * this class), not part of the encode hot path.
 */
#ifndef ZZFLATE_B200_DECODER_H
#define ZZFLATE_B200_DECODER_H

#include <stddef.h>
#include <stdint.h>
#include <vector>
#include "zzflate.h"

class Decoder
{
public:
    enum Status { Ok = 0, NeedMoreOutput = -1, BadHeader = -2, BadBlock = -3, BadCode = -4, BadDistance = -5,
                  BadChecksum = -6, Truncated = -7 };

    Decoder();

    /* Inflates one complete stream (for Gzip: all concatenated members).  `dict` (optional) is history that precedes the
     * stream, as for a shard of a larger stream.  Returns Ok and *destLen = bytes produced, or an error. */
    Status Inflate(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, Format format,
                   const uint8_t* dict = nullptr, size_t dictLen = 0);

    struct code { uint8_t length; uint16_t symbol; };      /* decoder.h:13: length of the code, symbol it decodes to */

private:
    class BitStream                                        /* decoder.h:18-22 */
    {
    public:
        void Reset(const uint8_t* p, size_t n) { cur = p; end = p + n; acc = 0; bits = 0; overrun = false; }
        uint16_t Peek16Bits() { Fill(); return (uint16_t)acc; }
        uint32_t Peek(int n) { Fill(); return (uint32_t)(acc & ((1ull << n) - 1)); }
        void Skip(int n) { Fill(); if (n > bits) { overrun = true; n = bits; } acc >>= n; bits -= n; }
        uint32_t Read(int n) { const uint32_t v = Peek(n); Skip(n); return v; }
        void AlignToByte() { Skip(bits & 7); }
        /* byte-aligned raw access (stored blocks, trailers) */
        bool ReadBytes(uint8_t* out, size_t n);
        size_t Consumed(const uint8_t* base) const { return (size_t)(cur - base) - (size_t)(bits >> 3); }
        bool overrun;
    private:
        void Fill() { while (bits <= 56 && cur < end) { acc |= (uint64_t)*cur++ << bits; bits += 8; } }
        const uint8_t* cur; const uint8_t* end; uint64_t acc; int bits;
    };

    bool BuildTable(const uint8_t* lengths, int n, std::vector<code>& table);
    Status InflateBlocks(uint8_t* dest, size_t cap, size_t& pos, const uint8_t* dict, size_t dictLen);
    code ReadCode(const std::vector<code>& table);         /* decoder.h:24-30 */

    std::vector<code> symbolCodes;                         /* decoder.h:16: literal/length codes by peeked bits */
    std::vector<code> lengthCodes;                         /* decoder.h:15: here the distance codes */
    BitStream inputStream;
};

/* Convenience wrapper in the style of ZzFlateEncode: *destLen is capacity on entry and bytes produced on return (~0 on
 * error). */
void ZzFlateDecode(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, Format format);

#endif
