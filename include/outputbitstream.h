/* outputbitstream.h -- host-side bit sink with the reference's public surface (zzflate/outputbitstream.h:14-24,47-160):
 * `code` (an LSB-first bit string) and an `outputbitstream` over a caller-owned buffer.  The GPU emitter (K-EMIT)
 * produces the same layout on the device; this host class exists so that callers and tests written against the
 * reference's bit writer (zztest/TestBitOutput.cpp:7-36) compile and behave the same:
 *   - bits are appended LSB-first into a 64-bit accumulator (outputbitstream.h:83-98);
 *   - memory is written in whole little-endian 64-bit words only, so nothing reaches the buffer before 64 bits
 *     have accumulated or Flush() is called (TestBitOutput.cpp:17: buffer[0] == 0 before Flush);
 *   - Flush() pads to a byte boundary and drains the accumulator bytewise (outputbitstream.h:100-124).
 * New code (written from SURVEY A.1), host only, no growable-buffer mode: the callback API streams from the GPU
 * (zzgpu_deflate_sink) instead. */
#ifndef ZZFLATE_B200_OUTPUTBITSTREAM_H
#define ZZFLATE_B200_OUTPUTBITSTREAM_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

struct code          /* outputbitstream.h:14-24 */
{
    code() = default;
    code(int length_, uint32_t bits_) : length(length_), bits(bits_) {}
    int32_t length;
    uint32_t bits;
};

struct outputbitstream
{
    outputbitstream(uint8_t* buffer, size_t byteCount) : start(buffer), pos(buffer), end(buffer + byteCount) {}
    outputbitstream(const outputbitstream&) = delete;

    void AppendToBitStream(code c) { AppendToBitStream(c.bits, c.length); }
    void AppendToBitStream(uint64_t bits, int32_t bitCount)          /* bitCount <= 57 per call, as in the reference */
    {
        acc |= bits << used;
        used += bitCount;
        if (used >= 64) {
            storeWord(acc);
            used -= 64;
            acc = used ? bits >> (bitCount - used) : 0;
        }
    }
    void PadToByte() { if (used & 7) AppendToBitStream(0, 8 - (used & 7)); }
    void Flush()
    {
        PadToByte();
        while (used > 0) { if (pos < end) *pos++ = (uint8_t)acc; acc >>= 8; used -= 8; }
        acc = 0; used = 0;
    }
    void WriteU8(uint8_t v) { Flush(); if (pos < end) *pos++ = v; }
    bool WriteU16(uint16_t v) { Flush(); if (end - pos < 2) return false; pos[0] = (uint8_t)v; pos[1] = (uint8_t)(v >> 8); pos += 2; return true; }
    void WriteU32(uint32_t v) { WriteU16((uint16_t)v); WriteU16((uint16_t)(v >> 16)); }
    void WriteBigEndianU32(uint32_t v) { WriteU8((uint8_t)(v >> 24)); WriteU8((uint8_t)(v >> 16)); WriteU8((uint8_t)(v >> 8)); WriteU8((uint8_t)v); }
    void WriteBytes(const uint8_t* source, int length)
    {
        Flush();
        const size_t room = (size_t)(end - pos), n = (size_t)length < room ? (size_t)length : room;
        memcpy(pos, source, n); pos += n;
    }
    uint8_t* streamStart() { return start; }
    size_t byteswritten() const { return (size_t)(pos - start); }
    uint64_t BitsWritten() const { return (uint64_t)(pos - start) * 8 + (uint64_t)used; }

private:
    void storeWord(uint64_t w) { for (int i = 0; i < 8 && pos < end; ++i) { *pos++ = (uint8_t)w; w >>= 8; } }
    uint8_t* start; uint8_t* pos; uint8_t* end;
    uint64_t acc = 0;
    int used = 0;
};

#endif
