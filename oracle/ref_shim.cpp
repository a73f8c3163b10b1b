// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI shim around the UNMODIFIED reference sources under /root/reference/zzflate.
// It is compiled together with those sources (where they lie) into oracle/_ref/libzzref.so
// by oracle/Makefile.  Nothing of the reference is copied into this repository: this
// file only *calls* the reference through its own headers.
//
// The reference keeps its tokens / code tables / hash table private (encoder.h:65-77);
// the survey's recipe (`#define private public`) is used to reach them so that the
// restatement in oracle/zz_oracle.c can be pinned stage by stage, not just on final bytes.
//
// Entry points (all `zzref_*`):
//   zzref_encode          whole-stream ZzFlateEncode           (zzflate.cpp:225)
//   zzref_encode_callback whole-stream ZzFlateEncodeToCallback (zzflate.cpp:197)
//   zzref_chunk_encode    E(c) of SURVEY A.7: fresh Encoder + dictionary priming + alignment block
//   zzref_chunk_tokens    as above, also dumps comprecords / code tables of the last block
//   zzref_calc_lengths    CalcLengths      (huffman.cpp:122)
//   zzref_from_lengths    FromLengths      (huffman.cpp:191)
//   zzref_generate        huffman::generate (huffman.h:49)
//   zzref_adler32x / zzref_combine / zzref_crc32   (adler.cpp:17,5 ; crc.cpp:24)
//   zzref_find_distance / zzref_read_lut            (encoder.cpp:51 ; encoder.h:93)
//   zzref_merged_length_codes                       (encoder.cpp:126)
//   zzref_bitstream_kat   drives outputbitstream    (outputbitstream.h:83-124)

#include <cstdint>
#include <cstddef>
#include <cstring>
#include <vector>
#include <string>
#include <memory>
#include <functional>
#include <algorithm>
#include <iterator>
#include <cassert>

#define private public
#include "encoder.h"
#undef private
#include "zzflate.h"
#include "crc.h"

#define ZZREF_API extern "C" __attribute__((visibility("default")))

namespace {
// Same arithmetic as Encoder::CalcHash (encoder.cpp:11-17); that function is `inline` in the
// reference's .cpp and therefore not linkable from here.  Used only for the level-1 priming loop,
// which the reference does not have at all (it has no dictionary API).
inline unsigned refhash(const uint8_t* p)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return (((v << 8) >> 8) * 0x00d68664u) >> (32 - 13);
}
}

ZZREF_API size_t zzref_encode(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                              int format, int level, int threaded)
{
    Config cfg = { (Format)format, (uint8_t)level, threaded != 0 };
    size_t len = cap;
    ZzFlateEncode(dest, &len, src, n, &cfg);
    return len;
}

ZZREF_API size_t zzref_encode_callback(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                                       int format, int level, int threaded)
{
    Config cfg = { (Format)format, (uint8_t)level, threaded != 0 };
    size_t total = 0;
    bool overflow = false;
    ZzFlateEncodeToCallback(src, n, &cfg, [&](const uint8_t* b, size_t c) -> bool {
        if (total + c > cap) { overflow = true; return false; }
        memcpy(dest + total, b, c);
        total += c;
        return false;
    });
    return overflow ? ~(size_t)0 : total;
}

// One chunk in reference-equivalent mode.  `chunk` must point INTO the full
// contiguous input (history readable before it, >= 8 readable bytes after chunk+n).
// level 2/3: AddHashEntries(chunk, -dict, dict) primes the table (L2 convention table[h(k)] = k).
// level 1  : the loop below primes with the L1 convention table[h(i+1)] = i (encoder.cpp:344-346).
static size_t chunk_encode(Encoder& e, const uint8_t* chunk, size_t n, size_t dict, int level, int final)
{
    if (dict > 0) {
        if (level >= 2) {
            e.AddHashEntries(chunk, -(int)dict, (int)dict);
        } else if (level == 1) {
            for (int i = -(int)dict; i < 0; ++i)
                e.hashtable[refhash(chunk + i + 1)] = i;
        }
    }
    if (final) {
        e.AddData(chunk, chunk + n, true);
    } else {
        e.AddData(chunk, chunk + n - 1, false);
        e.SetLevel(0);
        e.AddData(chunk + n - 1, chunk + n, false);
    }
    e.stream.Flush();
    return e.stream.byteswritten();
}

ZZREF_API size_t zzref_chunk_encode(const uint8_t* chunk, size_t n, size_t dict, int level, int final,
                                    uint8_t* out, size_t cap)
{
    auto e = std::make_unique<Encoder>(level, out, (int64_t)cap);
    return chunk_encode(*e, chunk, n, dict, level, final);
}

// Token dump of the (single) level>=2 block of a chunk: records as (literals, backoffset, length)
// triples of uint32, plus the literal/length and distance code tables as (length,bits) pairs.
// Returns the number of records, or -1 if the chunk did not go through WriteBlock2Pass.
ZZREF_API int zzref_chunk_tokens(const uint8_t* chunk, size_t n, size_t dict, int level, int final,
                                 uint8_t* out, size_t cap, size_t* outLen,
                                 uint32_t* records, int maxRecords,
                                 int32_t* litCodes /*286*2*/, int32_t* distCodes /*30*2*/)
{
    if (level < 2) return -1;
    auto e = std::make_unique<Encoder>(level, out, (int64_t)cap);
    for (auto& c : e->codes) { c.length = 0; c.bits = 0; }
    for (auto& c : e->dcodes) { c.length = 0; c.bits = 0; }
    e->validRecords = 0;
    *outLen = chunk_encode(*e, chunk, n, dict, level, final);
    int count = e->validRecords;
    for (int i = 0; i < count && i < maxRecords; ++i) {
        records[3 * i + 0] = e->comprecords[i].literals;
        records[3 * i + 1] = e->comprecords[i].backoffset;
        records[3 * i + 2] = e->comprecords[i].length;
    }
    for (int i = 0; i < 286; ++i) { litCodes[2 * i] = e->codes[i].length; litCodes[2 * i + 1] = (int32_t)e->codes[i].bits; }
    for (int i = 0; i < 30; ++i) { distCodes[2 * i] = e->dcodes[i].length; distCodes[2 * i + 1] = (int32_t)e->dcodes[i].bits; }
    return count;
}

ZZREF_API void zzref_calc_lengths(const int* freqs, int n, int maxLength, int* lengthsOut)
{
    std::vector<int> f(freqs, freqs + n), l;
    CalcLengths(f, l, maxLength);
    for (int i = 0; i < n; ++i) lengthsOut[i] = l[i];
}

// RLE of a code-length sequence; records written as (value,payLoad) byte pairs; freqs[19] accumulated.
ZZREF_API int zzref_from_lengths(const int* lengths, int n, int* freqs19, uint8_t* recordsOut, int maxRecords)
{
    std::vector<int> l(lengths, lengths + n), f(freqs19, freqs19 + 19);
    auto recs = FromLengths(l, f);
    for (int i = 0; i < 19; ++i) freqs19[i] = f[i];
    int count = (int)recs.size();
    for (int i = 0; i < count && i < maxRecords; ++i) {
        recordsOut[2 * i] = recs[i].value;
        recordsOut[2 * i + 1] = recs[i].payLoad;
    }
    return count;
}

ZZREF_API void zzref_generate(const int* lengths, int n, int32_t* codesOut /* n*2: length,bits */)
{
    std::vector<int> l(lengths, lengths + n);
    std::vector<code> c(n);
    for (auto& x : c) { x.length = 0; x.bits = 0; }
    huffman::generate<code>(l, &c[0]);
    for (int i = 0; i < n; ++i) { codesOut[2 * i] = c[i].length; codesOut[2 * i + 1] = (int32_t)c[i].bits; }
}

ZZREF_API void zzref_default_table_lengths(int* out288)
{
    auto l = huffman::defaultTableLengths();
    for (int i = 0; i < 288; ++i) out288[i] = l[i];
}

ZZREF_API unsigned zzref_reverse(unsigned value, int len) { return huffman::reverse(value, len); }

ZZREF_API uint32_t zzref_adler32x(uint32_t start, const uint8_t* d, size_t n) { return adler32x(start, d, n); }
ZZREF_API uint32_t zzref_combine(uint32_t a, uint32_t b, size_t lenB) { return combine(a, b, lenB); }
ZZREF_API uint32_t zzref_crc32(const uint8_t* d, size_t n, uint32_t start) { return crc32(d, n, start); }

ZZREF_API int zzref_find_distance(int offset) { return Encoder::FindDistance(offset); }
ZZREF_API int zzref_read_lut(int offset) { return Encoder::ReadLut(offset); }

// Static tables, for pinning the procedurally generated ones in the restatement.
ZZREF_API void zzref_tables(int16_t* lengthCode259, int8_t* lengthExtra259, int8_t* lengthExtraBits259,
                            uint8_t* order19, uint8_t* extraDist30, uint8_t* extraLen286, uint16_t* distBase30,
                            int32_t* codesF286x2, int32_t* lcodesF259x2, int32_t* dcodesF30x2)
{
    for (int i = 0; i < 259; ++i) {
        lengthCode259[i] = Encoder::lengthTable[i].code;
        lengthExtra259[i] = Encoder::lengthTable[i].extraBits;
        lengthExtraBits259[i] = Encoder::lengthTable[i].extraBitLength;
        lcodesF259x2[2 * i] = Encoder::lcodes_f[i].length; lcodesF259x2[2 * i + 1] = (int32_t)Encoder::lcodes_f[i].bits;
    }
    for (int i = 0; i < 19; ++i) order19[i] = Encoder::order[i];
    for (int i = 0; i < 30; ++i) {
        extraDist30[i] = Encoder::extraDistanceBits[i];
        distBase30[i] = Encoder::distanceTable[i];
        dcodesF30x2[2 * i] = Encoder::dcodes_f[i].length; dcodesF30x2[2 * i + 1] = (int32_t)Encoder::dcodes_f[i].bits;
    }
    for (int i = 0; i < 286; ++i) {
        extraLen286[i] = Encoder::extraLengthBits[i];
        codesF286x2[2 * i] = Encoder::codes_f[i].length; codesF286x2[2 * i + 1] = (int32_t)Encoder::codes_f[i].bits;
    }
}

ZZREF_API void zzref_merged_length_codes(const int32_t* symbolCodes286x2, int32_t* lcodes259x2)
{
    code sym[286], l[259];
    for (int i = 0; i < 286; ++i) { sym[i].length = symbolCodes286x2[2 * i]; sym[i].bits = (uint32_t)symbolCodes286x2[2 * i + 1]; }
    Encoder::CreateMergedLengthCodes(l, sym);
    for (int i = 0; i < 259; ++i) { lcodes259x2[2 * i] = l[i].length; lcodes259x2[2 * i + 1] = (int32_t)l[i].bits; }
}

// Bit-writer known-answer driver: appends (bits,count) pairs, optionally flushes, returns bytes written
// after Flush (or 0 before).  Mirrors zztest/TestBitOutput.cpp:7-36.
ZZREF_API size_t zzref_bitstream_kat(const uint64_t* bits, const int* counts, int n, int flush, uint8_t* buf, size_t cap)
{
    outputbitstream s(buf, cap);
    for (int i = 0; i < n; ++i) s.AppendToBitStream(bits[i], counts[i]);
    if (!flush) return 0;
    s.Flush();
    return s.byteswritten();
}
