// Host inflater behind include/decoder.h (the role zzflate/decoder.h:4-33 reserved): RFC 1951 blocks, zlib / gzip framing.
// Written from the RFCs; table layout as the reference's stub sketches it (peek 16 bits, index a 65 536-entry table).
#include "../../include/decoder.h"
#include "../../include/zzgpu.h"

#include <cstring>

namespace {

const uint16_t kLenBase[29] = { 3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258 };
const uint8_t kLenExtra[29] = { 0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0 };
const uint16_t kDistBase[30] = { 1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577 };
const uint8_t kDistExtra[30] = { 0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13 };
const uint8_t kOrder[19] = { 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15 };

uint32_t adler32_host(uint32_t adler, const uint8_t* p, size_t n)
{
    uint32_t a = adler & 0xFFFF, b = adler >> 16;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521; b %= 65521;
    }
    return (b << 16) | a;
}

uint32_t crc32_host(uint32_t crc, const uint8_t* p, size_t n)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int j = 0; j < 8; ++j) c = (c >> 1) ^ ((c & 1) * 0xEDB88320u); table[i] = c; }
        init = true;
    }
    crc = ~crc;
    while (n--) crc = (crc >> 8) ^ table[(crc ^ *p++) & 0xFF];
    return ~crc;
}

}  // namespace

Decoder::Decoder() : symbolCodes(65536), lengthCodes(65536) {}

bool Decoder::BitStream::ReadBytes(uint8_t* out, size_t n)
{
    while (n && bits >= 8) { *out++ = (uint8_t)acc; acc >>= 8; bits -= 8; --n; }
    if ((size_t)(end - cur) < n) { overrun = true; return false; }
    memcpy(out, cur, n); cur += n;
    return true;
}

// Canonical code (RFC 1951 3.2.2) spread over a table indexed by the next 16 stream bits (LSB-first, so the code's bits are
// reversed).  Codes are at most 15 bits long.  Returns false for an over-subscribed set of lengths; an incomplete set is
// accepted (zlib does the same for a single distance code), unused entries decode to length 0 = invalid.
bool Decoder::BuildTable(const uint8_t* lengths, int n, std::vector<code>& table)
{
    int count[16] = { 0 };
    for (int i = 0; i < n; ++i) count[lengths[i]]++;
    count[0] = 0;
    unsigned next[16], c = 0;
    long left = 1;
    for (int b = 1; b < 16; ++b) { left = left * 2 - count[b]; if (left < 0) return false; c = (c + (unsigned)count[b - 1]) << 1; next[b] = c; }
    for (auto& e : table) { e.length = 0; e.symbol = 0; }
    for (int i = 0; i < n; ++i) {
        const int len = lengths[i];
        if (!len) continue;
        unsigned v = next[len]++, rev = 0;
        for (int k = 0; k < len; ++k) rev |= ((v >> k) & 1u) << (len - 1 - k);
        for (unsigned idx = rev; idx < 65536u; idx += 1u << len) { table[idx].length = (uint8_t)len; table[idx].symbol = (uint16_t)i; }
    }
    return true;
}

Decoder::code Decoder::ReadCode(const std::vector<code>& table)
{
    const uint16_t value = inputStream.Peek16Bits();
    const code c = table[value];
    inputStream.Skip(c.length);
    return c;
}

Decoder::Status Decoder::InflateBlocks(uint8_t* dest, size_t cap, size_t& pos, const uint8_t* dict, size_t dictLen)
{
    const size_t start = pos;
    for (;;) {
        const uint32_t final = inputStream.Read(1), type = inputStream.Read(2);
        if (inputStream.overrun) return Truncated;
        if (type == 0) {
            inputStream.AlignToByte();
            uint8_t h[4];
            if (!inputStream.ReadBytes(h, 4)) return Truncated;
            const unsigned len = h[0] | (h[1] << 8), nlen = h[2] | (h[3] << 8);
            if ((len ^ 0xFFFFu) != nlen) return BadBlock;
            if (pos + len > cap) return NeedMoreOutput;
            if (!inputStream.ReadBytes(dest + pos, len)) return Truncated;
            pos += len;
        } else if (type == 1 || type == 2) {
            uint8_t lens[320];
            if (type == 1) {
                for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                BuildTable(lens, 288, symbolCodes);
                for (int i = 0; i < 30; ++i) lens[i] = 5;
                BuildTable(lens, 30, lengthCodes);
            } else {
                const int hlit = (int)inputStream.Read(5) + 257, hdist = (int)inputStream.Read(5) + 1, hclen = (int)inputStream.Read(4) + 4;
                if (hlit > 286 || hdist > 30) return BadBlock;
                uint8_t cl[19] = { 0 };
                for (int i = 0; i < hclen; ++i) cl[kOrder[i]] = (uint8_t)inputStream.Read(3);
                if (!BuildTable(cl, 19, symbolCodes)) return BadCode;
                int n = 0;
                while (n < hlit + hdist) {
                    const code c = ReadCode(symbolCodes);
                    if (c.length == 0 || inputStream.overrun) return inputStream.overrun ? Truncated : BadCode;
                    if (c.symbol < 16) lens[n++] = (uint8_t)c.symbol;
                    else {
                        int rep; uint8_t v = 0;
                        if (c.symbol == 16) { if (n == 0) return BadCode; v = lens[n - 1]; rep = 3 + (int)inputStream.Read(2); }
                        else if (c.symbol == 17) rep = 3 + (int)inputStream.Read(3);
                        else rep = 11 + (int)inputStream.Read(7);
                        if (n + rep > hlit + hdist) return BadCode;
                        while (rep--) lens[n++] = v;
                    }
                }
                uint8_t dl[30];
                memcpy(dl, lens + hlit, (size_t)hdist);
                if (!BuildTable(lens, hlit, symbolCodes)) return BadCode;
                if (!BuildTable(dl, hdist, lengthCodes)) return BadCode;
            }
            for (;;) {
                const code c = ReadCode(symbolCodes);
                if (c.length == 0) return inputStream.overrun ? Truncated : BadCode;
                if (inputStream.overrun) return Truncated;
                if (c.symbol < 256) {
                    if (pos >= cap) return NeedMoreOutput;
                    dest[pos++] = (uint8_t)c.symbol;
                } else if (c.symbol == 256) break;
                else {
                    const int ls = c.symbol - 257;
                    if (ls >= 29) return BadCode;
                    const size_t len = kLenBase[ls] + inputStream.Read(kLenExtra[ls]);
                    const code d = ReadCode(lengthCodes);
                    if (d.length == 0 || d.symbol >= 30) return BadCode;
                    const size_t dist = kDistBase[d.symbol] + inputStream.Read(kDistExtra[d.symbol]);
                    if (inputStream.overrun) return Truncated;
                    if (dist > (pos - start) + dictLen) return BadDistance;
                    if (pos + len > cap) return NeedMoreOutput;
                    for (size_t k = 0; k < len; ++k, ++pos) {
                        const size_t back = pos - start;                // bytes this call produced
                        dest[pos] = dist <= back ? dest[pos - dist] : dict[dictLen - (dist - back)];
                    }
                }
            }
        } else return BadBlock;
        if (final) return Ok;
    }
}

Decoder::Status Decoder::Inflate(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, Format format,
                                 const uint8_t* dict, size_t dictLen)
{
    const size_t cap = *destLen;
    size_t pos = 0;
    *destLen = 0;
    inputStream.Reset(source, sourceLen);
    bool firstMember = true;
    for (;;) {
        const size_t memberStart = pos;
        if (format == Zlib) {
            uint8_t h[2];
            if (!inputStream.ReadBytes(h, 2)) return Truncated;
            if ((h[0] & 0x0F) != 8 || ((h[0] << 8) | h[1]) % 31 != 0 || (h[1] & 0x20)) return BadHeader;
        } else if (format == Gzip) {
            uint8_t h[10];
            if (!inputStream.ReadBytes(h, 10)) return firstMember ? Truncated : Ok;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8) return BadHeader;
            const uint8_t flg = h[3];
            if (flg & 4) { uint8_t x[2]; if (!inputStream.ReadBytes(x, 2)) return Truncated; size_t n = x[0] | (x[1] << 8); uint8_t b; while (n--) if (!inputStream.ReadBytes(&b, 1)) return Truncated; }
            for (int bit = 8; bit <= 16; bit <<= 1)
                if (flg & bit) { uint8_t b; do { if (!inputStream.ReadBytes(&b, 1)) return Truncated; } while (b); }
            if (flg & 2) { uint8_t x[2]; if (!inputStream.ReadBytes(x, 2)) return Truncated; }
        }
        const Status st = InflateBlocks(dest, cap, pos, firstMember ? dict : nullptr, firstMember ? dictLen : 0);
        if (st != Ok) return st;
        inputStream.AlignToByte();
        if (format == Zlib) {
            uint8_t t[4];
            if (!inputStream.ReadBytes(t, 4)) return Truncated;
            const uint32_t want = ((uint32_t)t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
            if (adler32_host(1, dest + memberStart, pos - memberStart) != want) return BadChecksum;
        } else if (format == Gzip) {
            uint8_t t[8];
            if (!inputStream.ReadBytes(t, 8)) return Truncated;
            const uint32_t crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
            const uint32_t isize = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
            if (crc32_host(0, dest + memberStart, pos - memberStart) != crc || (uint32_t)(pos - memberStart) != isize) return BadChecksum;
        }
        firstMember = false;
        if (format != Gzip || inputStream.Consumed(source) >= sourceLen) break;     // gzip: further members may follow
    }
    *destLen = pos;
    return Ok;
}

void ZzFlateDecode(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, Format format)
{
    Decoder d;
    size_t n = *destLen;
    if (d.Inflate(dest, &n, source, sourceLen, format) != Decoder::Ok) { *destLen = ~(size_t)0; return; }
    *destLen = n;
}

extern "C" {
/* plain-C wrapper: returns bytes produced, or (size_t)-1 - status on error (status < 0: Decoder::Status) */
ZZGPU_API size_t zz_c_decode(uint8_t* dest, size_t cap, const uint8_t* src, size_t n, int format, const uint8_t* dict, size_t dictLen, int* status)
{
    Decoder d;
    size_t len = cap;
    const Decoder::Status st = d.Inflate(dest, &len, src, n, (Format)format, dict, dictLen);
    if (status) *status = (int)st;
    return st == Decoder::Ok ? len : ~(size_t)0;
}
}
