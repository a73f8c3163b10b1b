// Host driver: the reference's public C++ API (include/zzflate.h, encoder.h, crc.h) on top of the C-ABI
// of include/zzgpu.h.  Replaces zzflate/zzflate.cpp (framing, partitioning, stitch, trailer) -- the
// compression itself happens in the kernels behind zzgpu_deflate_ex.
#include "../../include/zzflate.h"
#include "../../include/encoder.h"
#include "../../include/crc.h"
#include "../../include/zzgpu.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

namespace {

const size_t kCallbackPiece = 1000000;          // outputbitstream.h:183

// zlib: CM=8, CINFO=7, FCHECK so that the 16-bit header is a multiple of 31 (zzflate.cpp:30-36) -> 78 01
// gzip: fixed 10 bytes, no mtime, XFL=0, OS=255 (zzflate.cpp:28)
size_t headerBytes(Format f, uint8_t* h)
{
    switch (f) {
    case Zlib: {
        const unsigned cmf = 8u | (7u << 4);
        unsigned flg = 0;
        flg |= 31u - ((cmf * 256u + flg) % 31u);
        h[0] = (uint8_t)cmf; h[1] = (uint8_t)flg;
        return 2;
    }
    case Gzip: {
        static const uint8_t gz[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xFF };
        memcpy(h, gz, 10);
        return 10;
    }
    default:
        return 0;
    }
}

// Adler-32 big-endian / CRC-32 + ISIZE little-endian (zzflate.cpp:170-192)
size_t trailerBytes(Format f, uint32_t adler, uint32_t crc, size_t n, uint8_t* t)
{
    if (f == Zlib) {
        t[0] = (uint8_t)(adler >> 24); t[1] = (uint8_t)(adler >> 16); t[2] = (uint8_t)(adler >> 8); t[3] = (uint8_t)adler;
        return 4;
    }
    if (f == Gzip) {
        const uint32_t isize = (uint32_t)n;
        for (int i = 0; i < 4; ++i) { t[i] = (uint8_t)(crc >> (8 * i)); t[4 + i] = (uint8_t)(isize >> (8 * i)); }
        return 8;
    }
    return 0;
}

struct Shard {
    size_t off = 0, len = 0;
    int device = 0;
    bool final = false;
    std::vector<uint8_t> out;
    size_t outLen = 0;
    uint32_t adler0 = 0, crc = 0;
    int status = ZZGPU_OK;
};

// Raw deflate of source[0,n) into dest (capacity cap).  threaded => one contiguous range of chunks per
// visible GPU, each primed with the bytes before it, stitched in order (zzflate.cpp:97-154).
int deflateStream(uint8_t* dest, size_t cap, const uint8_t* source, size_t n, int level, bool threaded,
                  Format format, size_t* outLen, uint32_t* adler, uint32_t* crc)
{
    const int want = (format == Zlib ? 1 : 0) | (format == Gzip ? 2 : 0);
    const uint32_t chunk = ZZGPU_DEFAULT_CHUNK, dict = ZZGPU_DEFAULT_DICT;
    const size_t nchunks = (n + chunk - 1) / chunk;
    int ndev = threaded ? zzgpu_device_count() : 1;
    if ((size_t)ndev > nchunks) ndev = (int)std::max<size_t>(nchunks, 1);
    if (ndev <= 1) {
        uint32_t a0 = 0, c = 0;
        int rc = zzgpu_deflate_ex(source, n, 0, 1, ZZGPU_MEM_HOST, dest, cap, ZZGPU_MEM_HOST, level, chunk, dict, want,
                                  outLen, &a0, &c, nullptr);
        if (rc) return rc;
        *adler = zzgpu_adler32_combine(1, a0, n);
        *crc = c;
        return ZZGPU_OK;
    }
    std::vector<Shard> shards((size_t)ndev);
    const size_t per = (nchunks + ndev - 1) / ndev;
    for (int g = 0; g < ndev; ++g) {
        Shard& s = shards[(size_t)g];
        const size_t c0 = std::min(nchunks, per * g), c1 = std::min(nchunks, per * (g + 1));
        s.off = std::min(n, c0 * chunk); s.len = std::min(n, c1 * chunk) - s.off;
        s.device = g; s.final = (c1 == nchunks);
    }
    std::vector<std::thread> threads;
    for (auto& s : shards) {
        threads.emplace_back([&s, source, level, want]() {
            if (s.len == 0 && !s.final) return;
            s.status = zzgpu_init(s.device);
            if (s.status) return;
            s.out.resize(zzgpu_bound(s.len, level, ZZGPU_DEFAULT_CHUNK));
            s.status = zzgpu_deflate_ex(source + s.off, s.len, s.off, s.final ? 1 : 0, ZZGPU_MEM_HOST,
                                        s.out.data(), s.out.size(), ZZGPU_MEM_HOST, level,
                                        ZZGPU_DEFAULT_CHUNK, ZZGPU_DEFAULT_DICT, want, &s.outLen, &s.adler0, &s.crc, nullptr);
        });
    }
    for (auto& t : threads) t.join();
    size_t pos = 0;
    uint32_t a = 1, c = 0;
    for (auto& s : shards) {
        if (s.status) return s.status;
        if (pos + s.outLen > cap) return ZZGPU_E_CAPACITY;
        memcpy(dest + pos, s.out.data(), s.outLen);
        pos += s.outLen;
        a = zzgpu_adler32_combine(a, s.adler0, s.len);
        c = zzgpu_crc32_combine(c, s.crc, s.len);
    }
    *outLen = pos; *adler = a; *crc = c;
    return ZZGPU_OK;
}

int distanceSymbol(int offset)
{
    if (offset < 1 || offset > 32768) return -1;
    const int t = offset - 1;
    if (t < 4) return t;
    int nb = 31 - __builtin_clz((unsigned)t);
    return 2 * nb + ((t >> (nb - 1)) & 1);
}

}  // namespace

void ZzFlateEncode(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, const Config* config)
{
    uint8_t h[10], t[8];
    const size_t hl = headerBytes(config->format, h);
    const size_t tl = config->format == Zlib ? 4 : config->format == Gzip ? 8 : 0;
    if (config->level > 3 || *destLen < hl + tl) { *destLen = ~(size_t)0; return; }     // zzflate.cpp:230-234
    memcpy(dest, h, hl);
    size_t body = 0; uint32_t adler = 1, crc = 0;
    const int rc = deflateStream(dest + hl, *destLen - hl - tl, source, sourceLen, config->level, config->threaded,
                                 config->format, &body, &adler, &crc);
    if (rc != ZZGPU_OK) { *destLen = ~(size_t)0; return; }
    trailerBytes(config->format, adler, crc, sourceLen, t);
    memcpy(dest + hl + body, t, tl);
    *destLen = hl + body + tl;
}

void ZzFlateEncodeToCallback(const uint8_t* source, size_t sourceLen, const Config* config,
                             std::function<bool(const uint8_t*, size_t)> callback)
{
    if (config->level > 3) return;                                                     // zzflate.cpp:201-202
    std::vector<uint8_t> buf(zzgpu_bound(sourceLen, config->level, ZZGPU_DEFAULT_CHUNK));
    size_t body = 0; uint32_t adler = 1, crc = 0;
    const int rc = deflateStream(buf.data(), buf.size(), source, sourceLen, config->level, config->threaded,
                                 config->format, &body, &adler, &crc);
    if (rc != ZZGPU_OK) return;
    uint8_t h[10], t[8];
    const size_t hl = headerBytes(config->format, h);
    callback(h, hl);
    for (size_t pos = 0; pos < body; pos += kCallbackPiece)
        callback(buf.data() + pos, std::min(kCallbackPiece, body - pos));
    const size_t tl = trailerBytes(config->format, adler, crc, sourceLen, t);
    callback(t, tl);
}

uint32_t adler32x(uint32_t startValue, const uint8_t* data, size_t len)
{
    uint32_t a = startValue;
    if (zzgpu_checksums(data, len, ZZGPU_MEM_HOST, startValue, 0, &a, nullptr) != ZZGPU_OK) return ~0u;
    return a;
}

uint32_t combine(uint32_t first, uint32_t second, size_t lenSecond) { return zzgpu_adler32_combine(first, second, lenSecond); }

uint32_t crc32(const uint8_t* buffer, size_t length, uint32_t startValue)
{
    uint32_t c = startValue;
    if (zzgpu_checksums(buffer, length, ZZGPU_MEM_HOST, 1, startValue, nullptr, &c) != ZZGPU_OK) return ~0u;
    return c;
}

Encoder::Encoder(int level_, uint8_t* outputBuffer, int64_t bytes) : level(level_)
{
    stream.start = outputBuffer;
    stream.capacity = outputBuffer ? (size_t)bytes : 0;
}

bool Encoder::AddData(const uint8_t* start, const uint8_t* end, bool final)
{
    if (level < 0 || level > 3) return false;
    const size_t n = (size_t)(end - start);
    if (n == 0 && !final) return true;
    const size_t history = (start == lastEnd) ? contiguous : 0;
    const size_t need = zzgpu_bound(n, level, ZZGPU_DEFAULT_CHUNK);
    uint8_t* out; size_t cap;
    if (stream.start && stream.owned.empty()) {
        out = stream.start + stream.written; cap = stream.capacity - stream.written;
    } else {
        stream.owned.resize(stream.written + need);
        stream.start = stream.owned.data();
        out = stream.start + stream.written; cap = need;
    }
    size_t outLen = 0;
    const int rc = zzgpu_deflate_ex(start, n, history, final ? 1 : 0, ZZGPU_MEM_HOST, out, cap, ZZGPU_MEM_HOST, level,
                                    ZZGPU_DEFAULT_CHUNK, ZZGPU_DEFAULT_DICT, 0, &outLen, nullptr, nullptr, nullptr);
    if (rc != ZZGPU_OK) return false;
    stream.written += outLen;
    contiguous = history + n;
    lastEnd = end;
    return true;
}

bool Encoder::AddDataGzip(const uint8_t* start, const uint8_t* end, uint32_t& adler, bool final)
{
    if (!AddData(start, end, final)) return false;
    adler = adler32x(adler, start, (size_t)(end - start));
    return true;
}

int Encoder::FindDistance(int offset) { return distanceSymbol(offset); }
int Encoder::ReadLut(int offset) { return offset == 0 ? 255 : distanceSymbol(offset); }     // luts.cpp:116: index 0 holds 255

void Encoder::CreateMergedLengthCodes(code* lCodes, code* symbolCodes)
{
    static const int base[29] = { 3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258 };
    static const int ext[29] = { 0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0 };
    for (int len = 0; len < 259; ++len) {
        int sym = 0, eb = 0, ev = 0;                      // lengths 0..2 map to symbol 0 with no extra bits (luts.cpp:6)
        if (len >= 3) {
            int s = 28;
            if (len < 258) { s = 0; while (s < 27 && base[s + 1] <= len) ++s; }
            sym = 257 + s; eb = ext[s]; ev = len - base[s];
        }
        lCodes[len].length = symbolCodes[sym].length + eb;
        lCodes[len].bits = ((uint32_t)ev << symbolCodes[sym].length) | symbolCodes[sym].bits;
    }
}

// ---- plain-C wrappers so that tests (ctypes) and other languages can drive the C++ API above ----
extern "C" {

ZZGPU_API size_t zz_c_encode(uint8_t* dest, size_t cap, const uint8_t* src, size_t n, int format, int level, int threaded)
{
    Config cfg = { (Format)format, (uint8_t)level, threaded != 0 };
    size_t len = cap;
    ZzFlateEncode(dest, &len, src, n, &cfg);
    return len;
}

typedef int (*zz_c_callback)(const uint8_t*, size_t, void*);

ZZGPU_API void zz_c_encode_to_callback(const uint8_t* src, size_t n, int format, int level, int threaded,
                                       zz_c_callback cb, void* user)
{
    Config cfg = { (Format)format, (uint8_t)level, threaded != 0 };
    ZzFlateEncodeToCallback(src, n, &cfg, [cb, user](const uint8_t* b, size_t c) -> bool { return cb(b, c, user) != 0; });
}

ZZGPU_API uint32_t zz_c_adler32x(uint32_t start, const uint8_t* d, size_t n) { return adler32x(start, d, n); }
ZZGPU_API uint32_t zz_c_combine(uint32_t a, uint32_t b, size_t lenB) { return combine(a, b, lenB); }
ZZGPU_API uint32_t zz_c_crc32(const uint8_t* d, size_t n, uint32_t start) { return crc32(d, n, start); }
ZZGPU_API int zz_c_find_distance(int offset) { return Encoder::FindDistance(offset); }
ZZGPU_API int zz_c_read_lut(int offset) { return Encoder::ReadLut(offset); }

ZZGPU_API void zz_c_merged_length_codes(const int32_t* symbolCodes286x2, int32_t* lcodes259x2)
{
    code sym[286], l[259];
    for (int i = 0; i < 286; ++i) { sym[i].length = symbolCodes286x2[2 * i]; sym[i].bits = (uint32_t)symbolCodes286x2[2 * i + 1]; }
    Encoder::CreateMergedLengthCodes(l, sym);
    for (int i = 0; i < 259; ++i) { lcodes259x2[2 * i] = l[i].length; lcodes259x2[2 * i + 1] = (int32_t)l[i].bits; }
}

ZZGPU_API void* zz_c_encoder_new(int level, uint8_t* out, int64_t cap) { return new Encoder(level, out, cap); }
ZZGPU_API void zz_c_encoder_free(void* e) { delete (Encoder*)e; }
ZZGPU_API int zz_c_encoder_add_data(void* e, const uint8_t* start, size_t n, int final) { return ((Encoder*)e)->AddData(start, start + n, final != 0) ? 1 : 0; }
ZZGPU_API void zz_c_encoder_set_level(void* e, int level) { ((Encoder*)e)->SetLevel(level); }
ZZGPU_API size_t zz_c_encoder_bytes(void* e) { ((Encoder*)e)->stream.Flush(); return ((Encoder*)e)->stream.byteswritten(); }
ZZGPU_API const uint8_t* zz_c_encoder_data(void* e) { return ((Encoder*)e)->stream.streamStart(); }

}  // extern "C"
