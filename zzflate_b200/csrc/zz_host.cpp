// Host driver: the reference's public C++ API (include/zzflate.h, encoder.h, crc.h) on top of the C-ABI
// of include/zzgpu.h.  Replaces zzflate/zzflate.cpp (framing, partitioning, stitch, trailer) -- the
// compression itself happens in the kernels behind zzgpu_deflate_ex.
#include "../../include/zzflate.h"
#include "../../include/encoder.h"
#include "../../include/crc.h"
#include "../../include/zzgpu.h"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
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
    size_t outLen = 0;
    uint32_t adler0 = 0, crc = 0;
    int status = ZZGPU_OK;
};

// Contiguous ranges of whole chunks, one per device that gets any (zzflate.cpp:67-78 divideInRanges at chunk
// granularity).  Only shards with data exist, so exactly the last one carries BFINAL.
std::vector<Shard> partition(size_t n, int ndevAvail, uint32_t chunk)
{
    const size_t nchunks = (n + chunk - 1) / chunk;
    size_t ndev = (size_t)std::max(ndevAvail, 1);
    if (ndev > nchunks) ndev = std::max<size_t>(nchunks, 1);
    const size_t per = std::max<size_t>((nchunks + ndev - 1) / ndev, 1);
    ndev = std::max<size_t>((nchunks + per - 1) / per, 1);          // devices that actually receive chunks
    std::vector<Shard> shards(ndev);
    for (size_t g = 0; g < ndev; ++g) {
        Shard& s = shards[g];
        const size_t c0 = std::min(nchunks, per * g), c1 = std::min(nchunks, per * (g + 1));
        s.off = std::min(n, c0 * chunk); s.len = std::min(n, c1 * chunk) - s.off;
        s.device = (int)g; s.final = (g + 1 == ndev);
    }
    return shards;
}

struct SinkCtx { const std::function<bool(const uint8_t*, size_t)>* cb; };

int sinkTrampoline(const uint8_t* p, size_t len, void* user)
{
    (*static_cast<SinkCtx*>(user)->cb)(p, len);
    return 0;
}

// Raw deflate of source[0,n).  Output goes either to dest (capacity cap) or, when `cb` is given, to the callback in
// order, in pieces of at most kCallbackPiece bytes.  threaded => one contiguous range of chunks per visible GPU, each
// primed with the bytes before it (zzflate.cpp:97-132).  Shard 0 streams straight to its destination while it is
// encoded; the other shards stay in their GPU's memory until the sizes of the shards before them are known and are
// then copied to their final place (no temporary, no host-side memmove as in zzflate.cpp:136-154).
int deflateStream(uint8_t* dest, size_t cap, const std::function<bool(const uint8_t*, size_t)>* cb,
                  const uint8_t* source, size_t n, int level, bool threaded,
                  Format format, size_t* outLen, uint32_t* adler, uint32_t* crc)
{
    const int want = (format == Zlib ? 1 : 0) | (format == Gzip ? 2 : 0);
    const uint32_t chunk = ZZGPU_DEFAULT_CHUNK, dict = ZZGPU_DEFAULT_DICT;
    SinkCtx sctx = { cb };
    std::vector<Shard> shards = partition(n, threaded ? zzgpu_device_count() : 1, chunk);
    if (shards.size() <= 1) {
        uint32_t a0 = 0, c = 0;
        int rc;
        if (cb) rc = zzgpu_deflate_sink(source, n, 0, 1, level, chunk, dict, want, sinkTrampoline, &sctx, kCallbackPiece, outLen, &a0, &c, nullptr);
        else rc = zzgpu_deflate_ex(source, n, 0, 1, ZZGPU_MEM_HOST, dest, cap, ZZGPU_MEM_HOST, level, chunk, dict, want,
                                   outLen, &a0, &c, nullptr);
        if (rc) return rc;
        *adler = zzgpu_adler32_combine(1, a0, n);
        *crc = c;
        return ZZGPU_OK;
    }
    // one host thread per device; each keeps its device context until its stream has been fetched
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> sized(shards.size(), 0);          // shard g's size is known (or it failed)
    std::vector<size_t> place(shards.size(), 0);        // final offset of shard g
    bool abortAll = false;
    size_t emitted = 0;                                  // callback mode: shards handed over so far
    std::vector<std::thread> threads;
    for (size_t g = 0; g < shards.size(); ++g) {
        threads.emplace_back([&, g]() {
            Shard& s = shards[g];
            auto publish = [&](bool ok) {
                std::lock_guard<std::mutex> lk(mu);
                sized[g] = 1; if (!ok) abortAll = true;
                cv.notify_all();
            };
            s.status = zzgpu_init(s.device);
            if (s.status) { publish(false); return; }
            if (g == 0) {
                // the first shard's place is known: it goes out while it is being encoded
                if (cb) s.status = zzgpu_deflate_sink(source, s.len, 0, 0, level, chunk, dict, want, sinkTrampoline, &sctx, kCallbackPiece,
                                                      &s.outLen, &s.adler0, &s.crc, nullptr);
                else s.status = zzgpu_deflate_ex(source, s.len, 0, 0, ZZGPU_MEM_HOST, dest, cap, ZZGPU_MEM_HOST, level, chunk, dict, want,
                                                 &s.outLen, &s.adler0, &s.crc, nullptr);
                { std::lock_guard<std::mutex> lk(mu); emitted = 1; }
                publish(s.status == ZZGPU_OK);
                return;
            }
            s.status = zzgpu_deflate_hold(source + s.off, s.len, s.off, s.final ? 1 : 0, level, chunk, dict, want,
                                          &s.outLen, &s.adler0, &s.crc, nullptr);
            publish(s.status == ZZGPU_OK);
            if (s.status) return;
            size_t at = 0;
            {
                std::unique_lock<std::mutex> lk(mu);
                // direct mode: the sizes of all earlier shards; callback mode: all earlier shards handed over
                cv.wait(lk, [&] {
                    if (abortAll) return true;
                    if (cb) return emitted == g;
                    for (size_t k = 0; k < g; ++k) if (!sized[k]) return false;
                    return true;
                });
                if (abortAll) { lk.unlock(); zzgpu_release(); return; }
                for (size_t k = 0; k < g; ++k) at += shards[k].outLen;
                place[g] = at;
            }
            if (cb) {
                s.status = zzgpu_fetch(nullptr, 0, sinkTrampoline, &sctx, kCallbackPiece);
                std::lock_guard<std::mutex> lk(mu);
                emitted = g + 1; if (s.status) abortAll = true;
                cv.notify_all();
            } else if (at + s.outLen > cap) {
                zzgpu_release();
                s.status = ZZGPU_E_CAPACITY;
            } else {
                s.status = zzgpu_fetch(dest + at, cap - at, nullptr, nullptr, 0);
            }
        });
    }
    for (auto& t : threads) t.join();
    size_t pos = 0;
    uint32_t a = 1, c = 0;
    for (auto& s : shards) {
        if (s.status) return s.status;
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


// A gzip member records its input length modulo 2^32 (ISIZE, RFC 1952 2.3.1), so inputs of 4 GiB and more are written
// as several members (RFC 1952 2.2: "a gzip file consists of a series of members"): every member is a complete
// deflate stream with its own header, CRC-32 and ISIZE, and inflaters concatenate them.  (The reference truncates ISIZE,
// zzflate.cpp:183-186.)  Zlib and raw deflate have no such field and stay one stream.
size_t g_gzipMemberBytes = ((size_t)0xFFFFFFFFu / ZZGPU_DEFAULT_CHUNK) * ZZGPU_DEFAULT_CHUNK;      // 4 GiB - 64 KiB

size_t memberLimit(Format f, size_t n)
{
    return (f == Gzip && n > g_gzipMemberBytes) ? g_gzipMemberBytes : n;
}

}  // namespace

void ZzFlateEncode(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, const Config* config)
{
    uint8_t h[10], t[8];
    const size_t hl = headerBytes(config->format, h);
    const size_t tl = config->format == Zlib ? 4 : config->format == Gzip ? 8 : 0;
    if (config->level > 3 || *destLen < hl + tl) { *destLen = ~(size_t)0; return; }     // zzflate.cpp:230-234
    const size_t cap = *destLen;
    const size_t member = memberLimit(config->format, sourceLen);
    size_t pos = 0, off = 0;
    do {
        const size_t len = std::min(member, sourceLen - off);
        if (cap - pos < hl + tl) { *destLen = ~(size_t)0; return; }
        memcpy(dest + pos, h, hl);
        size_t body = 0; uint32_t adler = 1, crc = 0;
        const int rc = deflateStream(dest + pos + hl, cap - pos - hl - tl, nullptr, source + off, len, config->level, config->threaded,
                                     config->format, &body, &adler, &crc);
        if (rc != ZZGPU_OK) { *destLen = ~(size_t)0; return; }
        trailerBytes(config->format, adler, crc, len, t);
        memcpy(dest + pos + hl + body, t, tl);
        pos += hl + body + tl;
        off += len;
    } while (off < sourceLen);
    *destLen = pos;
}

// Header, the stream in pieces of <= 1 000 000 bytes as they leave the GPU (the first piece arrives while later
// input is still being copied in and encoded), trailer.  Nothing of the size of the whole stream is ever allocated.
void ZzFlateEncodeToCallback(const uint8_t* source, size_t sourceLen, const Config* config,
                             std::function<bool(const uint8_t*, size_t)> callback)
{
    if (config->level > 3) return;                                                     // zzflate.cpp:201-202
    if (zzgpu_device_count() <= 0) return;                                             // no CPU fallback: nothing is delivered
    uint8_t h[10], t[8];
    const size_t hl = headerBytes(config->format, h);
    const size_t member = memberLimit(config->format, sourceLen);
    size_t off = 0;
    do {
        const size_t len = std::min(member, sourceLen - off);
        callback(h, hl);
        size_t body = 0; uint32_t adler = 1, crc = 0;
        const int rc = deflateStream(nullptr, 0, &callback, source + off, len, config->level, config->threaded,
                                     config->format, &body, &adler, &crc);
        if (rc != ZZGPU_OK) return;                   // a failed stream gets no trailer, so no inflater accepts it
        const size_t tl = trailerBytes(config->format, adler, crc, len, t);
        callback(t, tl);
        off += len;
    } while (off < sourceLen);
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
    const size_t need = zzgpu_bound(n, level, ZZGPU_DEFAULT_CHUNK);
    uint8_t* out; size_t cap;
    if (stream.start && stream.owned.empty()) {
        out = stream.start + stream.written; cap = stream.capacity - stream.written;
    } else {
        stream.owned.resize(stream.written + need);
        stream.start = stream.owned.data();
        out = stream.start + stream.written; cap = need;
    }
    // The dictionary of this call is the private copy of what the previous calls were fed (the reference keeps its hash
    // table across calls, encoder.cpp:248,320-327): the caller's earlier buffers are never read again.
    size_t outLen = 0;
    const int rc = zzgpu_deflate_hist(start, n, tail.data(), tail.size(), final ? 1 : 0, out, cap, level,
                                      ZZGPU_DEFAULT_CHUNK, ZZGPU_DEFAULT_DICT, &outLen);
    if (rc != ZZGPU_OK) return false;
    stream.written += outLen;
    const size_t keep = (size_t)ZZGPU_DEFAULT_DICT + 288;
    if (n >= keep) tail.assign(end - keep, end);
    else {
        const size_t old = std::min(tail.size(), keep - n);
        tail.erase(tail.begin(), tail.end() - (std::ptrdiff_t)old);
        tail.insert(tail.end(), start, end);
    }
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

// shard partition of the multi-GPU driver: writes (offset, length, final) triples, returns the number of shards
ZZGPU_API int zz_c_partition(size_t n, int ndev, uint32_t chunk, uint64_t* triples, int maxShards)
{
    const std::vector<Shard> sh = partition(n, ndev, chunk ? chunk : ZZGPU_DEFAULT_CHUNK);
    for (size_t g = 0; g < sh.size() && (int)g < maxShards; ++g) { triples[3 * g] = sh[g].off; triples[3 * g + 1] = sh[g].len; triples[3 * g + 2] = sh[g].final ? 1 : 0; }
    return (int)sh.size();
}

// tests: gzip member size (default 4 GiB - 64 KiB; 0 restores the default)
ZZGPU_API void zz_c_set_gzip_member_bytes(size_t bytes)
{
    g_gzipMemberBytes = bytes ? (bytes / ZZGPU_DEFAULT_CHUNK) * ZZGPU_DEFAULT_CHUNK : ((size_t)0xFFFFFFFFu / ZZGPU_DEFAULT_CHUNK) * ZZGPU_DEFAULT_CHUNK;
    if (g_gzipMemberBytes == 0) g_gzipMemberBytes = ZZGPU_DEFAULT_CHUNK;
}

ZZGPU_API void* zz_c_encoder_new(int level, uint8_t* out, int64_t cap) { return new Encoder(level, out, cap); }
ZZGPU_API void zz_c_encoder_free(void* e) { delete (Encoder*)e; }
ZZGPU_API int zz_c_encoder_add_data(void* e, const uint8_t* start, size_t n, int final) { return ((Encoder*)e)->AddData(start, start + n, final != 0) ? 1 : 0; }
ZZGPU_API void zz_c_encoder_set_level(void* e, int level) { ((Encoder*)e)->SetLevel(level); }
ZZGPU_API size_t zz_c_encoder_bytes(void* e) { ((Encoder*)e)->stream.Flush(); return ((Encoder*)e)->stream.byteswritten(); }
ZZGPU_API const uint8_t* zz_c_encoder_data(void* e) { return ((Encoder*)e)->stream.streamStart(); }

}  // extern "C"
