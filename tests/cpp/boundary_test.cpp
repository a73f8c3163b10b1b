// Boundary proof: a C++14 caller written against include/zzflate.h, encoder.h, crc.h, outputbitstream.h, huffman.h,
// linked to libzzflate_b200.so.  Re-hosts the reference's own tests with plain checks (gtest is not installed):
//   testroundtrip / testroundtripgzip       zztest/Test.cpp:202-282   (zlib inflate is the judge, Test.cpp:87-140)
//   Adler.Combine                            zztest/Test.cpp:301-313
//   ZzFlate.SmallZerBouffer                  zztest/Test.cpp:330-338
//   ZzFlate.TestDistanceSearch               zztest/TestHuffman.cpp:9-31
//   ZzFlate.GenerateHuffman                  zztest/TestHuffman.cpp:34-50
//   BitOutput.TestSimple/TestSimple2/TrivHuffman   zztest/TestBitOutput.cpp:7-48
// usage: boundary_test <input file>      (exit code 0 = all passed; needs a CUDA device: there is no CPU fallback)
#include "../../include/zzflate.h"
#include "../../include/encoder.h"
#include "../../include/crc.h"
#include "../../include/huffman.h"
#include "../../include/outputbitstream.h"
#include "../../include/decoder.h"

#include <zlib.h>
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static int ZlibUncompress(uint8_t* dest, size_t* destLen, const uint8_t* source, size_t sourceLen, bool gzip)
{
    z_stream s{};
    if (inflateInit2(&s, gzip ? 15 + 16 : 15) != Z_OK) return Z_MEM_ERROR;
    s.next_in = const_cast<Bytef*>(source); s.next_out = dest;
    size_t inLeft = sourceLen, outLeft = *destLen;
    int err = Z_OK;
    while (err == Z_OK) {
        s.avail_in = (uInt)std::min<size_t>(inLeft, 1u << 30); s.avail_out = (uInt)std::min<size_t>(outLeft, 1u << 30);
        const uInt in0 = s.avail_in, out0 = s.avail_out;
        err = inflate(&s, Z_NO_FLUSH);
        inLeft -= in0 - s.avail_in; outLeft -= out0 - s.avail_out;
        if (in0 == s.avail_in && out0 == s.avail_out) break;
    }
    *destLen -= outLeft;
    inflateEnd(&s);
    return err == Z_STREAM_END ? Z_OK : (err == Z_OK ? Z_BUF_ERROR : err);
}

static long testroundtrip(const std::vector<uint8_t>& in, Config config)
{
    std::vector<uint8_t> compressed;
    if (config.level == 1) {
        compressed.resize(std::max<size_t>(200, in.size() + in.size() / 8 + 64));      // level 1 may expand (9-bit literals)
        size_t len = compressed.size();
        ZzFlateEncode(compressed.data(), &len, in.data(), in.size(), &config);
        if (len == ~(size_t)0) return -1;
        compressed.resize(len);
    } else {
        ZzFlateEncodeToCallback(in.data(), in.size(), &config, [&compressed](const uint8_t* b, size_t n) -> bool {
            compressed.insert(compressed.end(), b, b + n);
            return false;
        });
    }
    std::vector<uint8_t> out(in.size() + 1);
    size_t outLen = out.size();
    if (ZlibUncompress(out.data(), &outLen, compressed.data(), compressed.size(), config.format == Gzip) != Z_OK) return -1;
    if (outLen != in.size() || !std::equal(in.begin(), in.end(), out.begin())) return -1;
    // and through the library's own inflater (include/decoder.h: the role zzflate/decoder.h was meant to have)
    std::vector<uint8_t> out2(in.size() + 1);
    size_t outLen2 = out2.size();
    ZzFlateDecode(out2.data(), &outLen2, compressed.data(), compressed.size(), config.format);
    if (outLen2 != in.size() || !std::equal(in.begin(), in.end(), out2.begin())) return -1;
    return (long)compressed.size();
}

int main(int argc, char** argv)
{
    std::vector<uint8_t> file;
    if (argc > 1) { std::ifstream f(argv[1], std::ios::binary); file.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>()); }
    if (file.empty()) { for (int i = 0; i < 300000; ++i) file.push_back((uint8_t)("the quick brown fox jumps over the lazy dog "[(i * 7 + i / 13) % 44])); }

    {   // Adler.Combine
        const std::vector<unsigned char> asdf = { 0, 1, 23, 30, 4, 69, 145, 32, 216 };
        const uint32_t total = adler32x(1, asdf.data(), 9);
        const int split = 5;
        CHECK(total == combine(adler32x(1, asdf.data(), split), adler32x(0, asdf.data() + split, 9 - split), 9 - split));
        CHECK(total == adler32(1, asdf.data(), 9));
        CHECK(crc32(asdf.data(), 9) == (uint32_t)::crc32(0, asdf.data(), 9));
        CHECK(crc32(asdf.data() + 4, 5, crc32(asdf.data(), 4)) == (uint32_t)::crc32(0, asdf.data(), 9));    // chaining via startValue
    }
    // ZzGzip.Simple, ZzFlate.UserHuffman (callback API), ZzFlate.FixedHuffman, levels 0 and 3, raw deflate via Encoder
    CHECK(testroundtrip(file, { Gzip, 1, false }) > 0);
    CHECK(testroundtrip(file, { Zlib, 2, false }) > 0);
    CHECK(testroundtrip(file, { Zlib, 1, false }) > 0);
    CHECK(testroundtrip(file, { Zlib, 3, true }) > 0);
    CHECK(testroundtrip(file, { Gzip, 0, false }) > 0);
    {   // error convention: level > 3 and a destination that cannot hold the header (zzflate.cpp:230-234)
        std::vector<uint8_t> dst(64); size_t len = dst.size();
        Config bad = { Zlib, 4, false };
        ZzFlateEncode(dst.data(), &len, file.data(), file.size(), &bad);
        CHECK(len == ~(size_t)0);
        Config gz = { Gzip, 2, false }; len = 5;
        ZzFlateEncode(dst.data(), &len, file.data(), file.size(), &gz);
        CHECK(len == ~(size_t)0);
    }
    for (int size = 1; size <= 512; size *= 2) {         // ZzFlate.SmallZerBouffer
        const std::vector<uint8_t> zeros((size_t)size, 0);
        CHECK(testroundtrip(zeros, { Zlib, 2, false }) > 0);
    }
    {   // Encoder fed incrementally, level switch in between (zzflate.cpp:116-120 uses exactly this)
        std::vector<uint8_t> out(file.size() + file.size() / 8 + 4096);
        Encoder enc(2, out.data(), (int64_t)out.size());
        const size_t cut = file.size() / 2;
        CHECK(enc.AddData(file.data(), file.data() + cut, false));
        enc.SetLevel(1);
        CHECK(enc.AddData(file.data() + cut, file.data() + file.size(), true));
        enc.stream.Flush();
        std::vector<uint8_t> back(file.size() + 1);
        z_stream s{};
        CHECK(inflateInit2(&s, -15) == Z_OK);
        s.next_in = enc.stream.streamStart(); s.avail_in = (uInt)enc.stream.byteswritten();
        s.next_out = back.data(); s.avail_out = (uInt)back.size();
        CHECK(inflate(&s, Z_FINISH) == Z_STREAM_END);
        CHECK(s.total_out == file.size() && std::equal(file.begin(), file.end(), back.begin()));
        inflateEnd(&s);
    }
    {   // ZzFlate.TestDistanceSearch
        int failCount = 0;
        for (int distance = 1; distance <= 32768; ++distance) failCount += Encoder::FindDistance(distance) != Encoder::ReadLut(distance);
        CHECK(failCount == 0);
    }
    {   // ZzFlate.GenerateHuffman
        const std::vector<int> lengths = huffman::defaultTableLengths();
        std::vector<code> result(lengths.size());
        huffman::generate<code>(lengths, result.data());
        CHECK(result[0].bits == huffman::reverse(0x30, lengths[0]));
        CHECK(result[143].bits == huffman::reverse(0xBF, lengths[143]));
        CHECK(result[144].bits == huffman::reverse(0x190, lengths[144]));
        CHECK(result[255].bits == huffman::reverse(0x1FF, lengths[255]));
        CHECK(result[256].bits == huffman::reverse(0, lengths[256]));
        CHECK(result[279].bits == huffman::reverse(0x17, lengths[279]));
        CHECK(result[280].bits == huffman::reverse(0xC0, lengths[280]));
        CHECK(result[287].bits == huffman::reverse(0xC7, lengths[287]));
        // the merged length codes the fixed-Huffman path uses are built from these (fixedhuffmanluts.cpp:8-46)
        std::vector<code> lcodes(259);
        Encoder::CreateMergedLengthCodes(lcodes.data(), result.data());
        CHECK(lcodes[3].length == 7 && lcodes[258].length == 8 && lcodes[11].length == 7 + 1);
    }
    {   // BitOutput.TestSimple / TestSimple2 / TrivHuffman
        std::vector<uint8_t> buffer(100);
        outputbitstream strm(buffer.data(), 100);
        strm.AppendToBitStream(1, 1);
        strm.AppendToBitStream(0, 2);
        CHECK(buffer[0] == 0);
        strm.Flush();
        CHECK(buffer[0] == 1);
        std::vector<uint8_t> buffer2(100);
        outputbitstream strm2(buffer2.data(), 100);
        strm2.AppendToBitStream(3, 2);
        strm2.AppendToBitStream(0, 2);
        strm2.AppendToBitStream(15, 4);
        strm2.Flush();
        CHECK(buffer2[0] == 0xF3);
        const std::vector<int> lens = { 2, 1, 3, 3 };
        std::vector<code> codes(lens.size());
        huffman::generate<code>(lens, codes.data());
        CHECK(codes[1].bits == 0);
    }
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "boundary ok", failures);
    return failures ? 1 : 0;
}
