/* encoder.h -- the part of the reference's zzflate/encoder.h that callers and tests use directly
 * (encoder.h:10-12,79-97): the free checksum functions, and an Encoder object that can be fed
 * incrementally.  Private state of the reference class (hash table, records, code tables) lives on the
 * device and is not mirrored.
 */
#ifndef ZZFLATE_B200_ENCODER_H
#define ZZFLATE_B200_ENCODER_H

#include <stddef.h>
#include <stdint.h>
#include <vector>

/* Adler-32 continuing from startValue (adler.cpp:17-43; true Adler-32, i.e. without the reference's
 * overflow above ~362 MiB). */
uint32_t adler32x(uint32_t startValue, const uint8_t* data, size_t len);

/* Adler-32 of A||B from adler(A) and adler(B computed with start value 0) (adler.cpp:5-15). */
uint32_t combine(uint32_t first, uint32_t second, size_t lenSecond);

struct code          /* outputbitstream.h:14-24 : LSB-first bit string */
{
    int32_t length;
    uint32_t bits;
};

/* Output side of an Encoder: what the reference exposes as `outputbitstream stream` (Flush,
 * byteswritten, streamStart).  Chunks always end byte-aligned here, so Flush has nothing left to pad. */
struct outputbytestream
{
    void Flush() {}
    size_t byteswritten() const { return written; }
    uint8_t* streamStart() { return start; }

    uint8_t* start = nullptr;
    size_t capacity = 0;
    size_t written = 0;
    std::vector<uint8_t> owned;      /* used when the Encoder was built without an output buffer */
};

struct Encoder
{
    Encoder(int level, uint8_t* outputBuffer = nullptr, int64_t bytes = 0);      /* encoder.cpp:527 */

    /* Encodes [start,end) behind what was already written; `final` marks the last chunk BFINAL.  Data that
     * directly follows the previous call in memory is primed with it as dictionary (the reference keeps
     * its hash table across calls, encoder.cpp:248).  Returns false on error (encoder.cpp:539-551). */
    bool AddData(const uint8_t* start, const uint8_t* end, bool final);
    void SetLevel(int newlevel) { level = newlevel; }
    bool AddDataGzip(const uint8_t* start, const uint8_t* end, uint32_t& adler, bool final);   /* encoder.cpp:554 */

    static int FindDistance(int offset);                                  /* encoder.cpp:51-61 */
    static int ReadLut(int offset);                                       /* encoder.h:93 */
    static void CreateMergedLengthCodes(code* lCodes, code* symbolCodes); /* encoder.cpp:126-133 */

    outputbytestream stream;

private:
    int level;
    const uint8_t* lastEnd = nullptr;
    size_t contiguous = 0;           /* bytes before lastEnd that belong to this stream */
};

#endif
