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
#include "outputbitstream.h"      /* struct code, host bit writer */

/* Adler-32 continuing from startValue (adler.cpp:17-43; true Adler-32, i.e. without the reference's
 * overflow above ~362 MiB). */
uint32_t adler32x(uint32_t startValue, const uint8_t* data, size_t len);

/* Adler-32 of A||B from adler(A) and adler(B computed with start value 0) (adler.cpp:5-15). */
uint32_t combine(uint32_t first, uint32_t second, size_t lenSecond);

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

    /* Encodes [start,end) behind what was already written; `final` marks the last chunk BFINAL.  Successive calls
     * continue one stream: each call is primed with the last 32 KiB (+288 bytes) the Encoder was fed, of which it
     * keeps a private copy -- the caller's earlier buffers may be freed or reused (the reference keeps its hash table
     * across calls, encoder.cpp:248,320-327, and silently requires the earlier bytes to stay in place in front of
     * `start`).  Every call ends byte-aligned (E-mode, zzgpu.h), which the reference does not do between calls.
     * Returns false on error (encoder.cpp:539-551). */
    bool AddData(const uint8_t* start, const uint8_t* end, bool final);
    void SetLevel(int newlevel) { level = newlevel; }
    bool AddDataGzip(const uint8_t* start, const uint8_t* end, uint32_t& adler, bool final);   /* encoder.cpp:554 */

    static int FindDistance(int offset);                                  /* encoder.cpp:51-61 */
    static int ReadLut(int offset);                                       /* encoder.h:93 */
    static void CreateMergedLengthCodes(code* lCodes, code* symbolCodes); /* encoder.cpp:126-133 */

    outputbytestream stream;

private:
    int level;
    std::vector<uint8_t> tail;       /* the last <= 32 KiB + 288 bytes of the stream so far */
};

#endif
