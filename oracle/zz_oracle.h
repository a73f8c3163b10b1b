/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the zzflate deflate-encode hot path.
 *
 * Plain C restatement of the reference algorithm (jandevaan/zzflate), each function citing the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (zzflate_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement byte-for-byte against
 * the unmodified reference compiled into oracle/_ref/libzzref.so (whole streams, per-chunk E(c)
 * streams, token dumps, Huffman lengths, RLE records, static tables, checksums), and
 * tests/golden/ holds vectors generated from that library for the GPU box, where /root/reference
 * does not exist.
 *
 * Defect policies -- where the reference itself is wrong this restatement is
 * *correct* and reports the defect class in zzo_chunk_info.defects:
 *   R1 batch without a match loses a byte      -> the byte is counted
 *   R2 level-1 short match overruns block end  -> clamped to the block end
 *   R3 adler32x overflows above ~362 MiB       -> true Adler-32 (zzo_adler32x_literal keeps the overflow)
 *   R4 backward extension reads before input   -> clamped at global offset 0
 *   R6 backward extension >= 259               -> capped at 258
 *   R7 empty input emits no block              -> one final empty stored block
 */
#ifndef ZZ_ORACLE_H
#define ZZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ZZO_ZLIB = 0, ZZO_GZIP = 1, ZZO_DEFLATE = 2 };          /* zzflate.h:8 */

enum {
    ZZO_DEFECT_R1 = 1,   /* a FirstPass batch produced no match (encoder.cpp:433-438) */
    ZZO_DEFECT_R2 = 2,   /* level-1 match clamped at block end (encoder.cpp:350-354)  */
    ZZO_DEFECT_R4 = 4,   /* backward extension stopped by the start of the input      */
    ZZO_DEFECT_R6 = 8    /* backward extension reached 259 (encoder.cpp:404-424)      */
};

enum { ZZO_BLOCK_STORED = 0, ZZO_BLOCK_FIXED = 1, ZZO_BLOCK_DYNAMIC = 2 };

typedef struct zzo_chunk_info {
    int defects;            /* OR of ZZO_DEFECT_*                                            */
    int block_type;         /* type of the chunk's main block (level>=2: dynamic or stored)   */
    int n_records;          /* reference-style records (literals, backoffset, length)         */
    int n_matches;          /* records with length != 0                                       */
    int64_t block_bits;     /* LengthCounter total of the dynamic block (encoder.cpp:267-269) */
    int lit_freq[286];
    int dist_freq[30];
    int lit_len[286];
    int dist_len[30];
    int meta_len[19];
    /* optional dumps (may be NULL).  records: 3 uint32 per record; matches: start,len,dist per match */
    uint32_t* records; int max_records;
    uint32_t* matches; int max_matches;
    /* optional dump of the hash candidate of every position j in [0,n): candidate distance, 0 = none
     * (level >= 2 only; parse-independent view used to pin the GPU match finder) */
} zzo_chunk_info;

/* static tables (luts.cpp, fixedhuffmanluts.cpp) regenerated procedurally */
void zzo_tables(int16_t* lengthCode259, int8_t* lengthExtra259, int8_t* lengthExtraBits259,
                uint8_t* order19, uint8_t* extraDist30, uint8_t* extraLen286, uint16_t* distBase30,
                int32_t* codesF286x2, int32_t* lcodesF259x2, int32_t* dcodesF30x2);
int zzo_find_distance(int offset);       /* encoder.cpp:51-61 */
int zzo_read_lut(int offset);            /* encoder.h:93, luts.cpp:116 */
unsigned zzo_hash(const uint8_t* p);     /* encoder.cpp:11-17 */

/* bit writer (outputbitstream.h:83-124) */
size_t zzo_bitstream_kat(const uint64_t* bits, const int* counts, int n, int flush, uint8_t* buf, size_t cap);

/* Huffman (huffman.cpp:67-216, huffman.h:49-81) */
void zzo_calc_lengths(const int* freqs, int n, int maxLength, int* lengthsOut);
int  zzo_calc_lengths_iters(const int* freqs, int n, int maxLength, int* lengthsOut); /* returns #tree builds */
int  zzo_from_lengths(const int* lengths, int n, int* freqs19, uint8_t* recordsOut, int maxRecords);
void zzo_generate(const int* lengths, int n, int32_t* codesOut);
unsigned zzo_reverse(unsigned value, int len);
void zzo_default_table_lengths(int* out288);
void zzo_merged_length_codes(const int32_t* symbolCodes286x2, int32_t* lcodes259x2);

/* checksums (adler.cpp, crc.cpp) */
uint32_t zzo_adler32(uint32_t start, const uint8_t* d, size_t n);          /* true Adler-32 */
uint32_t zzo_adler32x_literal(uint32_t start, const uint8_t* d, size_t n); /* adler.cpp:17-43 incl. R3 overflow */
uint32_t zzo_combine(uint32_t first, uint32_t second, size_t lenSecond);   /* adler.cpp:5-15 */
uint32_t zzo_crc32(const uint8_t* d, size_t n, uint32_t start);            /* crc.cpp:24-33 */
uint32_t zzo_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);   /* new: reference has none */

/* One chunk in reference-equivalent mode, E(c) of SURVEY A.7.
 * chunk points into the full contiguous input; global_off = its offset there (dict <= global_off). */
size_t zzo_chunk_encode(const uint8_t* chunk, size_t n, size_t dict, uint64_t global_off,
                        int level, int final, uint8_t* out, size_t cap, zzo_chunk_info* info);

/* Hash candidates of a chunk as the level>=2 tokeniser sees them when every position 1..n-1 is
 * inserted (SURVEY A.2 "parallel formulation"): cand[j] = distance to the nearest earlier inserted
 * position with the same hash, 0 if none or >= 32768. */
void zzo_chunk_candidates(const uint8_t* chunk, size_t n, size_t dict, uint16_t* cand);

/* Upper bound of the chunked stream for n input bytes (header and trailer included). */
size_t zzo_bound(size_t n, int level, size_t chunk);

/* The product's stream definition: header || E(0) || E(1) || ... || trailer, chunks of `chunk` bytes,
 * each primed with the preceding min(dict, offset) bytes.  Returns bytes written or ~0 on error. */
size_t zzo_stream_chunked(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                          int format, int level, size_t chunk, size_t dict, int* defects);

/* Restatement of the reference's NON-threaded whole-stream path (zzflate.cpp:84-95,225): one Encoder,
 * level>=2 blocks capped at 500000 bytes with hash-table carry-over (encoder.cpp:516-523,248).
 * Exists to pin the restatement against ZzFlateEncode on the reference's own corpus.
 * `src` needs >= 8 readable bytes after src+n. */
size_t zzo_stream_reference(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                            int format, int level, int* defects);

/* Multi-threaded CPU run of the chunked stream (pthread pool over chunks): bench.py's cpu_baseline
 * when kind == "port".  Same bytes as zzo_stream_chunked. */
size_t zzo_stream_chunked_mt(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                             int format, int level, size_t chunk, size_t dict, int threads);

#ifdef __cplusplus
}
#endif
#endif
