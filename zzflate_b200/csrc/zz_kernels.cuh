// Shared declarations between the sm_100a kernels (zz_kernels.cu) and the C-ABI layer (zz_cabi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zz {

constexpr int kHashBits = 13;                 // encoder.h:41
constexpr int kHashSize = 1 << kHashBits;
constexpr int kMaxChunk = 65536;
constexpr int kMaxDict = 32768;
constexpr int kMaxMatch = 258;                // encoder.h:47
constexpr int kMaxDistance = 32768;           // encoder.h:46
constexpr int kBatch = 16384;                 // encoder.cpp:227
constexpr int kPreExtra = 288;                // history kept beyond the dictionary for backward extension (<= 259 + word slop)
constexpr int kMaxTokens = kMaxChunk / 4 + 8; // every match covers >= 4 positions
constexpr int kHistStride = 320;              // 286 lit/len + 30 dist + pad
constexpr int kHdrBytes = 640;                // >= (17 + 57 + 316*14) / 8

// Per-chunk products of the Huffman stage (K-HUFF), consumed by K-EMIT.
struct ChunkCodes {
    uint32_t lit[286];        // bits | len << 16   (code stored bit-reversed, LSB-first ready; huffman.h:74-78)
    uint32_t dist[30];
    uint8_t lens[336];        // 286 lit/len, 30 dist, 19 code-length-code lengths (debug tap / tests)
    uint8_t hdr[kHdrBytes];   // dynamic block header bit string (3 + 14 + 57 + RLE'd lengths), LSB-first
};

struct ChunkState {
    uint32_t ntok;            // matches found by K-MATCH
    uint32_t flags;
    uint32_t block_type;      // 0 stored, 1 fixed, 2 dynamic
    uint32_t hdr_bits;        // bits in ChunkCodes::hdr
    uint64_t total_bits;      // LengthCounter total of the dynamic block (encoder.cpp:267-269)
    uint32_t out_bytes;       // size of E(c)
    uint32_t nlit;            // literals of the main block (K-MATCH writes them in order into the chunk's info row)
    uint64_t out_off;         // byte offset of E(c) in the output stream
};

struct Job {
    const uint8_t* src;       // stream position 0 of this call
    uint64_t n;               // bytes in this call
    uint64_t history;         // bytes readable before src that belong to the same stream
    uint32_t chunk, dict;
    uint64_t first_chunk;     // first chunk of this batch
    uint32_t nchunks;         // chunks in this batch
    int final_stream;         // the call's last chunk carries BFINAL
    int level;
    int want_checksums;       // bit0 adler, bit1 crc
    int mode;                 // 0: reference-equivalent (E-mode), 1: free mode (header counts trimmed to the codes in use)
    // scratch, indexed by slot = chunk - first_chunk
    uint16_t* cand;           // [slots][chunk]     candidate distance per position, 0 = none
    uint8_t* info;            // [slots][chunk]     K-INFO: 0 = unusable, else 1 + min(forward match length, 32); then the literal stream
    uint32_t* tokA;           // [slots][kMaxTokens] start | length << 16
    uint16_t* tokD;           // [slots][kMaxTokens] distance
    uint32_t* hist;           // [slots][kHistStride]
    ChunkCodes* codes;        // [slots]
    ChunkState* state;        // [slots]
    uint8_t* dst;
    uint64_t cap;
    uint64_t* total;          // [0] running output size, [1] error flags, [2] matches, [3] stored chunks
    uint32_t* ck;             // [2 * chunks of the whole call]: Adler-32 with start 0, stand-alone CRC-32
};

// launch wrappers (zz_kernels.cu); each returns the number of kernels launched
int launch_candidates(const Job& job, cudaStream_t s);
int launch_lz(const Job& job, cudaStream_t s);        // match info + greedy parse, fused
int launch_huffman(const Job& job, cudaStream_t s);
int launch_offsets(const Job& job, cudaStream_t s);
int launch_emit(const Job& job, cudaStream_t s);
int launch_checksums(const Job& job, cudaStream_t s);
int launch_stored(const Job& job, cudaStream_t s);    // level 0: stored blocks + checksum partials, one kernel
int launch_fixed(const Job& job, cudaStream_t s);     // level 1: bits into the scratch slot
int launch_gather(const Job& job, cudaStream_t s);    // level 1: scratch slot -> stream
cudaError_t configure_kernels();
bool set_kernel_option(const char* name, int value);   // true if the name is known

// host-side checksum folds
uint32_t adler32_combine(uint32_t first, uint32_t second, size_t lenSecond);
uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
uint32_t crc32_shift_operator(uint64_t len);
uint32_t crc32_apply_shift(uint32_t crc, uint32_t op);

}  // namespace zz
