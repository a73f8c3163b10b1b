/* zzgpu.h -- C-ABI seam between the host driver (zzflate.h / encoder.h API) and the sm_100a kernels.
 *
 * Plain C: pointers and sizes only, explicit status codes, no STL and no torch types.  These are the
 * entry points a maintainer of the reference would bind in place of its CPU hot path:
 *
 *   zzgpu_deflate / zzgpu_deflate_ex   replace  WriteDeflateStream           zzflate/zzflate.cpp:81-156
 *                                      (i.e. Encoder::AddData -> WriteDeflateBlock -> WriteBlock2Pass /
 *                                       WriteBlockFixedHuff / WriteUncompressedBlock,
 *                                       zzflate/encoder.cpp:539,506,217,329,482, over all partitions,
 *                                       plus the stitch of zzflate.cpp:136-154)
 *   zzgpu_checksums                    replaces adler32x / crc32              zzflate/adler.cpp:17, crc.cpp:24
 *   zzgpu_adler32_combine              replaces combine                       zzflate/adler.cpp:5
 *   zzgpu_crc32_combine                new (the reference can only chain through startValue, crc.cpp:24-26)
 *   zzgpu_bound                        sizing rule callers of ZzFlateEncode apply by hand (zztest/Test.cpp:147,209,254)
 *
 * Stream definition ("E-mode"): the input is cut into `chunk`-byte pieces; chunk c is encoded
 * exactly as a fresh reference Encoder would encode it after priming its hash table with the preceding
 * min(dict, offset) bytes, non-final chunks ending with the reference's own 1-byte stored block so every
 * chunk ends byte-aligned (zzflate.cpp:116-120).  Output = E(0) || E(1) || ...  (raw deflate; the zlib /
 * gzip header and trailer are added by the host driver).
 *
 * There is no CPU fallback: every entry point that computes returns ZZGPU_E_NO_DEVICE when no CUDA
 * device is usable.
 */
#ifndef ZZGPU_H
#define ZZGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ZZGPU_API __attribute__((visibility("default")))
#else
#define ZZGPU_API
#endif

enum {
    ZZGPU_OK = 0,
    ZZGPU_E_NO_DEVICE = 1,   /* no usable CUDA device / driver */
    ZZGPU_E_CUDA = 2,        /* a CUDA call failed; see zzgpu_last_error() */
    ZZGPU_E_ARG = 3,         /* bad level / chunk / dict / NULL pointer */
    ZZGPU_E_CAPACITY = 4,    /* destination too small (the reference silently truncates here) */
    ZZGPU_E_NOMEM = 5
};

enum { ZZGPU_MEM_HOST = 0, ZZGPU_MEM_DEVICE = 1 };

#define ZZGPU_DEFAULT_CHUNK 65536u
#define ZZGPU_DEFAULT_DICT  32768u
#define ZZGPU_MAX_CHUNK     65536u
#define ZZGPU_MAX_DICT      32768u

#define ZZGPU_NSTAGES 10

/* Per-call statistics (all optional outputs). */
typedef struct zzgpu_stats {
    uint64_t chunks;          /* chunks encoded */
    uint64_t stored_chunks;   /* chunks whose main block took the stored fallback (encoder.cpp:271-274) */
    uint64_t matches;         /* LZ77 matches emitted */
    uint64_t kernel_launches; /* kernels launched by this call */
    float device_ms;          /* CUDA-event time of the device pipeline (kernels only) */
    float total_ms;           /* CUDA-event time including H2D / D2H copies when buffers are on the host */
    uint64_t h2d_bytes, d2h_bytes;
    /* CUDA-event time per pipeline stage, summed over the call's batches (ZZGPU_STAGE_*) */
    float stage_ms[ZZGPU_NSTAGES];
    uint32_t stage_launches[ZZGPU_NSTAGES];
} zzgpu_stats;

enum { ZZGPU_STAGE_CAND = 0, ZZGPU_STAGE_PARSE = 1, ZZGPU_STAGE_HUFF = 2, ZZGPU_STAGE_OFFS = 3, ZZGPU_STAGE_EMIT = 4,
       ZZGPU_STAGE_CKSUM = 5, ZZGPU_STAGE_FIXED = 6, ZZGPU_STAGE_GATHER = 7, ZZGPU_STAGE_INFO = 8, ZZGPU_STAGE_LZ = 9 };

/* Select the device used by the calling thread's subsequent calls (default: current CUDA device).
 * Creates the per-device context (stream, scratch) lazily.  Returns ZZGPU_E_NO_DEVICE without a GPU. */
ZZGPU_API int zzgpu_init(int device);
ZZGPU_API void zzgpu_shutdown(void);
ZZGPU_API int zzgpu_device_count(void);
ZZGPU_API const char* zzgpu_strerror(int status);
ZZGPU_API const char* zzgpu_last_error(void);

/* Tuning knobs that do not change the produced bytes: "tma", "spec", "region" (K-LZ variants, 0/1), "lanes" (1-4 kernel
 * streams of the host-buffer path), "piece_first" / "piece_mid" / "piece_last" (piece schedule, chunks), "segment_mib".
 * Unknown names return ZZGPU_E_ARG. */
ZZGPU_API int zzgpu_set_option(const char* name, int value);

/* Diagnostic counters of the calling thread's most recent streaming call (tests):
 *   "sink_pieces"            pieces (H2D -> kernels -> D2H units) of the call
 *   "sink_first_h2d_done"    pieces whose host-to-device copy had completed when the first data slice reached the sink
 * Unknown names return -1. */
ZZGPU_API long long zzgpu_get_counter(const char* name);

/* Worst-case size of the raw deflate stream for n input bytes (A.6: 65 546 bytes per 65 536-byte chunk at
 * levels 0/2/3; 9 bits per literal at level 1). */
ZZGPU_API size_t zzgpu_bound(size_t n, int level, uint32_t chunk);

/* Raw deflate of src[0,n) as one complete stream (last chunk carries BFINAL).
 * src_mem / dst_mem: ZZGPU_MEM_HOST or ZZGPU_MEM_DEVICE.  adler / crc (optional) receive the Adler-32
 * (start value 1) and CRC-32 of the input.  Returns ZZGPU_OK and *out_len, or an error status. */
ZZGPU_API int zzgpu_deflate(const uint8_t* src, size_t n, int src_mem,
                            uint8_t* dst, size_t cap, int dst_mem,
                            int level, uint32_t chunk, uint32_t dict,
                            size_t* out_len, uint32_t* adler, uint32_t* crc, zzgpu_stats* stats);

/* One shard of a larger stream (multi-GPU / streaming use): `history` bytes are readable immediately
 * before src and belong to the same stream (they prime the first chunk's dictionary); `final` tells whether
 * this shard ends the stream.  Checksums returned are those of the shard alone, computed with start
 * value 0 (Adler) / as a stand-alone CRC, ready for zzgpu_adler32_combine / zzgpu_crc32_combine.
 * `want_checksums`: bit0 Adler-32, bit1 CRC-32. */
ZZGPU_API int zzgpu_deflate_ex(const uint8_t* src, size_t n, size_t history, int final, int src_mem,
                               uint8_t* dst, size_t cap, int dst_mem,
                               int level, uint32_t chunk, uint32_t dict, int want_checksums,
                               size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats);

/* zzgpu_deflate_ex with a mode.  0 = reference-equivalent ("E-mode": every chunk byte-identical to the reference encoder,
 * what every other entry point produces).  1 = free mode: the same tokens and code lengths, but the dynamic
 * block header carries HLIT / HDIST / HCLEN trimmed to the codes in use (the reference always writes 286 / 30 / 19,
 * encoder.cpp:283-290), which also shortens the run-length coded length sequence.  (A per-chunk choice of the fixed code
 * was measured and dropped: with this parse -- the last 258 bytes of a block are always literals, encoder.cpp:222 -- it
 * never beat both the dynamic and the stored form.)  Levels 2 and 3; other levels ignore the mode. */
ZZGPU_API int zzgpu_deflate_mode(const uint8_t* src, size_t n, size_t history, int final, int src_mem,
                                 uint8_t* dst, size_t cap, int dst_mem,
                                 int level, uint32_t chunk, uint32_t dict, int want_checksums, int mode,
                                 size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats);

/* Same as zzgpu_deflate_ex on host buffers, with the `hist_len` bytes of preceding stream given by their own pointer
 * (they need not lie in front of src).  Used by Encoder::AddData, which keeps a private copy of the last
 * 32 KiB + 288 bytes it was fed (the reference keeps its hash table across calls, encoder.cpp:248,320-327). */
ZZGPU_API int zzgpu_deflate_hist(const uint8_t* src, size_t n, const uint8_t* hist, size_t hist_len, int final,
                                 uint8_t* dst, size_t cap, int level, uint32_t chunk, uint32_t dict,
                                 size_t* out_len);

/* Streaming variant (replaces the growable-buffer mode of outputbitstream behind ZzFlateEncodeToCallback,
 * zzflate/zzflate.cpp:197-222, outputbitstream.h:171-190): host input, the stream is handed to `sink` in order, in
 * slices of at most `slice` bytes (0 = 1 000 000, the reference's buffer size), while later pieces are still being
 * copied in and encoded.  Slice pointers are valid only during the sink call; its return value is ignored. */
typedef int (*zzgpu_sink_fn)(const uint8_t* data, size_t len, void* user);
ZZGPU_API int zzgpu_deflate_sink(const uint8_t* src, size_t n, size_t history, int final,
                                 int level, uint32_t chunk, uint32_t dict, int want_checksums,
                                 zzgpu_sink_fn sink, void* user, size_t slice,
                                 size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats);

/* Two-phase variant for shards whose place in the caller's buffer depends on earlier shards (stitch of
 * zzflate.cpp:136-154 without a temporary and without a host-side memmove): zzgpu_deflate_hold encodes host input
 * and keeps the stream in device memory, reporting its size; zzgpu_fetch then copies it to its final host address
 * (dst != NULL) or hands it to a sink (dst == NULL), and releases the device's context.  Between the two calls the
 * calling thread owns the device's context: other threads' calls on that device wait.  zzgpu_release drops a held
 * stream without fetching it. */
ZZGPU_API int zzgpu_deflate_hold(const uint8_t* src, size_t n, size_t history, int final,
                                 int level, uint32_t chunk, uint32_t dict, int want_checksums,
                                 size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats);
ZZGPU_API int zzgpu_fetch(uint8_t* dst, size_t cap, zzgpu_sink_fn sink, void* user, size_t slice);
ZZGPU_API void zzgpu_release(void);

/* Size limits: host-buffer calls are processed in segments of at most 2 GiB of input (device staging is a ring of
 * that size, so any input length works); device-resident calls use the caller's buffers in place; zzgpu_deflate_hold
 * keeps zzgpu_bound(n) bytes of device memory until the fetch. */

/* Adler-32 (continuing from adler_start, reference convention adler32x(start,...)) and CRC-32
 * (continuing from crc_start, reference convention crc32(buf,len,start)) of a buffer, on the GPU. */
ZZGPU_API int zzgpu_checksums(const uint8_t* src, size_t n, int src_mem,
                              uint32_t adler_start, uint32_t crc_start,
                              uint32_t* adler, uint32_t* crc);

/* Host-side O(1)/O(log n) folds used when stitching shards. */
ZZGPU_API uint32_t zzgpu_adler32_combine(uint32_t first, uint32_t second_start0, size_t len_second);
ZZGPU_API uint32_t zzgpu_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);

/* Debug / test taps: run the level>=2 pipeline on a device-resident or host buffer and copy out the
 * intermediate per-chunk products of chunk `chunk_index` (any pointer may be NULL):
 *   cand      uint16[chunk]       hash candidate distance per position (0 = none)
 *   tokens    uint32[3*max]       (start, length, distance) per match, *n_tokens receives the count
 *   hist      uint32[316]         literal/length (286) then distance (30) frequencies
 *   lengths   uint8[335]          code lengths: 286 lit/len, 30 dist, 19 code-length code
 *   info      uint32[4]           block_type, header_bits, out_bytes, total_bits(low 32) */
ZZGPU_API int zzgpu_debug_chunk(const uint8_t* src, size_t n, int src_mem, int level,
                                uint32_t chunk, uint32_t dict, uint64_t chunk_index,
                                uint16_t* cand, uint32_t* tokens, uint32_t max_tokens, uint32_t* n_tokens,
                                uint32_t* hist, uint8_t* lengths, uint32_t* info);

#ifdef __cplusplus
}
#endif
#endif
