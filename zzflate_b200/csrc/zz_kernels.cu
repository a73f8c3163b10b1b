// sm_100a kernels of the deflate-encode hot path.
//
// Pipeline per batch of chunks (one chunk = up to 64 KiB of input + up to 32 KiB of dictionary):
//
//   K-CAND   k_candidates  hash heads: nearest earlier position with the same 13-bit hash, per position
//                          (replaces CalcHash / the table probe+insert of FirstPass / AddHashEntries,
//                           zzflate/encoder.cpp:11-17,388-390,474-480)
//   K-LZ     k_lz          per-position match info against the candidate (the compare part of FirstPass: remain /
//                          countMatchBackward; encoder.cpp:81-102,391-403) and, fused with it, the greedy acceptance as the
//                          orbit of a successor function, exact lengths and backward extension of the taken matches,
//                          histograms, literal stream (FirstPass, GetFrequencies; encoder.cpp:375-471); the window of every
//                          sub-batch arrives in shared memory by TMA bulk copies (cp.async.bulk + mbarrier)
//   K-HUFF   k_huffman     code lengths (libstdc++-heap Huffman with the reference's limiter), canonical
//            k_huffman_lanes  codes, code-length RLE, exact block size, stored fallback decision, block header
//                          (huffman.cpp:67-216, huffman.h:49-81, encoder.cpp:171-187,250-293); warp per chunk for
//                          small launches, thread per chunk for full batches
//   K-OFFS   k_offsets     exclusive scan of chunk sizes -> output offsets (stitch of zzflate.cpp:136-154)
//   K-EMIT   k_emit2       bit emission of header, records, EOB and the aligning stored block
//                          (WriteRecords / WriteDistance / StartBlock / outputbitstream; encoder.cpp:135-169,
//                           UncompressedFallback / WriteUncompressedBlock; encoder.cpp:305-317,482-502), symbol-parallel
//   K-FIXED  k_fixed, k_gather   level 1: WriteBlockFixedHuff (encoder.cpp:329-373)
//   K-CKSUM  k_checksums   per-chunk Adler-32 / CRC-32 partials (adler.cpp:17, crc.cpp:24)
//   K-STORED k_stored      level 0: stored blocks and the checksum partials in one kernel (encoder.cpp:482-502)
//
// Everything is integer work; nothing here is a dense contraction, so tensor cores are not used.
#include "zz_kernels.cuh"
#include <cstdio>
#include <cstring>

namespace zz {

namespace {

// ------------------------------------------------------------------------------------------------
// geometry of one chunk
// ------------------------------------------------------------------------------------------------
struct Geom {
    long long off;   // offset of the chunk in this call's input
    int n;           // bytes in the chunk
    int dict;        // dictionary bytes that prime the hash table
    int pre;         // readable history kept in the window (dict + slack for backward extension)
    int body;        // bytes of the main block: n-1 for non-final chunks (zzflate.cpp:116), n for the final one
    int final;
    int t0;          // tokeniser target: positions >= t0 are never probed (encoder.cpp:222)
};

__device__ __forceinline__ Geom chunk_geom(const Job& job, unsigned slot)
{
    Geom g;
    unsigned long long c = job.first_chunk + slot;
    g.off = (long long)(c * job.chunk);
    unsigned long long rem = job.n - (unsigned long long)g.off;
    g.n = (int)(rem < job.chunk ? rem : job.chunk);
    unsigned long long before = (unsigned long long)g.off + job.history;
    g.dict = (int)(before < job.dict ? before : job.dict);
    unsigned long long pre = (unsigned long long)g.dict + kPreExtra;
    g.pre = (int)(before < pre ? before : pre);
    g.final = ((unsigned long long)g.off + g.n == job.n) && job.final_stream;
    g.body = g.final ? g.n : g.n - 1;
    g.t0 = g.body > kMaxMatch ? g.body - kMaxMatch : 0;
    return g;
}

__device__ __forceinline__ unsigned hash3(unsigned v24)            // encoder.cpp:11-17
{
    return (v24 * 0x00d68664u) >> (32 - kHashBits);
}

// distance -> distance symbol / extra bits (luts.cpp:64,79-116 regenerated arithmetically)
__device__ __forceinline__ int dist_symbol(int d, int& extraBits, int& extraVal)
{
    int t = d - 1;
    if (t < 4) { extraBits = 0; extraVal = 0; return t; }
    int nb = 31 - __clz(t);
    extraBits = nb - 1;
    extraVal = t & ((1 << extraBits) - 1);
    return 2 * nb + ((t >> extraBits) & 1);
}

// match length -> length symbol / extra bits (luts.cpp:5-58)
__device__ __forceinline__ int len_symbol(int len, int& extraBits, int& extraVal)
{
    int l = len - 3;
    if (l < 8) { extraBits = 0; extraVal = 0; return 257 + l; }
    if (len == 258) { extraBits = 0; extraVal = 0; return 285; }
    int nb = 31 - __clz(l);
    extraBits = nb - 2;
    extraVal = l & ((1 << extraBits) - 1);
    return 257 + 4 * extraBits + 4 + ((l >> extraBits) & 3);
}

// ------------------------------------------------------------------------------------------------
// K-CAND : hash heads.  One warp walks a run of consecutive chunks, 32 positions per step:
//   candidate(j) = nearest earlier inserted position with the same 13-bit hash
//                = nearest lower lane of the same step with that hash (__match_any_sync), else table[h].
// Every position is inserted except position 0 of each chunk (never probed nor inserted, encoder.cpp:384);
// the positions before the first chunk of the run only prime the table (AddHashEntries, encoder.cpp:474).
//
// The table holds stream positions modulo 65536 in 16 bits (16 KiB per warp instead of the reference's
// 32 KiB of ints): distances below 65536 are exact modulo 65536, and every 32768 positions a sweep retires the
// entries that are 32768 or more behind (invalid from then on, encoder.cpp:392), so nothing ever aliases.
// With the default geometry (64 KiB chunks, 32 KiB dictionary) the table a fresh reference Encoder would
// hold after priming chunk c equals, on every entry that can still yield a valid distance, the table after
// walking chunk c-1 -- so a run of chunks needs one priming pass only.
// Input is staged through shared memory with cp.async (LDGSTS), double buffered, 1 KiB tiles.
// ------------------------------------------------------------------------------------------------
constexpr int kCandTile = 1024;
constexpr int kEmptySlot = -(1 << 30);
constexpr int kCandStage = kCandTile + 16;
constexpr int kCandFlight = 32;           // steps of 32 positions whose table exchanges are in flight together

__device__ __forceinline__ void cp_async4(void* smemDst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// stages bytes gsrc[0, nbytes) at st + ((uintptr)gsrc & 3); bytes outside [validLo, validHi) read as zero
__device__ __forceinline__ void stage_tile(uint8_t* st, const uint8_t* gsrc, int nbytes, const uint8_t* validLo,
                                           const uint8_t* validHi, int lane)
{
    const int phase = (int)(reinterpret_cast<uintptr_t>(gsrc) & 3);
    const uint8_t* g0 = gsrc - phase;
    const int words = (phase + nbytes + 3) >> 2;
    for (int w = lane; w < words; w += 32) {
        const uint8_t* ga = g0 + 4 * w;
        if (ga >= validLo && ga + 4 <= validHi) {
            cp_async4(st + 4 * w, ga);
        } else {
            unsigned v = 0;
            for (int k = 0; k < 4; ++k) if (ga + k >= validLo && ga + k < validHi) v |= (unsigned)ga[k] << (8 * k);
            *reinterpret_cast<unsigned*>(st + 4 * w) = v;
        }
    }
    cp_async_commit();
}

__global__ void __launch_bounds__(32) k_candidates(Job job, int run)
{
    __shared__ int table[kHashSize];
    __shared__ __align__(16) uint8_t stage[2][kCandStage];
    const int lane = threadIdx.x;
    const unsigned firstSlot = blockIdx.x * (unsigned)run;
    unsigned lastSlot = firstSlot + (unsigned)run; if (lastSlot > job.nchunks) lastSlot = job.nchunks;
    const Geom g0 = chunk_geom(job, firstSlot);
    const Geom gl = chunk_geom(job, lastSlot - 1);
    const uint8_t* base0 = job.src + g0.off;                                  // position q = 0 of the run
    // positions q in [-dict, qEnd) relative to the run start; runs longer than one chunk only with 64 KiB chunks,
    // where the chunk-relative position is q & 0xFFFF and the candidate rows of the run are contiguous
    const int qEnd = (int)(lastSlot - 1 - firstSlot) * (int)job.chunk + gl.n;
    const int qStart = -g0.dict;
    const unsigned startMask = run > 1 ? 0xFFFFu : 0xFFFFFFFFu;               // q & startMask == 0 <=> first byte of a chunk
    uint16_t* candOut = job.cand + (size_t)firstSlot * job.chunk;
    const uint8_t* validLo = job.src - job.history;
    const uint8_t* validHi = job.src + job.n;
    const unsigned ltMask = (1u << lane) - 1u;

    for (int i = lane; i < kHashSize; i += 32) table[i] = kEmptySlot;
    int buf = 0;
    stage_tile(stage[0], base0 + qStart, kCandTile + 8, validLo, validHi, lane);
    for (int q0 = qStart; q0 < qEnd; q0 += kCandTile) {
        cp_async_wait_all();
        __syncwarp();
        if (q0 + kCandTile < qEnd) stage_tile(stage[buf ^ 1], base0 + q0 + kCandTile, kCandTile + 8, validLo, validHi, lane);
        const unsigned* sw = reinterpret_cast<const unsigned*>(stage[buf]);
        const int phase = (int)(reinterpret_cast<uintptr_t>(base0 + q0) & 3);
        const int tileEnd = min(q0 + kCandTile, qEnd);
        // Fast path: a full tile of positions that are all probed and inserted (no priming, no chunk start, no tail).
        // kCandFlight steps (a whole 1 KiB tile) are in flight at once: their hashes are independent, the exchanges are issued
        // back to back (shared-memory operations of one warp complete in program order, so step u+1 sees the slots as
        // step u left them) and the ascending-order check is made once for the group.  If any lane received a position
        // above its own, the slots the group touched are restored from the pre-group values the lanes hold (exactly one
        // lane per touched slot received a position below the group) and the group is redone step by step with explicit
        // same-hash group resolution.
        if (q0 > 0 && q0 + kCandTile <= qEnd && ((unsigned)q0 & startMask) != 0) {
            const int ob = phase + lane;
            const unsigned* swl = sw + (ob >> 2);
            const int sh = (ob & 3) * 8;
            uint16_t* outp = candOut + q0 + lane;
            constexpr int U = kCandFlight, G = 32 * U;        // steps in flight, positions per group
            for (int g4 = 0; g4 < kCandTile / G; ++g4) {
                const int qs = q0 + g4 * G;
                unsigned h[U]; int old[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned v = __funnelshift_r(swl[g4 * (G / 4) + u * 8], swl[g4 * (G / 4) + u * 8 + 1], sh) & 0xFFFFFFu;
                    h[u] = hash3(v);
                }
                // A run of one byte value (all G hashes equal): every position's candidate is its predecessor and the last
                // one stays in the slot -- one exchange instead of G exchanges on the same word.
                if (__all_sync(0xffffffffu, h[0] == h[U - 1] && h[0] == h[U / 2])) {
                    bool uni = h[0] == (unsigned)__shfl_sync(0xffffffffu, (int)h[0], 0);
#pragma unroll
                    for (int u = 1; u < U; ++u) uni &= h[u] == h[0];
                    if (__all_sync(0xffffffffu, uni)) {
                        int pre = 0;
                        if (lane == 0) pre = atomicExch(&table[h[0]], qs + G - 1);
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int d = (u == 0 && lane == 0) ? qs - pre : 1;
                            outp[g4 * G + u * 32] = (uint16_t)(d < kMaxDistance ? d : 0);
                        }
                        continue;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) old[u] = atomicExch(&table[h[u]], qs + u * 32 + lane);
                bool bad = false;
#pragma unroll
                for (int u = 0; u < U; ++u) bad |= old[u] > qs + u * 32 + lane;
                if (__ballot_sync(0xffffffffu, bad)) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (old[u] < qs) table[h[u]] = old[u];
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int pre = table[h[u]];
                        const unsigned grp = __match_any_sync(0xffffffffu, h[u]);
                        const unsigned lower = grp & ltMask;
                        old[u] = lower ? qs + u * 32 + 31 - __clz(lower) : pre;
                        __syncwarp();
                        if ((grp >> lane) == 1u) table[h[u]] = qs + u * 32 + lane;
                        __syncwarp();
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int d = qs + u * 32 + lane - old[u];
                    outp[g4 * G + u * 32] = (uint16_t)(d < kMaxDistance ? d : 0);
                }
            }
            buf ^= 1;
            continue;
        }
#pragma unroll 2
        for (int qs = q0; qs < tileEnd; qs += 32) {
            const int q = qs + lane;
            const int o = phase + (q - q0);
            const unsigned v = __funnelshift_r(sw[o >> 2], sw[(o >> 2) + 1], (o & 3) * 8) & 0xFFFFFFu;
            const unsigned h = hash3(v);
            const bool inRange = q < qEnd;
            // position 0 of a chunk is neither probed nor inserted (encoder.cpp:384); positions before the run
            // only prime the table (AddHashEntries, encoder.cpp:474)
            const bool act = inRange && (q < 0 || ((unsigned)q & startMask) != 0);
            // One shared-memory exchange per position does probe and insert at once.  Lanes of a step that share a
            // hash are serialised by the hardware; if that happens in ascending lane order each lane receives the
            // position of the nearest lower lane with its hash (or the slot's previous content) -- exactly the
            // candidate -- and the highest lane's position stays in the slot.  Any other order hands some lane a
            // position above its own, which is detected and the step is redone with explicit group resolution.
            int old = act ? atomicExch(&table[h], q) : kEmptySlot;
            if (__ballot_sync(0xffffffffu, act && old > q)) {
                const unsigned grp = __match_any_sync(0xffffffffu, act ? h : (0x10000u + lane));
                const unsigned fromBefore = __ballot_sync(0xffffffffu, act && old < qs) & grp;   // exactly one lane per group
                const int pre = __shfl_sync(0xffffffffu, old, fromBefore ? __ffs(fromBefore) - 1 : lane);
                const unsigned lower = grp & ltMask;
                old = lower ? qs + 31 - __clz(lower) : pre;
                __syncwarp();
                if (act && (grp >> lane) == 1u) table[h] = q;       // the highest position of a group owns the slot
                __syncwarp();
            }
            if (inRange && q >= 0) {
                const int d = q - old;
                candOut[q] = (uint16_t)((act && d < kMaxDistance) ? d : 0);
            }
        }
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------
// shared-memory window helpers
// ------------------------------------------------------------------------------------------------
constexpr int kPreCap = kMaxDict + kPreExtra;                       // 33056, multiple of 16
static_assert(kPreCap % 16 == 0, "window base must keep 16-byte phase");

// Copies src[lo, hi) (positions relative to the chunk start) into win so that position i lands at byte
// wb + i, where wb = kPreCap + ((uintptr)(chunk start) & 15): global and shared 16-byte phases agree and the
// interior moves as 128-bit loads/stores.
__device__ __forceinline__ void load_window(uint8_t* win, int wb, const uint8_t* chunk0, int lo, int hi, int padTo)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    int s_lo = wb + lo, s_hi = wb + hi;
    int a_lo = (s_lo + 15) & ~15, a_hi = s_hi & ~15;
    if (a_lo >= a_hi) {
        for (int s = s_lo + tid; s < s_hi; s += nt) win[s] = chunk0[s - wb];
    } else {
        for (int s = s_lo + tid; s < a_lo; s += nt) win[s] = chunk0[s - wb];
        for (int s = a_hi + tid; s < s_hi; s += nt) win[s] = chunk0[s - wb];
        const uint4* gsrc = reinterpret_cast<const uint4*>(chunk0 + (a_lo - wb));
        uint4* sdst = reinterpret_cast<uint4*>(win + a_lo);
        const int nvec = (a_hi - a_lo) >> 4;
        for (int k = tid; k < nvec; k += nt) sdst[k] = __ldg(gsrc + k);
    }
    for (int s = s_hi + tid; s < wb + padTo; s += nt) win[s] = 0;
}

__device__ __forceinline__ unsigned ld4(const uint8_t* win, int o)          // unaligned 4-byte load
{
    const unsigned* w = reinterpret_cast<const unsigned*>(win) + (o >> 2);
    return __funnelshift_r(w[0], w[1], (o & 3) * 8);
}

__device__ __forceinline__ unsigned long long ld8(const uint8_t* win, int o)
{
    const unsigned* w = reinterpret_cast<const unsigned*>(win) + (o >> 2);
    unsigned a = w[0], b = w[1], c = w[2];
    int sh = (o & 3) * 8;
    unsigned lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    return ((unsigned long long)hi << 32) | lo;
}

// ------------------------------------------------------------------------------------------------
// match info helpers (the compare part of FirstPass, encoder.cpp:391-403)
// ------------------------------------------------------------------------------------------------
constexpr int kCapLen = 32;                        // cap of the parallel forward compare
// forward match length beyond the first 4 bytes, capped at kCapLen - 4 (oj, op already advanced by 4)
__device__ __forceinline__ int fwd_more(const uint8_t* win, int oj, int op)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned long long x = ld8(win, oj + 8 * k) ^ ld8(win, op + 8 * k);
        const unsigned xl = (unsigned)x, xh = (unsigned)(x >> 32);
        if (xl) return 8 * k + ((__ffs(xl) - 1) >> 3);
        if (xh) return 8 * k + 4 + ((__ffs(xh) - 1) >> 3);
    }
    const unsigned x1 = ld4(win, oj + 24) ^ ld4(win, op + 24);
    return x1 ? 24 + ((__ffs(x1) - 1) >> 3) : kCapLen - 4;
}

// ------------------------------------------------------------------------------------------------
// helpers of the parse
// ------------------------------------------------------------------------------------------------
constexpr int kLongGap = kMaxMatch - kCapLen + 1;   // literal gap from which fwd + backward extension can exceed 258

// succ(b) = j + fwd holds while the match length fwd + lb stays below the 258 cap.  A match is measured exactly when its
// forward part reached the 32-byte compare cap, or when the pending literal run is so long that the backward extension
// alone could push fwd + lb over 258 (then the next state is j - lb + 258 < j + fwd).
__device__ __forceinline__ bool needs_exact(int fwd, int gap) { return fwd >= kCapLen || gap >= kLongGap; }
constexpr unsigned kNone16 = 0xFFFFu;

// unaligned 8-byte little-endian load from global memory; bytes outside [lo, hi) read as zero
__device__ __forceinline__ unsigned long long gload8(const uint8_t* p, const uint8_t* lo, const uint8_t* hi)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint8_t* al = reinterpret_cast<const uint8_t*>(a & ~(uintptr_t)7);
    if (al >= lo && al + 16 <= hi) {
        const unsigned long long x = __ldg(reinterpret_cast<const unsigned long long*>(al));
        const unsigned long long y = __ldg(reinterpret_cast<const unsigned long long*>(al) + 1);
        const int sh = (int)(a & 7) * 8;
        return sh ? ((x >> sh) | (y << (64 - sh))) : x;
    }
    unsigned long long v = 0;
    for (int k = 0; k < 8; ++k) if (p + k >= lo && p + k < hi) v |= (unsigned long long)p[k] << (8 * k);
    return v;
}


// unaligned 4-byte little-endian load from global memory; bytes outside [lo, hi) read as zero
__device__ __forceinline__ unsigned gload4(const uint8_t* p, const uint8_t* lo, const uint8_t* hi)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint8_t* al = reinterpret_cast<const uint8_t*>(a & ~(uintptr_t)3);
    if (al >= lo && al + 8 <= hi) {
        const unsigned x = __ldg(reinterpret_cast<const unsigned*>(al));
        const unsigned y = __ldg(reinterpret_cast<const unsigned*>(al) + 1);
        return __funnelshift_r(x, y, (int)(a & 3) * 8);
    }
    unsigned v = 0;
    for (int k = 0; k < 4; ++k) if (p + k >= lo && p + k < hi) v |= (unsigned)p[k] << (8 * k);
    return v;
}

// [lo, hi) is the readable stream
struct Stream { const uint8_t* lo; const uint8_t* hi; };

// First position the walk would take from state b (b = end of the previous match), or -1.
// info == 0 marks an unusable position, otherwise info - 1 = min(forward length, 32); a usable position j is
// taken when j - b >= 4 - min(fwd, 4).
__device__ __forceinline__ int probe_next(const uint8_t* info, const unsigned* okbits, const uint16_t* nzw,
                                          int ntiles, int base, int b)
{
    const int r = b - base;
#pragma unroll
    for (int k = 1; k <= 3; ++k)
        if ((int)info[r + k] >= 5 - k) return b + k;
    const int x = r + 4;
    const int w = x >> 5;
    if (w >= ntiles) return -1;
    const unsigned bits = okbits[w] & (~0u << (x & 31));
    if (bits) return base + w * 32 + __ffs(bits) - 1;
    const unsigned w2 = nzw[w + 1];
    if (w2 == kNone16) return -1;
    return base + (int)w2 * 32 + __ffs(okbits[w2]) - 1;
}

// ------------------------------------------------------------------------------------------------
// K-LZ : match info and greedy parse fused, one CTA per chunk (replaces K-INFO + K-MATCH; FirstPass, countMatchBackward,
// remain, GetFrequencies; encoder.cpp:81-102,375-471).
//
// The CTA walks the chunk in the reference's batches of 16 384 positions (WriteBlock2Pass, encoder.cpp:225-234).  For
// each batch the window [batch start - 33 056, batch end + 304) is brought into shared memory with one-dimensional
// bulk copies (cp.async.bulk + mbarrier: the TMA engine moves the aligned interior, threads patch the ragged edges), and
// everything the parse needs is read from there: candidate compares, exact lengths of long matches, backward extension.
// A batch is processed in sub-batches of up to 8192 positions:
//   A  info[j] = 0 / 1 + min(forward match length, 32) for every position (the compare part of FirstPass, as in the
//      former K-INFO: four positions per thread, long compares queued per warp)
//   C  A state b is the end of the previous match; from b the walk takes the first position j > b that is usable and far
//      enough (j - b >= 4 - min(fwd, 4)) and moves to succ(b) = j + fwd.  The parse is the orbit of succ from the
//      entry state.  succ is evaluated on the fly from info and a bitmap of usable positions, only for states that are
//      visited.  Orbits that start at different states merge after a few matches (the walk re-synchronises), so 128
//      lanes first follow succ speculatively, each through its own 64-state segment from the segment's first state,
//      marking what they visit (chains); each lane then follows succ from the state where its left neighbour's chain entered
//      its segment until it steps on its own chain (links).  One walk from the true entry state follows succ only until it
//      steps on a chain; from there the chains are the orbit as far as the links connect them, so the walk continues behind
//      the last connected segment.  Matches of 32 bytes or more (or behind a literal gap so long that the backward extension
//      could hit the 258 cap) end a chain; while they are few they are measured exactly by the chain's warp (all lanes
//      compare, up to 258 bytes forwards and backwards) and the chain goes on, otherwise the true walk measures them as it
//      meets them, runs of equal 258-byte matches 32 at a time.
//   D  the visited states are compacted and expanded to tokens in parallel (backward extension bounded by the pending
//      literals, encoder.cpp:404-416).
// After the last batch: histograms and the literal stream for K-EMIT, as before.
// ------------------------------------------------------------------------------------------------
constexpr int kLzThreads = 384;                           // three CTAs per SM: the serial parts of one overlap the parallel parts of the others
constexpr int kLzWarps = kLzThreads / 32;
constexpr int kSub = 4 * 4 * kLzThreads;                  // states per sub-batch (tile-aligned base up to 95 below its first position): 6144
constexpr int kLzBack = kPreCap;                          // bytes kept before the sub-batch's first probed position
constexpr int kLzAhead = 304;                             // bytes kept behind its last one (258 + 8-byte loads + slack)
constexpr int kLzWin = 32 + kLzBack + kSub + kLzAhead + 16;
static_assert(kLzWin % 16 == 0 && kLzWin < 65536, "window offsets are kept in 16 bits");
constexpr int kSubTiles = kSub / 32;
constexpr int kSubWords = kSubTiles + 5;
#ifndef ZZ_SEG_STATES
#define ZZ_SEG_STATES 64
#endif
constexpr int kSegStates = ZZ_SEG_STATES;                 // states per lane in the speculative pass
constexpr int kChains = kSub / kSegStates;
constexpr int kChaseWarps = kChains / 32;
static_assert(kChains % 32 == 0 && kChaseWarps <= 8 && kChaseWarps < kLzWarps && kSubTiles <= kLzThreads && kSub % (4 * kLzThreads) == 0, "geometry");
constexpr int kLzQueue = 32 + 128;
constexpr int kExCap = 256;                               // exactly measured long matches remembered per sub-batch
constexpr int kRegionMax = kSub + 512;                    // positions whose literals / match symbols are counted in one go beside the walk
constexpr int kRegionWords = kRegionMax / 32 + 8;
constexpr int kHistCopies = 4;                            // private histogram copies (warps share them round robin)
constexpr int kWorkerThreads = (kLzWarps - kChaseWarps) * 32;
constexpr int kLzScratch = kLzWarps * kLzQueue * 4;       // phase A: long-compare queues; afterwards: bitmaps, nzw, state list
static_assert(4 * kSubWords * 4 + 544 + (kSub / 4 + 40) * 2 <= kLzScratch, "bitmaps + nzw + state list fit the queue space");
constexpr int kLzSmem = kLzWin + (kSub + 64) + (kSub + 64) * 2 + kLzScratch;
static_assert(3 * (kLzSmem + 2048) <= 227 * 1024, "three K-LZ CTAs per SM");

// phase shares of a K-LZ CTA (diagnostic build -DZZ_PHASE_TIMING: thread 0's clock at the phase boundaries)
#ifdef ZZ_PHASE_TIMING
__device__ unsigned long long g_lzPhase[12];
#define PT(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); phaseAcc[k] += (unsigned long long)(t_ - phaseT); phaseT = t_; } } while (0)
#else
#define PT(k)
#endif
#ifdef ZZ_LZ_CHECKS
__device__ unsigned g_lzDebug[8];
#define LZ_CHECK(cond, code, a, b) do { if (!(cond)) { if (atomicCAS(&g_lzDebug[0], 0u, (unsigned)(code)) == 0u) { g_lzDebug[1] = (unsigned)(a); g_lzDebug[2] = (unsigned)(b); g_lzDebug[3] = blockIdx.x; g_lzDebug[4] = threadIdx.x; } } } while (0)
#define LZ_GUARD(var, limit, code, a, b) if (++(var) > (limit)) { LZ_CHECK(false, code, a, b); break; }
#else
#define LZ_CHECK(cond, code, a, b) do { } while (0)
#define LZ_GUARD(var, limit, code, a, b)
#endif

struct LzShared {
    int pos;            // start of the next FirstPass batch
    int ntok;
    int nexcl;          // batch starts that were never inserted into the hash table
    int excl[4];
    int npend;          // positions found in phase A whose candidate is an excluded batch start (never inserted: encoder.cpp:384)
    int pendJ[8];
    int b;              // current state of the walk
    int npre;           // tokens written directly by warp 0 in this sub-batch (first probe / far entry)
    int exN;            // entries of the exact list
    int endsAt;         // orbit state without successor in the sub-batch (it yields no token), or -1
    // GetFrequencies beside the walk: positions below histPos / tokens below histTok are counted and their literals written;
    // [histPos, regEnd) / [histTok, regTok) are final (everything below the walk's state is) and wait for the next walk
    int histPos, histTok, regEnd, regTok, histStop;
    unsigned nlits, regLits;
    int err;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// one-dimensional bulk copy global -> shared through the TMA engine; all three of dst, src, bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// named barrier of the chase warps; the non-aligned form counts arrivals per thread, so it is safe even where the compiler
// has not reconverged a warp in front of it
__device__ __forceinline__ void chase_barrier() { __syncwarp(); asm volatile("barrier.sync 1, %0;" ::"n"(kChaseWarps * 32) : "memory"); }
__device__ __forceinline__ void worker_barrier() { __syncwarp(); asm volatile("barrier.sync 2, %0;" ::"n"(kWorkerThreads) : "memory"); }

// info of one position from the window (same definition as phase A); oj = window offset of the position
__device__ int info_of_w(const uint8_t* win, int oj, int d, int room)
{
    if (d == 0) return 0;
    const int op = oj - d;
    int fwd = 0;
    while (fwd < kCapLen) {
        const unsigned x = ld4(win, oj + fwd) ^ ld4(win, op + fwd);
        if (x) { fwd += (__ffs(x) - 1) >> 3; break; }
        fwd += 4;
    }
    bool ok = fwd >= 4;
    if (!ok) {
        const unsigned y = ld4(win, oj - 4) ^ ld4(win, op - 4);
        int back = y ? (__clz(y) >> 3) : 4;
        if (back > room) back = room;                    // bytes of real history before the candidate (R4 clamp)
        ok = fwd + back >= 4;
    }
    return ok ? fwd + 1 : 0;
}

// exact forward and backward match lengths (<= 258 each), all 32 lanes cooperate (remain(), countMatchBackward;
// encoder.cpp:81-102); window offsets
__device__ __forceinline__ void coop_lengths_w(const uint8_t* win, int oj, int op, int lane, bool wantBack, int& fwd, int& lb)
{
    const unsigned long long xa = ld8(win, oj + lane * 8) ^ ld8(win, op + lane * 8);
    unsigned long long xb = 0;
    unsigned ta = 1, tb = 1;
    if (wantBack) xb = ld8(win, oj - 8 - lane * 8) ^ ld8(win, op - 8 - lane * 8);
    if (lane < 2) ta = (unsigned)(win[oj + 256 + lane] ^ win[op + 256 + lane]);
    if (wantBack && lane < 2) tb = (unsigned)(win[oj - 257 - lane] ^ win[op - 257 - lane]);
    const unsigned ma = __ballot_sync(0xffffffffu, xa != 0);
    const unsigned mta = __ballot_sync(0xffffffffu, ta == 0);       // bit 0: byte 256 equal, bit 1: byte 257 equal
    if (ma) {
        const int src = __ffs(ma) - 1;
        const unsigned long long xs = __shfl_sync(0xffffffffu, xa, src);
        fwd = src * 8 + ((__ffsll((long long)xs) - 1) >> 3);
    } else {
        fwd = 256 + ((mta & 1u) ? ((mta & 2u) ? 2 : 1) : 0);
    }
    lb = 0;
    if (wantBack) {
        const unsigned mb = __ballot_sync(0xffffffffu, xb != 0);
        const unsigned mtb = __ballot_sync(0xffffffffu, tb == 0);
        if (mb) {
            const int src = __ffs(mb) - 1;
            const unsigned long long xs = __shfl_sync(0xffffffffu, xb, src);
            lb = src * 8 + (__clzll((long long)xs) >> 3);
        } else {
            lb = 256 + ((mtb & 1u) ? ((mtb & 2u) ? 2 : 1) : 0);
        }
    }
}

// single-thread backward match length from the window, at most `limit` bytes
__device__ __forceinline__ int back_upto_w(const uint8_t* win, int oj, int op, int limit)
{
    int lb = 0;
    while (lb < limit) {
        const unsigned y = ld4(win, oj - 4 - lb) ^ ld4(win, op - 4 - lb);
        const int c = y ? (__clz(y) >> 3) : 4;
        lb += c;
        if (c < 4) break;
    }
    return lb < limit ? lb : limit;
}

// single-thread forward match length from the window, at most 258 bytes
__device__ __forceinline__ int fwd_upto_w(const uint8_t* win, int oj, int op)
{
    int fwd = 0;
    while (fwd < kMaxMatch) {
        const unsigned x = ld4(win, oj + fwd) ^ ld4(win, op + fwd);
        if (x) { fwd += (__ffs(x) - 1) >> 3; break; }
        fwd += 4;
    }
    return fwd < kMaxMatch ? fwd : kMaxMatch;
}

// candidate of j (raw distance d) with the never-inserted batch starts removed from the hash chain (encoder.cpp:384: a
// FirstPass batch never inserts its first position unless a match of the previous batch covers it)
__device__ int lz_effective_cand(const uint16_t* cand, const LzShared* ps, int j, int d)
{
    for (;;) {
        if (d == 0) return 0;
        const int p = j - d;
        bool hit = false;
        for (int k = 0; k < ps->nexcl; ++k) hit |= (ps->excl[k] == p);
        if (!hit) return d;
        const int dd = p > 0 ? (int)__ldg(cand + p) : 0;
        if (dd == 0) return 0;
        d += dd;
        if (d >= kMaxDistance) return 0;
    }
}

// the arrays of a sub-batch
struct LzSub {
    const uint8_t* info; const unsigned* okbits; const uint16_t* nzw;
    int ntiles, base, B0, s1;
};

// First position the walk takes from state b, or -1 (same result as probe_next).  The info bytes of b+1..b+3 and the
// bitmap words from b+4 on are loaded together, so the common cases cost one shared-memory round trip instead of two:
// this sits on the serial path of the chains.
__device__ __forceinline__ int lz_probe(const LzSub& s, int b)
{
    const int r = b - s.base;
    const unsigned* iw = reinterpret_cast<const unsigned*>(s.info) + ((r + 1) >> 2);
    const unsigned a0 = iw[0], a1 = iw[1];
    const int x = r + 4, w = x >> 5;
    const unsigned k0 = s.okbits[w], k1 = s.okbits[w + 1];                  // okbits holds ntiles + 2 words
    const unsigned near3 = (__funnelshift_r(a0, a1, ((r + 1) & 3) * 8) + 0x007E7D7Cu) & 0x00808080u;   // bytes >= 4 / 3 / 2 (info <= 33: no carry)
    if (near3) return b + 1 + ((__ffs(near3) - 1) >> 3);
    const unsigned bits = __funnelshift_r(k0, k1, x & 31);                 // positions b+4 .. b+35
    if (bits) return b + 4 + __ffs(bits) - 1;
    if (w + 1 >= s.ntiles) return -1;
    const unsigned rest = k1 & (~0u << (x & 31));                          // the part of word w+1 the funnel shift did not cover
    if (rest) return s.base + (w + 1) * 32 + __ffs(rest) - 1;
    if (w + 2 > s.ntiles) return -1;
    const unsigned w2 = s.nzw[w + 2];
    if (w2 == kNone16) return -1;
    return s.base + (int)w2 * 32 + __ffs(s.okbits[w2]) - 1;
}

// succ(b): 0 = no position of the sub-batch is taken from b any more, 1 = the match must be measured exactly, else the next state
__device__ __forceinline__ unsigned lz_succ(const LzSub& s, int b, int& j)
{
    j = -1;
    if ((unsigned)(b - s.B0) >= (unsigned)(s.s1 - s.B0)) return 0u;
    j = lz_probe(s, b);
    if (j < 0) return 0u;
    const int fwd = (int)s.info[j - s.base] - 1;
    return needs_exact(fwd, j - b) ? 1u : (unsigned)(j + fwd);
}

// exact match taken at position j (candidate distance d) from state x, all lanes of the warp cooperate: returns the next state
__device__ __forceinline__ int lz_exact(const uint8_t* win, int wb, int pre, int x, int j, int d, int lane, int& fwdOut, int& lbOut)
{
    const int p = j - d;
    int maxBack = j - x;
    { const int room = p + pre; if (room < maxBack) maxBack = room; }     // R4: clamp at stream start
    if (maxBack > kMaxMatch) maxBack = kMaxMatch;                          // R6: cap (reference breaks at 259)
    int fwd, lb;
    coop_lengths_w(win, wb + j, wb + p, lane, maxBack > 0, fwd, lb);
    if (lb > maxBack) lb = maxBack;
    int m = fwd + lb; if (m > kMaxMatch) m = kMaxMatch;
    fwdOut = fwd; lbOut = lb;
    return j - lb + m;
}

__device__ __forceinline__ unsigned lz_lookup(const unsigned* exList, int exN, int x)
{
    unsigned nb = 0;
    if (exN > kExCap) exN = kExCap;
    for (int q = 0; q < exN; ++q) { const unsigned e = exList[q]; if ((int)(e >> 16) == x) nb = e & 0xFFFFu; }
    return nb;
}

// Adler-32 partial of the chunk (defined with K-CKSUM below).  The chunk went through this kernel's window a moment ago,
// so the pass finds it in L2 instead of HBM, and the separate checksum launch is gone when only Adler-32 is wanted.
__device__ __forceinline__ void lz_checksums(const uint8_t* p, int n, uint32_t* ck, uint8_t* scratch);

__global__ void __launch_bounds__(kLzThreads, 3) k_lz(Job job, int useTma, int useSpec, int useRegion)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* win = smem;
    uint8_t* info = win + kLzWin;
    uint16_t* dist = reinterpret_cast<uint16_t*>(info + kSub + 64);      // candidate distance per position, as the parse sees it
    uint8_t* scratch = reinterpret_cast<uint8_t*>(dist + kSub + 64);
    unsigned* okbits = reinterpret_cast<unsigned*>(scratch);
    unsigned* Sb = okbits + kSubWords;               // states visited by the speculative chains
    unsigned* Tb = Sb + kSubWords;                   // states of the orbit
    unsigned* Lb = Tb + kSubWords;                   // states visited by the links between chains
    uint16_t* nzw = reinterpret_cast<uint16_t*>(Lb + kSubWords);
    uint16_t* stateList = nzw + 272;
    __shared__ LzShared ps;
    __shared__ unsigned wsum[32];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ unsigned exList[kExCap];              // state << 16 | exact next state
    __shared__ unsigned stopKS[kChains];             // kind << 16 | last state of the chain
    __shared__ uint16_t stopT[kChains];              // where the chain left its segment (kind 3)
    __shared__ uint16_t linkArr[kChains];            // state where the link joined the chain, 0xFFFF = it did not
    __shared__ uint16_t segMp[kChains];              // first state of the chain that belongs to the orbit, 0xFFFF = none
    __shared__ unsigned k3Mask[8], lkMask[8], linkedW[8];      // one bit per chain: left its segment / its link joined / entered through the link
    __shared__ unsigned grpMin[kSubWords / 32 + 2];
    __shared__ unsigned hist4[kHistCopies * kHistStride];     // literal/length + distance histograms (private copies)
    __shared__ unsigned rcov[kRegionWords];                   // region: bit per position, covered by a match
    __shared__ uint16_t rbase[kRegionWords];                  // region: literals before each 32-position word

#ifdef ZZ_PHASE_TIMING
    __shared__ unsigned long long phaseAcc[12];
    __shared__ long long phaseT;
    if (threadIdx.x == 0) { for (int k = 0; k < 12; ++k) phaseAcc[k] = 0; phaseT = clock64(); }
#endif
    const unsigned slot = blockIdx.x;
    const Geom g = chunk_geom(job, slot);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    const uint8_t* chunk0 = job.src + g.off;
    const Stream strm = { job.src - job.history, job.src + job.n };
    const uint16_t* cand = job.cand + (size_t)slot * job.chunk;
    uint8_t* lits = job.info + (size_t)slot * job.chunk;           // the block's literal bytes in order
    unsigned* myh = hist4 + (warp % kHistCopies) * kHistStride;
    for (int i = tid; i < kHistCopies * kHistStride; i += kLzThreads) hist4[i] = 0;
    uint32_t* tokA = job.tokA + (size_t)slot * kMaxTokens;
    uint16_t* tokD = job.tokD + (size_t)slot * kMaxTokens;
    const int phase = (int)(reinterpret_cast<uintptr_t>(chunk0) & 15);

    if (tid == 0) {
        ps.pos = 0; ps.ntok = 0; ps.nexcl = 0; ps.npend = 0; ps.exN = 0; ps.err = 0;
        ps.histPos = 0; ps.histTok = 0; ps.regEnd = 0; ps.regTok = 0; ps.histStop = 0; ps.nlits = 0; ps.regLits = 0;
        for (int w = 0; w < 8; ++w) { k3Mask[w] = 0; lkMask[w] = 0; }
        mbar_init(&mbar, 1);
    }
    __syncthreads();
    unsigned parity = 0;

    const int t0 = g.t0;
    // ---- batches of the reference's WriteBlock2Pass loop (encoder.cpp:225-234) ----
    for (;;) {
        const int pos = ps.pos;
        if (pos >= t0) break;
        int E = pos + kBatch; if (E > t0) E = t0;
        const int B0 = pos + 1;                          // FirstPass: backRefEnd = j = startPos + 1
        const int nexcl = ps.nexcl;
        const int ex0 = nexcl > 0 ? ps.excl[0] : -1, ex1 = nexcl > 1 ? ps.excl[1] : -1, ex2 = nexcl > 2 ? ps.excl[2] : -1;
        if (tid == 0) ps.b = B0;
        __syncthreads();

        // ---- sub-batches ----
        bool firstSub = true;
        for (int s0 = B0; s0 < E; ) {
            int s1n = s0 + kSub; if (s1n > E) s1n = E;
            if (!firstSub && ps.b >= s1n) { s0 = s1n; continue; }     // a long match jumped over the whole sub-batch
            // ---- window of the sub-batch: position i lives at win[wb + i]; shared and global 16-byte phases agree ----
            const int off0 = s0 - kLzBack;
            const int wb = 16 + ((phase + off0) & 15) - off0;
            int lo = off0; if (lo < -g.pre) lo = -g.pre;
            int hi = s1n + kLzAhead - 16; if (hi > g.n) hi = g.n;
            bool tmaIssued = false;
            {
                const int s_lo = wb + lo, s_hi = wb + hi;
                const int a_lo = (s_lo + 15) & ~15, a_hi = s_hi & ~15;
                tmaIssued = useTma && a_lo < a_hi;
                if (tmaIssued) {
                    if (tid == 0) {
                        // the window was read through the generic proxy by the previous sub-batch: order those accesses before
                        // the asynchronous writes
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        const unsigned bytes = (unsigned)(a_hi - a_lo);
                        mbar_expect_tx(&mbar, bytes);
                        const uint8_t* gsrc = chunk0 + (a_lo - wb);
                        for (unsigned o = 0; o < bytes; o += 16384u) {
                            const unsigned len = bytes - o < 16384u ? bytes - o : 16384u;
                            bulk_g2s(win + a_lo + o, gsrc + o, len, &mbar);
                        }
                    }
                    for (int q = s_lo + tid; q < a_lo; q += kLzThreads) win[q] = chunk0[q - wb];
                    for (int q = a_hi + tid; q < s_hi; q += kLzThreads) win[q] = chunk0[q - wb];
                    for (int q = s_hi + tid; q < s_hi + 16; q += kLzThreads) win[q] = 0;
                } else {
                    load_window(win, wb, chunk0, lo, hi, hi + 16);
                }
            }
            if (tmaIssued) {
                unsigned spins = 0;
                while (!mbar_try_wait(&mbar, parity)) { if (++spins > (1u << 18)) { ps.err = 1; break; } }
                parity ^= 1u;
            }
            __syncthreads();
            PT(0);
            if (firstSub) {
                // ---- first probe of the batch: j == backRefEnd, no backward room (encoder.cpp:384-386) ----
                firstSub = false;
                if (warp == 0) {
                    int b = B0;
                    const int d = lz_effective_cand(cand, &ps, B0, (int)__ldg(cand + B0));     // the excluded start may be B0's own candidate
                    const int inf = info_of_w(win, wb + B0, d, B0 - d + g.pre);
                    if (inf >= 5) {
                        int fwd = inf - 1, lbNone;
                        if (fwd >= kCapLen) coop_lengths_w(win, wb + B0, wb + B0 - d, lane, false, fwd, lbNone);
                        const int tk = ps.ntok;
                        if (lane == 0) { tokA[tk] = (uint32_t)B0 | ((uint32_t)fwd << 16); tokD[tk] = (uint16_t)d; ps.ntok = tk + 1; }
                        b = B0 + fwd;
                    }
                    if (lane == 0) ps.b = b;
                }
                __syncthreads();
                if (ps.b >= s1n) { __syncthreads(); s0 = s1n; continue; }
            }
            PT(1);
            const int bIn = ps.b;
            int lowb = bIn > s0 - 64 ? bIn : s0 - 64; if (lowb > s0) lowb = s0;
            const int base = lowb & ~31;
            int s1 = base + kSub; if (s1 > s1n) s1 = s1n;           // the arrays hold kSub states from the tile-aligned base
            const int ntiles = (s1 - base + 31) >> 5;
            const int lim = ntiles * 32;

            // ---- A: match info of the positions [s0, s1); everything else in the array reads as "no match" ----
            {
                unsigned* queue = reinterpret_cast<unsigned*>(scratch) + warp * kLzQueue;
                int queued = 0;
                const unsigned* w32 = reinterpret_cast<const unsigned*>(win);
                const int iters = (lim + 4 * kLzThreads - 1) / (4 * kLzThreads);
                uint2 ddNext = make_uint2(0u, 0u);
                {
                    const int j0 = base + 4 * tid;
                    if (j0 + 3 >= s0 && j0 < s1) ddNext = __ldg(reinterpret_cast<const uint2*>(cand + j0));
                }
                for (int it = 0; it < iters; ++it) {
                    const int idx = 4 * tid + it * 4 * kLzThreads;
                    const int j0 = base + idx;
                    const bool inArr = idx < lim;
                    const bool live = inArr && j0 + 3 >= s0 && j0 < s1;
                    const uint2 dd = ddNext;
                    ddNext = make_uint2(0u, 0u);
                    {
                        const int jn = j0 + 4 * kLzThreads;
                        if (idx + 4 * kLzThreads < lim && jn + 3 >= s0 && jn < s1) ddNext = __ldg(reinterpret_cast<const uint2*>(cand + jn));
                    }
                    int d[4] = { (int)(dd.x & 0xFFFFu), (int)(dd.x >> 16), (int)(dd.y & 0xFFFFu), (int)(dd.y >> 16) };
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (!live || j0 + k < s0 || j0 + k >= s1) d[k] = 0;
                    if (nexcl) {
                        // a position whose candidate is a never-inserted batch start must look one entry further down the chain:
                        // noted here, recomputed after the pass (at most three such positions per chunk)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int p = j0 + k - d[k];
                            if (d[k] != 0 && (p == ex0 || p == ex1 || p == ex2)) { const int q = atomicAdd(&ps.npend, 1); if (q < 8) ps.pendJ[q] = j0 + k; }
                        }
                    }
                    if (inArr) *reinterpret_cast<uint2*>(dist + idx) = make_uint2((unsigned)d[0] | ((unsigned)d[1] << 16), (unsigned)d[2] | ((unsigned)d[3] << 16));
                    const int oj0 = wb + (live ? j0 : s0);
                    const unsigned* wj = w32 + (oj0 >> 2);
                    unsigned packed = 0, longMask = 0;
                    int op[4] = { 0, 0, 0, 0 };
                    if (live) {
                        const unsigned W0 = wj[-1], W1 = wj[0], W2 = wj[1], W3 = wj[2];
                        const int ph = (oj0 & 3) * 8;
                        const unsigned V0 = __funnelshift_r(W0, W1, ph), V1 = __funnelshift_r(W1, W2, ph), V2 = __funnelshift_r(W2, W3, ph);
                        unsigned pw0[4], pw1[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            op[k] = oj0 + k - d[k];
                            const unsigned* wp = w32 + (op[k] >> 2);
                            pw0[k] = wp[0]; pw1[k] = wp[1];
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const unsigned x = __funnelshift_r(V1, V2, 8 * k) ^ __funnelshift_r(pw0[k], pw1[k], op[k] * 8);
                            if (d[k] != 0) {
                                if (x == 0) longMask |= 1u << k;         // >= 4 bytes forwards: usable whatever lies behind; length from the queue
                                else {
                                    const int fwd = (__ffs(x) - 1) >> 3;
                                    const unsigned y = __funnelshift_r(V0, V1, 8 * k) ^ __funnelshift_r(w32[(op[k] >> 2) - 1], pw0[k], op[k] * 8);
                                    int back = y ? (__clz(y) >> 3) : 4;
                                    const int room = j0 + k - d[k] + g.pre;   // bytes of real history before the candidate (R4 clamp)
                                    if (back > room) back = room;
                                    if (fwd + back >= 4) packed |= (unsigned)(fwd + 1) << (8 * k);
                                }
                            }
                        }
                    }
                    // queued bytes are stored as 0 here and overwritten when the queue is drained: the drain comes after a
                    // __syncwarp(), which orders the two stores of the warp
                    if (inArr) *reinterpret_cast<unsigned*>(info + idx) = packed;
                    {   // append: a lane queues 0..4 positions; its slot = positions queued by lower lanes (three votes on the count's bits)
                        const unsigned cnt = (unsigned)__popc(longMask);
                        const unsigned b0 = __ballot_sync(0xffffffffu, cnt & 1u), b1 = __ballot_sync(0xffffffffu, cnt & 2u), b2 = __ballot_sync(0xffffffffu, cnt & 4u);
                        int slotq = queued + __popc(b0 & ltMask) + 2 * __popc(b1 & ltMask) + 4 * __popc(b2 & ltMask);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if ((longMask >> k) & 1u) queue[slotq++] = (unsigned)(oj0 + k) | ((unsigned)op[k] << 16);
                        queued += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                    }
                    __syncwarp();
                    int head = 0;
                    while (queued - head >= 32) {
                        const unsigned e = queue[head + lane]; const int oj = (int)(e & 0xFFFFu), opq = (int)(e >> 16);
                        info[oj - wb - base] = (uint8_t)(4 + fwd_more(win, oj + 4, opq + 4) + 1);
                        head += 32;
                    }
                    if (head) {                                      // keep the remainder (< 32 entries) at the front
                        const int rem = queued - head;
                        unsigned e = 0;
                        if (lane < rem) e = queue[head + lane];
                        __syncwarp();
                        if (lane < rem) queue[lane] = e;
                        queued = rem;
                    }
                    __syncwarp();
                }
                if (lane < queued) {
                    const unsigned e = queue[lane];
                    const int oj = (int)(e & 0xFFFFu), opq = (int)(e >> 16);
                    info[oj - wb - base] = (uint8_t)(4 + fwd_more(win, oj + 4, opq + 4) + 1);
                }
                if (tid < 16) *reinterpret_cast<unsigned*>(info + lim + 4 * tid) = 0u;      // look-ahead of the last states
            }
            __syncthreads();
            PT(2);
            // the queues are drained: their space now holds the bitmaps, nzw and the state list
            for (int t = tid; t < kSubWords; t += kLzThreads) { Sb[t] = 0; Tb[t] = 0; Lb[t] = 0; }
            if (tid < kChains) segMp[tid] = 0xFFFFu;
            if (tid < 8) linkedW[tid] = 0;
            if (tid < ps.npend && tid < 8) {                      // positions whose candidate changes with the parse
                const int pj = ps.pendJ[tid];
                const int pd = lz_effective_cand(cand, &ps, pj, (int)dist[pj - base]);
                dist[pj - base] = (uint16_t)pd;
                info[pj - base] = (uint8_t)info_of_w(win, wb + pj, pd, pj - pd + g.pre);
            }
            __syncthreads();
            if (tid == 0) { ps.npend = 0; ps.exN = 0; }
            for (int idx = tid; idx < lim + 64; idx += kLzThreads) {
                const unsigned m4 = __ballot_sync(0xffffffffu, info[idx] != 0);
                if (lane == 0) okbits[idx >> 5] = m4;
            }
            __syncthreads();
            {   // nzw[w] = next non-empty bitmap word at or after w (suffix minimum: within groups of 32 words by shuffles, then across)
                const int w = tid;
                unsigned v = (w < ntiles && okbits[w] != 0) ? (unsigned)w : kNone16;
                if (w < kSubWords + 27) {
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32 && u < v) v = u; }
                    if (lane == 0) grpMin[warp] = v;
                }
                __syncthreads();
                if (w <= ntiles) {
                    if (v == kNone16) for (int gq = warp + 1; gq <= (ntiles >> 5) && v == kNone16; ++gq) v = grpMin[gq];
                    nzw[w] = (uint16_t)v;
                }
            }
            __syncthreads();

            // the region that is final by now (everything below the walk's state) and not yet counted: done beside this walk
            const int rA = ps.histPos, rB = ps.regEnd, tA = ps.histTok, tB = ps.regTok;
            const unsigned nlit0 = ps.nlits;
            const int regSpan = rB - (rA & ~31);
            const bool regionNow = useRegion && !ps.histStop && rB > rA && regSpan <= kRegionMax;
            PT(3);

            // ---- C: the orbit of succ from the entry state ----
            if (warp < kChaseWarps) {
                __syncwarp();
                const LzSub sub = { info, okbits, nzw, ntiles, base, B0, s1 };
                // speculative chains: lane i follows succ from the first state of its segment and marks what it visits
                const int ci = tid;
                const int segLo = base + kSegStates * ci;
                const int segHi = ci == kChains - 1 ? base + lim : segLo + kSegStates;
                int kind = 0, state = 0, tgt = 0;                   // 1: succ == 0, 2: long match, 3: left the segment (tgt = next state)
                {
                    int x = segLo;
                    bool go = useSpec && segLo < s1;
                    bool givenUp = false;
                    int guardR = 0; (void)guardR;
                    for (;;) {
                        if (go) {
                            int guard1 = 0;
                            for (;;) {
                                LZ_GUARD(guard1, kSub, 1, x, base)
                                LZ_CHECK(x >= base && x < base + lim, 2, x, base);
                                atomicOr(&Sb[(x - base) >> 5], 1u << (x & 31));
                                int j;
                                const unsigned f = lz_succ(sub, x, j);
                                if (f == 0u) { kind = 1; state = x; break; }
                                if (f == 1u) { kind = 2; state = x; break; }
                                if ((int)f >= segHi || (int)f >= s1) { kind = 3; state = x; tgt = (int)f; break; }
                                x = (int)f;
                            }
                            go = false;
                        }
                        // long matches that stopped chains: measured here while they are few (text), left to the true walk otherwise
                        unsigned brk = __ballot_sync(0xffffffffu, kind == 2 && !givenUp);
                        if (!brk) break;
                        LZ_GUARD(guardR, 4 * kSegStates, 3, x, state)
                        if (__popc(brk) > 8) { givenUp = true; break; }
                        int myJ = 0, myD = 0;
                        if ((brk >> lane) & 1u) { myJ = lz_probe(sub, state); myD = (int)dist[myJ - base]; }
                        while (brk) {
                            const int L = __ffs(brk) - 1; brk &= brk - 1;
                            const int xs = __shfl_sync(0xffffffffu, state, L), js = __shfl_sync(0xffffffffu, myJ, L), ds = __shfl_sync(0xffffffffu, myD, L);
                            int fw, lb;
                            const int nb = lz_exact(win, wb, g.pre, xs, js, ds, lane, fw, lb);
                            if (lane == L) {
                                const int q = atomicAdd(&ps.exN, 1);
                                if (q < kExCap) exList[q] = ((unsigned)xs << 16) | (unsigned)(nb < 65535 ? nb : 65535);
                                if (nb >= segHi || nb >= s1) { kind = 3; tgt = nb; }
                                else { kind = 0; x = nb; go = true; }
                            }
                        }
                    }
                }
                stopKS[ci] = ((unsigned)kind << 16) | (unsigned)state;
                stopT[ci] = (uint16_t)(tgt < 65535 ? tgt : 65535);
                chase_barrier();
                PT(4);
                // links: lane i follows succ from the state where lane i-1's chain entered segment i until it steps on its own
                // chain.  If lane i-1's chain turns out to be the orbit, so is lane i's from that state on.
                int linkMp = -1;
                if (ci > 0 && kind != 0) {
                    const unsigned pks = stopKS[ci - 1];
                    const int pt = stopT[ci - 1];
                    if ((pks >> 16) == 3u && pt >= segLo && pt < segHi && pt < s1) {
                        const int exN = ps.exN;
                        int x = pt;
                        int guard2 = 0; (void)guard2;
                        for (;;) {
                            LZ_GUARD(guard2, kSegStates + 2, 4, x, segLo)
                            if ((Sb[(x - base) >> 5] >> (x & 31)) & 1u) { linkMp = x; break; }
                            atomicOr(&Lb[(x - base) >> 5], 1u << (x & 31));
                            int j;
                            unsigned f = lz_succ(sub, x, j);
                            if (f == 1u) f = lz_lookup(exList, exN, x);
                            if (f == 0u || (int)f >= segHi || (int)f >= s1) break;
                            x = (int)f;
                        }
                    }
                }
                linkArr[ci] = (uint16_t)(linkMp >= 0 ? linkMp : 0xFFFF);
                {
                    const unsigned m1 = __ballot_sync(0xffffffffu, kind == 3), m2 = __ballot_sync(0xffffffffu, linkMp >= 0);
                    if (lane == 0) { k3Mask[warp] = m1; lkMask[warp] = m2; }
                }
                chase_barrier();
                PT(5);
                if (warp == 0) {
                    __syncwarp();
                    int cur = bIn, npre = 0;
                    const int tokBase = ps.ntok;
                    bool alive = true;

                    if (cur < base) {
                        // entry state far behind the arrays: every usable position of the sub-batch is far enough, the first one is taken
                        const unsigned w0 = nzw[0];
                        if (w0 == kNone16) alive = false;                 // nothing to take: the state stays
                        else {
                            const int j = base + (int)w0 * 32 + __ffs(okbits[w0]) - 1;
                            const int d = dist[j - base];
                            int fw, lb;
                            const int nb = lz_exact(win, wb, g.pre, cur, j, d, lane, fw, lb);
                            int m = fw + lb; if (m > kMaxMatch) m = kMaxMatch;
                            if (lane == 0) { tokA[tokBase] = (uint32_t)(j - lb) | ((uint32_t)m << 16); tokD[tokBase] = (uint16_t)d; }
                            npre = 1;
                            cur = nb;
                        }
                    }

                    int endsAt = -1;                                       // orbit state with succ == 0 that ended the walk (it yields no token)
                    if (alive && cur < s1) {
                        // bit t of okRun: segment t is entered through its link provided segment t-1's chain is the orbit and leaves into it
                        unsigned okW[kChaseWarps];
#pragma unroll
                        for (int w = 0; w < kChaseWarps; ++w) okW[w] = ((k3Mask[w] << 1) | (w ? k3Mask[w - 1] >> 31 : 0u)) & lkMask[w];
                        int guard3 = 0; (void)guard3;
                        for (;;) {
                            if (cur >= s1) break;
                            // The lanes of this warp all carry the walk's state, but nothing makes them run in lockstep (a warp
                            // that split at an `if (lane == 0)` or in the links may stay split: seen on the hardware, where one
                            // group of lanes then missed the join another group had just recorded).  The walk is written so that
                            // every group takes the same path whatever the skew: what it writes is idempotent (bitmap ORs, the
                            // same values into segMp), and the one thing it reads back, segMp[i], counts as "not joined yet"
                            // also when it already holds this very state (a faster group of the same warp put it there).
                            LZ_GUARD(guard3, 2 * kSub, 5, cur, base)
                            LZ_CHECK(cur >= base, 6, cur, base);
                            const int r = cur - base;
                            int j;
                            const unsigned f0 = lz_succ(sub, cur, j);
                            const bool marked = (Sb[r >> 5] >> (r & 31)) & 1u;
                            int x = cur;
                            bool joined = false;
                            if (marked) {
                                int i = r / kSegStates; if (i > kChains - 1) i = kChains - 1;
                                const unsigned iKS = stopKS[i];
                                const unsigned mpNow = segMp[i];
                                if ((iKS >> 16) != 0u && (mpNow == 0xFFFFu || mpNow == (unsigned)cur)) {
                                    // the walk stepped on chain i: the chain is the orbit from here on, and so are the chains of the
                                    // following segments as long as each one's link joined it
                                    joined = true;
                                    // run of ones in okRun from bit i+1 on
                                    int last = i;
                                    if (i + 1 < kChains) {
                                        int t = i + 1;                                // ends as the first index that is not in the run
                                        bool stop = false;
#pragma unroll
                                        for (int w = 0; w < kChaseWarps; ++w) {
                                            if (!stop && (t >> 5) == w) {
                                                const int sh = t & 31, room = 32 - sh;
                                                const unsigned inv = ~(okW[w] >> sh);         // the sh bits shifted in at the top read as "end of word"
                                                const int z = inv ? __ffs(inv) - 1 : 32;
                                                if (z < room) { t += z; stop = true; } else t += room;
                                            }
                                        }
                                        last = t - 1;
                                    }
                                    if (lane == 0) segMp[i] = (uint16_t)cur;
                                    for (int t = i + 1 + lane; t <= last; t += 32) { segMp[t] = linkArr[t]; atomicOr(&linkedW[t >> 5], 1u << (t & 31)); }
                                    __syncwarp();
                                    const unsigned lKS = stopKS[last];
                                    const int lKind = (int)(lKS >> 16), lState = (int)(lKS & 0xFFFFu), lTgt = stopT[last];
                                    if (lKind == 1) { endsAt = lState; cur = lState; break; }
                                    if (lKind == 3) { cur = lTgt; continue; }
                                    x = lState;                            // long match at the end of the chain: measured below
                                    j = lz_probe(sub, x);
                                }
                            }
                            if (!joined) {
                                if (lane == 0) atomicOr(&Tb[r >> 5], 1u << (r & 31));
                                if (f0 == 0u) { endsAt = x; break; }
                                if (f0 != 1u) { cur = (int)f0; continue; }
                            }
                            // long match at state x (taken at position j): exact lengths, unless a chain has measured it already
                            {
                                const unsigned known = __shfl_sync(0xffffffffu, lz_lookup(exList, ps.exN, x), 0);     // lane 0's view: uniform
                                if (known) { cur = (int)known; continue; }
                            }
                            const int d = dist[j - base];
                            int fwd, lb;
                            int nb = lz_exact(win, wb, g.pre, x, j, d, lane, fwd, lb);
                            int exBase = 0;
                            if (lane == 0) { exBase = atomicAdd(&ps.exN, 1); if (exBase < kExCap) exList[exBase] = ((unsigned)x << 16) | (unsigned)(nb < 65535 ? nb : 65535); }
                            // Runs (RLE-like data): a full-length match whose successor states repeat it 258 bytes further on.  The
                            // equality between the two sides is measured once beyond the first 258 bytes (R, as far as 32 more matches
                            // can use it and the window reaches) and lane i-1 checks that state x + 258 i is an unmeasured long state
                            // whose probe position has the same gap and the same distance.  For those states fwd_i = min(258, R - 258 i)
                            // and the backward part is the gap again (its bytes lie inside [j, j + R)), so match i is
                            // (start x_i, length 258) exactly as the walk would find it, as long as R - 258 i >= 258 - gap.
                            const int gap = j - x;
                            const int p = j - d;
                            if (fwd == kMaxMatch && lb == gap && nb + 1 < s1) {
                                const int xi = x + kMaxMatch * (lane + 1);
                                bool ok = xi < s1;
                                int ji = -1;
                                if (ok) ok = lz_succ(sub, xi, ji) == 1u;
                                if (ok) ok = ji == xi + gap;
                                if (ok) ok = (int)dist[xi + gap - base] == d;
                                const unsigned okm = __ballot_sync(0xffffffffu, ok);
                                int K = okm == 0xffffffffu ? 32 : __ffs(~okm) - 1;            // candidates of the run
                                if (K > 0) {
                                    // R: equal bytes between the two sides from j on, measured only as far as K matches need it
                                    int Rcap = kMaxMatch * (K + 1) - gap;
                                    { const int limB = g.body - j; if (limB < Rcap) Rcap = limB; }
                                    { const int limW = hi - 272 - j; if (limW < Rcap) Rcap = limW; }        // what the window holds (a step reads 264 bytes beyond its offset)
                                    int R = Rcap;
                                    for (int off = 256; off < Rcap && R == Rcap; off += 1024) {
                                        unsigned long long xa[4];
#pragma unroll
                                        for (int u = 0; u < 4; ++u)
                                            xa[u] = off + 256 * u < Rcap ? ld8(win, wb + j + off + 256 * u + lane * 8) ^ ld8(win, wb + p + off + 256 * u + lane * 8) : 0ull;
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            const unsigned mm = __ballot_sync(0xffffffffu, xa[u] != 0);
                                            if (mm && R == Rcap) {
                                                const int src = __ffs(mm) - 1;
                                                const unsigned long long xs = __shfl_sync(0xffffffffu, xa[u], src);
                                                const int rr = off + 256 * u + src * 8 + ((__ffsll((long long)xs) - 1) >> 3);
                                                if (rr < R) R = rr;
                                            }
                                        }
                                    }
                                    // match i needs R - 258 i >= 258 - gap
                                    const int byR = (R - kMaxMatch + gap) / kMaxMatch;
                                    if (byR < K) K = byR < 0 ? 0 : byR;
                                    int runBase = 0;
                                    if (lane == 0 && K > 0) runBase = atomicAdd(&ps.exN, K);
                                    runBase = __shfl_sync(0xffffffffu, runBase, 0);
                                    if (lane < K) {
                                        const int nbi = xi + kMaxMatch;
                                        if (runBase + lane < kExCap) exList[runBase + lane] = ((unsigned)xi << 16) | (unsigned)(nbi < 65535 ? nbi : 65535);
                                        atomicOr(&Tb[(xi - base) >> 5], 1u << (xi & 31));
                                    }
                                    nb = x + kMaxMatch * (K + 1);
                                    __syncwarp();
                                }
                            }
                            cur = nb;
                        }
                        __syncwarp();
                    }
                    if (lane == 0) { ps.b = cur; ps.npre = npre; ps.endsAt = endsAt; }
                    PT(6);
                }
            } else if (regionNow) {
                // ---- beside the walk: GetFrequencies (encoder.cpp:442-471) for the region the previous sub-batches finished.  Its
                //      matches are counted token-parallel and mark the positions they cover; the literals in between are counted
                //      position-parallel from the window and written in order to the literal stream for K-EMIT ----
                const int wt = tid - kChaseWarps * 32;
                const int r0 = rA & ~31;
                const int words = (rB - r0 + 31) >> 5;
                LZ_CHECK(words <= kRegionWords && tB >= tA && wb + r0 >= 0 && wb + rB + 8 < kLzWin, 10, rA, rB);
                for (int w = wt; w < words; w += kWorkerThreads) rcov[w] = 0;
                worker_barrier();
                for (int k = tA + wt; k < tB; k += kWorkerThreads) {
                    const uint32_t t = tokA[k];
                    const int ms = (int)(t & 0xFFFF) - r0, ln = (int)(t >> 16), me = ms + ln - 1;
                    int eb, ev;
                    atomicAdd(&myh[len_symbol(ln, eb, ev)], 1u);
                    atomicAdd(&myh[286 + dist_symbol(tokD[k], eb, ev)], 1u);
                    for (int w = ms >> 5; w <= (me >> 5); ++w) {
                        unsigned m = 0xffffffffu;
                        if (w == (ms >> 5)) m &= 0xffffffffu << (ms & 31);
                        if (w == (me >> 5)) m &= 0xffffffffu >> (31 - (me & 31));
                        atomicOr(&rcov[w], m);
                    }
                }
                if (wt == 0) {                                       // positions outside [rA, rB) are not this region's literals
                    if (rA & 31) atomicOr(&rcov[0], (1u << (rA & 31)) - 1u);
                    if ((rB - r0) & 31) atomicOr(&rcov[words - 1], 0xffffffffu << ((rB - r0) & 31));
                }
                worker_barrier();
                if (warp == kChaseWarps) {                           // exclusive scan of the literal counts per word
                    constexpr int kPer = (kRegionWords + 31) / 32;
                    unsigned c[kPer], sum = 0;
#pragma unroll
                    for (int k = 0; k < kPer; ++k) { const int w = lane * kPer + k; c[k] = w < words ? (unsigned)__popc(~rcov[w]) : 0u; sum += c[k]; }
                    unsigned inc = sum;
                    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                    unsigned before = inc - sum;
#pragma unroll
                    for (int k = 0; k < kPer; ++k) { const int w = lane * kPer + k; if (w < words) rbase[w] = (uint16_t)before; before += c[k]; }
                    if (lane == 31) ps.regLits = inc;
                }
                worker_barrier();
                for (int q = wt * 4; q < words * 32; q += kWorkerThreads * 4) {
                    const unsigned cword = rcov[q >> 5];
                    const unsigned cw = (cword >> (q & 31)) & 0xFu;
                    if (cw == 0xFu) continue;                        // four covered positions: nothing to count
                    const unsigned v = ld4(win, wb + r0 + q);
                    unsigned rank = nlit0 + rbase[q >> 5] + (unsigned)__popc(~cword & ((1u << (q & 31)) - 1u));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (!((cw >> k) & 1u)) { const unsigned b = (v >> (8 * k)) & 0xFFu; atomicAdd(&myh[b], 1u); LZ_CHECK(rank < job.chunk, 9, rank, q); lits[rank++] = (uint8_t)b; }
                }
            }
            __syncthreads();
            // (no divergent statement may sit between a CTA barrier and the chase warps' named barriers: the bookkeeping is here)
            if (tid == 0) {
                if (regionNow) { ps.histPos = rB; ps.histTok = tB; ps.nlits = nlit0 + ps.regLits; }
                else if (regSpan > kRegionMax) ps.histStop = 1;        // too much at once: left to the end
            }
            PT(7);

            // ---- D: the orbit's states, compacted, then expanded to tokens in parallel ----
            {
                // states of the orbit in this thread's 32-state word: what the true walk marked, plus every joined chain from the
                // state where the orbit joined it, plus the link that led there
                unsigned word = 0;
                if (tid < ntiles) {
                    word = Tb[tid];
                    int t = (tid * 32) / kSegStates; if (t > kChains - 1) t = kChains - 1;
                    const unsigned mp = segMp[t];
                    if (mp != 0xFFFFu) {
                        const int wm = ((int)mp - base) >> 5;
                        if (tid > wm) word |= Sb[tid];
                        else if (tid == wm) word |= Sb[tid] & (0xffffffffu << (mp & 31u));
                        if (tid <= wm && ((linkedW[t >> 5] >> (t & 31)) & 1u)) word |= Lb[tid];
                    }
                    const int endsAt = ps.endsAt;
                    if (endsAt >= 0 && ((endsAt - base) >> 5) == tid) word &= ~(1u << (endsAt & 31));
                }
                const unsigned cnt = (unsigned)__popc(word);
                unsigned inc = cnt;
                for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                if (lane == 31) wsum[warp] = inc;
                __syncthreads();
                if (warp == 0) {
                    unsigned v = lane < kLzWarps ? wsum[lane] : 0u;
                    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
                    wsum[lane] = v;
                }
                __syncthreads();
                int out = (int)((warp ? wsum[warp - 1] : 0u) + inc - cnt);
                unsigned wbits = word;
                while (wbits) {
                    const int bit = __ffs(wbits) - 1; wbits &= wbits - 1;
                    LZ_CHECK(out < kSub / 4 + 40, 7, out, base);
                    stateList[out++] = (uint16_t)(base + tid * 32 + bit);
                }
                __syncthreads();
                const int nbt = (int)wsum[31];
                const int outBase = ps.ntok + ps.npre;
                for (int t = tid; t < nbt; t += kLzThreads) {
                    const int x = stateList[t];
                    const int j = probe_next(info, okbits, nzw, ntiles, base, x);
                    const int d = dist[j - base];
                    int fwd = (int)info[j - base] - 1;
                    int limit = j - x;                                     // pending literals (encoder.cpp:404)
                    { const int room = j - d + g.pre; if (room < limit) limit = room; }
                    if (limit > kMaxMatch) limit = kMaxMatch;
                    const int lb = back_upto_w(win, wb + j, wb + j - d, limit);
                    if (fwd >= kCapLen) fwd = fwd_upto_w(win, wb + j, wb + j - d);      // the compare of phase A stopped at 32 bytes
                    int m = fwd + lb;
                    if (m > kMaxMatch) m = kMaxMatch;
                    LZ_CHECK(outBase + t < kMaxTokens && j >= 0, 8, outBase + t, j);
                    tokA[outBase + t] = (uint32_t)(j - lb) | ((uint32_t)m << 16);
                    tokD[outBase + t] = (uint16_t)d;
                }
                __syncthreads();
                if (tid == 0) { ps.ntok = outBase + nbt; ps.regEnd = ps.b; ps.regTok = outBase + nbt; }
            }
            __syncthreads();
            s0 = s1;
            PT(8);
        }
        if (tid == 0) {
            const int finalB = ps.b;
            const int newpos = finalB > E ? finalB : E;
            if (finalB < E && newpos < t0 && ps.nexcl < 4) ps.excl[ps.nexcl++] = newpos;     // next batch start not covered by a match: never inserted
            ps.pos = newpos;
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0 && ps.err) atomicOr(reinterpret_cast<unsigned long long*>(&job.total[1]), 8ull);

    // ---- the rest of GetFrequencies (encoder.cpp:442-471): what the walks left uncounted (the last sub-batch's matches, the
    //      literals behind histPos including the block's tail, or everything if a region grew too large to be counted beside
    //      a walk).  Same scheme over the whole chunk: matches mark a coverage bitmap, literals are counted position-parallel
    //      (coalesced reads from global memory).  The batch arrays are dead: the space is reused. ----
    unsigned* cov = reinterpret_cast<unsigned*>(smem);             // bit per position: covered by a match (or counted already)
    unsigned* litBase = cov + kMaxChunk / 32;                      // literals before each 32-position word
    for (int i = tid; i < kMaxChunk / 32; i += kLzThreads) cov[i] = 0;
    __syncthreads();
    {
        const int ntok = ps.ntok, hp = ps.histPos, ht = ps.histTok;
        const unsigned nl0 = ps.nlits;
        for (int k = ht + tid; k < ntok; k += kLzThreads) {
            const uint32_t t = tokA[k];
            const int ms = (int)(t & 0xFFFF), ln = (int)(t >> 16), me = ms + ln - 1;
            int eb, ev;
            atomicAdd(&myh[len_symbol(ln, eb, ev)], 1u);
            atomicAdd(&myh[286 + dist_symbol(tokD[k], eb, ev)], 1u);
            for (int w = ms >> 5; w <= (me >> 5); ++w) {
                unsigned m = 0xffffffffu;
                if (w == (ms >> 5)) m &= 0xffffffffu << (ms & 31);
                if (w == (me >> 5)) m &= 0xffffffffu >> (31 - (me & 31));
                atomicOr(&cov[w], m);
            }
        }
        __syncthreads();
        // positions counted already, and positions at or beyond the block's end, are not literals here
        for (int w = tid; w < kMaxChunk / 32; w += kLzThreads) {
            const int lo = w * 32;
            unsigned m = 0;
            if (lo < hp) m |= lo + 32 <= hp ? 0xffffffffu : ((1u << (hp - lo)) - 1u);
            if (lo + 32 > g.body) m |= lo >= g.body ? 0xffffffffu : (0xffffffffu << (g.body - lo));
            if (m) cov[w] |= m;
        }
        __syncthreads();
        {   // exclusive scan of the literal counts per word (kWordsPerThread consecutive words per thread)
            constexpr int kWordsPerThread = (kMaxChunk / 32 + kLzThreads - 1) / kLzThreads;
            const int w0 = tid * kWordsPerThread;
            unsigned c[kWordsPerThread], sum = 0;
#pragma unroll
            for (int k = 0; k < kWordsPerThread; ++k) { c[k] = w0 + k < kMaxChunk / 32 ? (unsigned)__popc(~cov[w0 + k]) : 0u; sum += c[k]; }
            unsigned inc = sum;
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                unsigned v = lane < kLzWarps ? wsum[lane] : 0u;
                for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
                wsum[lane] = v;
            }
            __syncthreads();
            unsigned before = nl0 + (warp ? wsum[warp - 1] : 0u) + inc - sum;
#pragma unroll
            for (int k = 0; k < kWordsPerThread; ++k) if (w0 + k < kMaxChunk / 32) { litBase[w0 + k] = before; before += c[k]; }
            if (tid == kLzThreads - 1) job.state[slot].nlit = before;
        }
        __syncthreads();
        // literals: histogram, and the literal bytes written out in order for K-EMIT (position -> rank through the
        // coverage bitmap).  A 4-byte aligned chunk reads whole words (a word that holds a valid byte never leaves its page).
        const bool srcAligned = (reinterpret_cast<uintptr_t>(chunk0) & 3) == 0;
        for (int pos = (hp & ~3) + tid * 4; pos < g.body; pos += 4 * kLzThreads) {
            const unsigned cword = cov[pos >> 5];
            const unsigned cw = (cword >> (pos & 31)) & 0xFu;
            if (cw == 0xFu) continue;                                   // four covered positions: nothing to count
            const unsigned v = srcAligned ? __ldg(reinterpret_cast<const unsigned*>(chunk0 + pos)) : gload4(chunk0 + pos, strm.lo, strm.hi);
            unsigned rank = litBase[pos >> 5] + (unsigned)__popc(~cword & ((1u << (pos & 31)) - 1u));
            if (cw == 0u && (rank & 3u) == 0u && pos + 4 <= g.body) {   // four literals in a row at an aligned place (incompressible data): one store
                *reinterpret_cast<unsigned*>(lits + rank) = v;
#pragma unroll
                for (int k = 0; k < 4; ++k) atomicAdd(&myh[(v >> (8 * k)) & 0xFFu], 1u);
                continue;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (!((cw >> k) & 1u)) { const unsigned b = (v >> (8 * k)) & 0xFFu; atomicAdd(&myh[b], 1u); lits[rank++] = (uint8_t)b; }
        }
    }
    __syncthreads();
    for (int i = tid; i < 316; i += kLzThreads) {
        unsigned s = 0;
        for (int w = 0; w < kHistCopies; ++w) s += hist4[w * kHistStride + i];
        if (i == 256) s += 1;                                       // end-of-block (encoder.cpp:470)
        job.hist[(size_t)slot * kHistStride + i] = s;
    }
    if (tid == 0) job.state[slot].ntok = ps.ntok;
    if (job.want_checksums == 1) {
        __syncthreads();                                            // the coverage bitmap is dead: its space is the scratch
        lz_checksums(chunk0, g.n, job.ck + 2 * (job.first_chunk + slot), smem + 32768);
    }
#ifdef ZZ_PHASE_TIMING
    PT(9);
    if (threadIdx.x == 0) for (int k = 0; k < 12; ++k) atomicAdd(&g_lzPhase[k], phaseAcc[k]);
#endif
}

// ------------------------------------------------------------------------------------------------
// K-HUFF : one warp per chunk.  Tie-breaks follow the libstdc++ heap layout, so the tree build itself is serial
// (lane 0, everything in shared memory); the loops around it (frequency loads, leaf compaction, sums, canonical
// codes, copies to global memory) use all lanes.
// ------------------------------------------------------------------------------------------------
// Heap entries are packed as frequency << 10 | id.  Comparisons look at the frequency only: ties are decided by
// the heap layout alone, as in the reference (huffman.cpp:55-63).
#define HF(v) ((v) >> 10)

__device__ __forceinline__ void heap_push_(unsigned* h, int hole, int top, unsigned v)
{
    int parent = (hole - 1) / 2;
    while (hole > top && HF(h[parent]) > HF(v)) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = v;
}

__device__ __forceinline__ void heap_adjust(unsigned* h, int hole, int len, unsigned v)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        const unsigned cr = h[child], cl = h[child - 1];
        unsigned pick = cr;
        if (HF(cr) > HF(cl)) { child--; pick = cl; }
        h[hole] = pick;
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    heap_push_(h, hole, top, v);
}

__device__ __forceinline__ int warp_sum(int v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CalcLengths (huffman.cpp:122-154), called by the whole warp; freqs, lens and heap are shared-memory arrays of the
// warp, n <= 286.  The heap shrinks by one entry per merge while the tree grows by one node, so node t (tree id n + t)
// is stored in the slot the heap has just vacated (slot nsym-1-t) as left | right << 10 | depth << 20: heap, tree
// links and depths share one 286-word array.  Leaf depths go straight to `lens`.
__device__ __noinline__ void calc_lengths(const int* freqs, int n, int maxLength, uint8_t* lens, unsigned* heap, int lane)
{
    int total = 0;
    for (int i = lane; i < n; i += 32) total += freqs[i];
    total = warp_sum(total);
    const unsigned ltMask = (1u << lane) - 1u;
    int minFreq = 0;
    for (;;) {
        int rn = 0;                                              // leaves in symbol order
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const int f = i < n ? freqs[i] : 0;
            const unsigned m = __ballot_sync(0xffffffffu, f != 0);
            if (i < n) lens[i] = f ? 1 : 0;                      // a used symbol that stays at depth 0 gets length 1
            if (f) heap[rn + __popc(m & ltMask)] = ((unsigned)(f > minFreq ? f : minFreq) << 10) | (unsigned)i;
            rn += __popc(m);
        }
        __syncwarp();
        int maxDepth = 0;
        if (lane == 0) {
            const int nsym = rn;
            if (rn >= 2)
                for (int parent = (rn - 2) / 2; ; --parent) { heap_adjust(heap, parent, rn, heap[parent]); if (parent == 0) break; }
            int tn = n;                                          // tree id of the next internal node
            while (rn >= 2) {
                unsigned a, b;
                if (rn > 1) { const unsigned v = heap[rn - 1]; heap[rn - 1] = heap[0]; heap_adjust(heap, 0, rn - 1, v); }
                a = heap[--rn];
                if (rn > 1) { const unsigned v = heap[rn - 1]; heap[rn - 1] = heap[0]; heap_adjust(heap, 0, rn - 1, v); }
                b = heap[--rn];
                const unsigned r = ((HF(a) + HF(b)) << 10) | (unsigned)tn;
                heap[rn++] = r;
                heap_push_(heap, rn - 1, 0, r);
                heap[rn] = (a & 1023u) | ((b & 1023u) << 10);   // node tn lives in the slot just vacated (depth 0 for now)
                ++tn;
            }
            // depths, root first: the root is the last node created = slot 1, node t sits in slot nsym-1-t
            for (int sl = 1; sl < nsym; ++sl) {
                const unsigned nd = heap[sl];
                const int dd = (int)(nd >> 20) + 1;
                const int ids[2] = { (int)(nd & 1023u), (int)((nd >> 10) & 1023u) };
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int id = ids[c];
                    if (id < n) { lens[id] = (uint8_t)dd; if (id != 0 && dd > maxDepth) maxDepth = dd; }   // leaf 0 is not looked at (huffman.cpp:108)
                    else { const int cs = nsym - 1 - (id - n); heap[cs] = (heap[cs] & 0xFFFFFu) | ((unsigned)dd << 20); }
                }
            }
        }
        maxDepth = __shfl_sync(0xffffffffu, maxDepth, 0);
        __syncwarp();
        if (maxDepth <= maxLength) return;
        const int step = total / (1 << maxLength);
        minFreq += step > 1 ? step : 1;
    }
}

__device__ __forceinline__ unsigned bit_reverse(unsigned v, int len) { return __brev(v) >> (32 - len); }

// huffman::generate (huffman.h:49-81): canonical codes, bit-reversed.  out[i] = bits | len << 16 (0 if unused); `out` may
// be global memory (coalesced stores).  Symbols of one length take consecutive codes in index order: the rank of a
// symbol among the equal-length symbols of its 32-symbol step comes from __match_any_sync, the running next_code per
// length lives in `next` (16 words of the warp's shared memory).
__device__ __noinline__ void generate_codes(const uint8_t* lens, int n, uint32_t* out, unsigned* next, int lane)
{
    if (lane < 16) next[lane] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) atomicAdd(&next[lens[i]], 1u);
    __syncwarp();
    if (lane == 0) {
        unsigned blPrev = 0, bits = 0;                           // bl_count[0] = 0 (huffman.h:60)
        for (int b = 1; b < 16; ++b) { bits = (bits + blPrev) << 1; blPrev = next[b]; next[b] = bits; }
        next[0] = 0;
    }
    __syncwarp();
    const unsigned ltMask = (1u << lane) - 1u;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const int len = i < n ? lens[i] : 0;
        const unsigned grp = __match_any_sync(0xffffffffu, len);
        const unsigned code = next[len] + (unsigned)__popc(grp & ltMask);
        __syncwarp();
        if ((grp >> lane) == 1u) next[len] += (unsigned)__popc(grp);
        __syncwarp();
        if (i < n) out[i] = len ? (bit_reverse(code, len) | ((uint32_t)len << 16)) : 0u;
    }
}

// FromLengths / AddRecords (huffman.cpp:158-216): records as value | payLoad << 8 (one lane; shared-memory arrays)
__device__ __noinline__ int rle_lengths(const uint8_t* lens, int n, unsigned short* rec, int* freqs19)
{
    int vn = 0, current = -1, count = 0;
    for (int i = 0; i <= n; ++i) {
        const int v = i < n ? lens[i] : -2;
        if (v == current) { count++; continue; }
        if (count > 0) {
            if (current == 0) {
                while (count >= 3) { int w = count < 138 ? count : 138; count -= w; rec[vn++] = (unsigned short)((w < 11 ? 17 : 18) | (w << 8)); }
            } else {
                rec[vn++] = (unsigned short)current; count--;
                while (count >= 3) { int w = count < 6 ? count : 6; count -= w; rec[vn++] = (unsigned short)(16 | (w << 8)); }
            }
            for (int k = 0; k < count; ++k) rec[vn++] = (unsigned short)current;
        }
        current = v; count = 1;
    }
    for (int i = 0; i < vn; ++i) freqs19[rec[i] & 0xFF]++;
    return vn;
}

struct HdrWriter {
    uint8_t* out; unsigned long long acc; int used; int pos;
    __device__ void put(unsigned bits, int n) {
        acc |= (unsigned long long)bits << used; used += n;
        while (used >= 8) { out[pos++] = (uint8_t)acc; acc >>= 8; used -= 8; }
    }
    __device__ void flush() { if (used > 0) { out[pos++] = (uint8_t)acc; acc = 0; used = 0; } }
};

__device__ const uint8_t kOrder[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };   // luts.cpp:62

__device__ __forceinline__ int len_extra_bits(int sym) { return (sym < 265 || sym == 285) ? 0 : (sym - 261) >> 2; }
__device__ __forceinline__ int dist_extra_bits(int sym) { return sym < 4 ? 0 : (sym - 2) >> 1; }

__device__ __forceinline__ uint32_t stored_size(int body)           // WriteUncompressedBlock: <= 65535 bytes + 5 per block
{
    if (body <= 0) return 0;
    const int blocks = (body + 0xFFFF - 1) / 0xFFFF;
    return (uint32_t)(body + 5 * blocks);
}

constexpr int kHuffWarps = 4;
constexpr int kHuffThreads = kHuffWarps * 32;

struct HuffShared {                         // one per warp
    int freq[316];
    unsigned heap[288];
    uint32_t metaCodes[20];
    int metaF[20];
    unsigned next[16];
    unsigned short symRec[288], distRec[32];
    uint8_t lens[336];
    uint8_t hdr[kHdrBytes];
};

__global__ void __launch_bounds__(kHuffThreads) k_huffman(Job job)
{
    __shared__ __align__(16) HuffShared sh[kHuffWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned slot = blockIdx.x * kHuffWarps + warp;
    if (slot >= job.nchunks) return;                                 // warp-uniform; the kernel has no CTA barrier
    HuffShared& w = sh[warp];
    const Geom g = chunk_geom(job, slot);
    ChunkState& st = job.state[slot];
    ChunkCodes& cc = job.codes[slot];
    const uint32_t tail = g.final ? 0u : 6u;                        // aligning 1-byte stored block (zzflate.cpp:116-120)

    if (job.level == 0 || g.body == 0) {
        if (lane == 0) {
            st.block_type = 0; st.hdr_bits = 0; st.total_bits = 0;
            st.out_bytes = stored_size(g.body) + tail;
            if (job.level == 0) st.ntok = 0;
        }
        return;
    }

    const uint32_t* hist = job.hist + (size_t)slot * kHistStride;
    for (int i = lane; i < 316; i += 32) w.freq[i] = (int)hist[i];
    if (lane < 20) w.metaF[lane] = 0;
    __syncwarp();

    calc_lengths(w.freq, 286, 15, w.lens, w.heap, lane);
    generate_codes(w.lens, 286, cc.lit, w.next, lane);
    int nSym = 0;
    if (lane == 0) nSym = rle_lengths(w.lens, 286, w.symRec, w.metaF);
    nSym = __shfl_sync(0xffffffffu, nSym, 0);
    int bits = 0;
    for (int i = lane; i < 286; i += 32) bits += w.freq[i] * (w.lens[i] + len_extra_bits(i));

    calc_lengths(w.freq + 286, 30, 15, w.lens + 286, w.heap, lane);
    generate_codes(w.lens + 286, 30, cc.dist, w.next, lane);
    int nDist = 0;
    if (lane == 0) nDist = rle_lengths(w.lens + 286, 30, w.distRec, w.metaF);
    nDist = __shfl_sync(0xffffffffu, nDist, 0);
    if (lane < 30) bits += w.freq[286 + lane] * (w.lens[286 + lane] + dist_extra_bits(lane));
    __syncwarp();

    uint8_t* metaL = w.lens + 316;
    calc_lengths(w.metaF, 19, 7, metaL, w.heap, lane);
    generate_codes(metaL, 19, w.metaCodes, w.next, lane);
    if (lane == 0) w.lens[335] = 0;
    __syncwarp();
    for (int i = lane; i < 336 / 4; i += 32)
        reinterpret_cast<uint32_t*>(cc.lens)[i] = reinterpret_cast<const uint32_t*>(w.lens)[i];

    for (int i = lane; i < nSym + nDist; i += 32) {
        const int v = (i < nSym ? w.symRec[i] : w.distRec[i - nSym]) & 0xFF;
        bits += metaL[v] + (v == 16 ? 2 : v == 17 ? 3 : v == 18 ? 7 : 0);
    }
    const long long total = 3 + 5 + 5 + 4 + 3 * 19 + (long long)warp_sum(bits);
    const long long required = (total + 8) / 8;                      // encoder.cpp:269
    if (required >= g.body) {                                        // UncompressedFallback (encoder.cpp:271-274)
        if (lane == 0) {
            st.total_bits = (uint64_t)total;
            st.block_type = 0; st.hdr_bits = 0;
            st.out_bytes = stored_size(g.body) + tail;
        }
        return;
    }
    int hdrBits = 17 + 57;
    if (lane == 0) {
        HdrWriter hw; hw.out = w.hdr; hw.acc = 0; hw.used = 0; hw.pos = 0;
        hw.put(g.final ? 1u : 0u, 1); hw.put(2u, 2);                     // StartBlock (encoder.cpp:143-147)
        hw.put(29u, 5); hw.put(29u, 5); hw.put(15u, 4);                  // encoder.cpp:283-285
        for (int i = 0; i < 19; ++i) hw.put(metaL[kOrder[i]], 3);
        for (int pass = 0; pass < 2; ++pass) {                           // WriteLengths (encoder.cpp:20-35)
            const unsigned short* rec = pass ? w.distRec : w.symRec;
            const int cnt = pass ? nDist : nSym;
            for (int i = 0; i < cnt; ++i) {
                const int v = rec[i] & 0xFF, pl = rec[i] >> 8;
                const uint32_t c = w.metaCodes[v];
                hw.put(c & 0xFFFF, (int)(c >> 16)); hdrBits += (int)(c >> 16);
                if (v == 16) { hw.put((unsigned)(pl - 3), 2); hdrBits += 2; }
                else if (v == 17) { hw.put((unsigned)(pl - 3), 3); hdrBits += 3; }
                else if (v == 18) { hw.put((unsigned)(pl - 11), 7); hdrBits += 7; }
            }
        }
        hw.flush();
        while (hw.pos & 3) w.hdr[hw.pos++] = 0;
        st.total_bits = (uint64_t)total;
        st.block_type = 2;
        st.hdr_bits = (uint32_t)hdrBits;
        // dynamic block occupies `total` bits; a non-final chunk appends 3 header bits, pads, then LEN/NLEN + 1 byte
        st.out_bytes = g.final ? (uint32_t)((total + 7) / 8) : (uint32_t)((total + 3 + 7) / 8) + 5u;
    }
    hdrBits = __shfl_sync(0xffffffffu, hdrBits, 0);
    __syncwarp();
    for (int i = lane; i < (hdrBits + 31) / 32; i += 32)
        reinterpret_cast<uint32_t*>(cc.hdr)[i] = reinterpret_cast<const uint32_t*>(w.hdr)[i];
}

// ------------------------------------------------------------------------------------------------
// K-HUFF, batch variant: one THREAD per chunk, 32 chunks per CTA.  Same algorithm as the warp-per-chunk kernel above; the
// heaps of the 32 chunks are interleaved by lane ([index][lane]) so that they never conflict on a bank.  Each lane
// runs the whole chain serially, so a launch takes ~0.9 ms whatever its size, but all lanes of a warp do useful work:
// this variant is used for launches large enough to fill the GPU several times over, the warp-per-chunk kernel
// (half the latency) for the smaller pieces of the host-buffer pipeline.
// ------------------------------------------------------------------------------------------------
#ifndef ZZ_HUFF_LANES
#define ZZ_HUFF_LANES 32
#endif
constexpr int kHuffLanes = ZZ_HUFF_LANES;
#define HS(i) ((i) * kHuffLanes)

__device__ __forceinline__ void heap_push_L(unsigned* h, int hole, int top, unsigned v)
{
    int parent = (hole - 1) / 2;
    while (hole > top && HF(h[HS(parent)]) > HF(v)) {
        h[HS(hole)] = h[HS(parent)];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[HS(hole)] = v;
}

__device__ __forceinline__ void heap_adjustL(unsigned* h, int hole, int len, unsigned v)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        const unsigned cr = h[HS(child)], cl = h[HS(child - 1)];
        unsigned pick = cr;
        if (HF(cr) > HF(cl)) { child--; pick = cl; }
        h[HS(hole)] = pick;
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h[HS(hole)] = h[HS(child - 1)];
        hole = child - 1;
    }
    heap_push_L(h, hole, top, v);
}

// CalcLengths (huffman.cpp:122-154).  n <= 286.  `heap` points at this lane's shared-memory column.  The heap
// shrinks by one entry per merge while the tree grows by one node, so node t (tree id n + t) is stored in the slot the
// heap has just vacated (slot nsym-1-t) as left | right << 10 | depth << 20: heap, tree links and depths share one
// 286-word column and the latency-critical chain never leaves shared memory.  Leaf depths go straight to `lens`.
__device__ __noinline__ void calc_lengthsL(const int* freqs, int n, int maxLength, uint8_t* lens, unsigned* heap)
{
    int total = 0;
    for (int i = 0; i < n; ++i) total += freqs[i];
    int minFreq = 0;
    for (;;) {
        int rn = 0;
        for (int i = 0; i < n; ++i) {
            lens[i] = freqs[i] == 0 ? 0 : 1;                        // a used symbol that stays at depth 0 gets length 1
            if (freqs[i] == 0) continue;
            const unsigned f = (unsigned)(freqs[i] > minFreq ? freqs[i] : minFreq);
            heap[HS(rn++)] = (f << 10) | (unsigned)i;
        }
        const int nsym = rn;
        if (rn >= 2)
            for (int parent = (rn - 2) / 2; ; --parent) { heap_adjustL(heap, parent, rn, heap[HS(parent)]); if (parent == 0) break; }
        int tn = n;                                              // tree id of the next internal node
        while (rn >= 2) {
            unsigned a, b;
            if (rn > 1) { const unsigned v = heap[HS(rn - 1)]; heap[HS(rn - 1)] = heap[0]; heap_adjustL(heap, 0, rn - 1, v); }
            a = heap[HS(--rn)];
            if (rn > 1) { const unsigned v = heap[HS(rn - 1)]; heap[HS(rn - 1)] = heap[0]; heap_adjustL(heap, 0, rn - 1, v); }
            b = heap[HS(--rn)];
            const unsigned r = ((HF(a) + HF(b)) << 10) | (unsigned)tn;
            heap[HS(rn++)] = r;
            heap_push_L(heap, rn - 1, 0, r);
            heap[HS(rn)] = (a & 1023u) | ((b & 1023u) << 10);       // node tn lives in the slot just vacated (depth 0 for now)
            ++tn;
        }
        // depths, root first: the root is the last node created = slot 1, node t sits in slot nsym-1-t
        int maxDepth = 0;
        for (int sl = 1; sl < nsym; ++sl) {
            const unsigned nd = heap[HS(sl)];
            const int dd = (int)(nd >> 20) + 1;
            const int ids[2] = { (int)(nd & 1023u), (int)((nd >> 10) & 1023u) };
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int id = ids[c];
                if (id < n) { lens[id] = (uint8_t)dd; if (id != 0 && dd > maxDepth) maxDepth = dd; }   // leaf 0 is not looked at (huffman.cpp:108)
                else { const int cs = nsym - 1 - (id - n); heap[HS(cs)] = (heap[HS(cs)] & 0xFFFFFu) | ((unsigned)dd << 20); }
            }
        }
        if (maxDepth <= maxLength) return;
        int step = total / (1 << maxLength);
        minFreq += step > 1 ? step : 1;
    }
}

// huffman::generate (huffman.h:49-81): canonical codes, bit-reversed.  out[i] = bits | len << 16 (0 if unused)
__device__ __noinline__ void generate_codesL(const uint8_t* lens, int n, uint32_t* out)
{
    int blCount[16];
    for (int i = 0; i < 16; ++i) blCount[i] = 0;
    for (int i = 0; i < n; ++i) blCount[lens[i]]++;
    unsigned nextCode[16];
    unsigned bits = 0;
    blCount[0] = 0;
    nextCode[0] = 0;
    for (int b = 1; b < 16; ++b) { bits = (bits + blCount[b - 1]) << 1; nextCode[b] = bits; }
    for (int i = 0; i < n; ++i) {
        const int len = lens[i];
        if (len == 0) { out[i] = 0; continue; }
        out[i] = bit_reverse(nextCode[len], len) | ((uint32_t)len << 16);
        nextCode[len]++;
    }
}

constexpr int kHuffLThreads = kHuffLanes;
constexpr int kHuffLSmem = 286 * kHuffLanes * 4;

__global__ void __launch_bounds__(kHuffLThreads) k_huffman_lanes(Job job)
{
    const unsigned slot = blockIdx.x * kHuffLThreads + threadIdx.x;
    if (slot >= job.nchunks) return;
    const Geom g = chunk_geom(job, slot);
    ChunkState& st = job.state[slot];
    ChunkCodes& cc = job.codes[slot];
    const uint32_t tail = g.final ? 0u : 6u;                        // aligning 1-byte stored block (zzflate.cpp:116-120)

    if (job.level == 0 || g.body == 0) {
        st.block_type = 0; st.hdr_bits = 0; st.total_bits = 0;
        st.out_bytes = stored_size(g.body) + tail;
        if (job.level == 0) st.ntok = 0;
        return;
    }

    extern __shared__ __align__(16) uint8_t hsm[];
    unsigned* heap = reinterpret_cast<unsigned*>(hsm) + threadIdx.x;
    int freq[286];
    int metaF[19];
    unsigned short symRec[288], distRec[32];
    uint32_t metaCodes[19];
    uint8_t metaL[19];
    uint8_t lens[336];                  // thread-local while the trees are built; copied to cc.lens at the end

    const uint32_t* hist = job.hist + (size_t)slot * kHistStride;
    long long bits = 0;

    for (int i = 0; i < 19; ++i) metaF[i] = 0;
    for (int i = 0; i < 286; ++i) freq[i] = (int)hist[i];
    calc_lengthsL(freq, 286, 15, lens, heap);
    generate_codesL(lens, 286, cc.lit);
    // free mode: only the used part of each alphabet goes into the header (HLIT / HDIST; the reference always writes 286 / 30,
    // encoder.cpp:283-285)
    int nl = 286, nd = 30;
    if (job.mode) while (nl > 257 && lens[nl - 1] == 0) --nl;
    const int nSym = rle_lengths(lens, nl, symRec, metaF);
    for (int i = 0; i < 286; ++i) bits += (long long)freq[i] * (lens[i] + len_extra_bits(i));

    for (int i = 0; i < 30; ++i) freq[i] = (int)hist[286 + i];
    calc_lengthsL(freq, 30, 15, lens + 286, heap);
    generate_codesL(lens + 286, 30, cc.dist);
    if (job.mode) while (nd > 1 && lens[286 + nd - 1] == 0) --nd;
    const int nDist = rle_lengths(lens + 286, nd, distRec, metaF);
    for (int i = 0; i < 30; ++i) bits += (long long)freq[i] * (lens[286 + i] + dist_extra_bits(i));

    calc_lengthsL(metaF, 19, 7, metaL, heap);
    generate_codesL(metaL, 19, metaCodes);
    for (int i = 0; i < 19; ++i) lens[316 + i] = metaL[i];
    lens[335] = 0;
    for (int i = 0; i < 336 / 4; ++i)
        reinterpret_cast<uint32_t*>(cc.lens)[i] = lens[4 * i] | (lens[4 * i + 1] << 8) | (lens[4 * i + 2] << 16) | ((uint32_t)lens[4 * i + 3] << 24);

    int hclen = 19;
    if (job.mode) while (hclen > 4 && metaL[kOrder[hclen - 1]] == 0) --hclen;
    long long total = 3 + 5 + 5 + 4 + 3 * hclen + bits;
    for (int pass = 0; pass < 2; ++pass) {
        const unsigned short* rec = pass ? distRec : symRec;
        const int cnt = pass ? nDist : nSym;
        for (int i = 0; i < cnt; ++i) {
            const int v = rec[i] & 0xFF;
            total += metaL[v] + (v == 16 ? 2 : v == 17 ? 3 : v == 18 ? 7 : 0);
        }
    }
    st.total_bits = (uint64_t)total;
    const long long required = (total + 8) / 8;                      // encoder.cpp:269
    if (required >= g.body) {                                        // UncompressedFallback (encoder.cpp:271-274)
        st.block_type = 0; st.hdr_bits = 0;
        st.out_bytes = stored_size(g.body) + tail;
        return;
    }
    st.block_type = 2;
    HdrWriter w; w.out = cc.hdr; w.acc = 0; w.used = 0; w.pos = 0;
    w.put(g.final ? 1u : 0u, 1); w.put(2u, 2);                       // StartBlock (encoder.cpp:143-147)
    w.put((unsigned)(nl - 257), 5); w.put((unsigned)(nd - 1), 5); w.put((unsigned)(hclen - 4), 4);     // 29, 29, 15 in reference-equivalent mode (encoder.cpp:283-285)
    for (int i = 0; i < hclen; ++i) w.put(metaL[kOrder[i]], 3);
    int hdrBits = 17 + 3 * hclen;
    for (int pass = 0; pass < 2; ++pass) {                           // WriteLengths (encoder.cpp:20-35)
        const unsigned short* rec = pass ? distRec : symRec;
        const int cnt = pass ? nDist : nSym;
        for (int i = 0; i < cnt; ++i) {
            const int v = rec[i] & 0xFF, pl = rec[i] >> 8;
            const uint32_t c = metaCodes[v];
            w.put(c & 0xFFFF, (int)(c >> 16)); hdrBits += (int)(c >> 16);
            if (v == 16) { w.put((unsigned)(pl - 3), 2); hdrBits += 2; }
            else if (v == 17) { w.put((unsigned)(pl - 3), 3); hdrBits += 3; }
            else if (v == 18) { w.put((unsigned)(pl - 11), 7); hdrBits += 7; }
        }
    }
    w.flush();
    st.hdr_bits = (uint32_t)hdrBits;
    // dynamic block occupies `total` bits; a non-final chunk appends 3 header bits, pads, then LEN/NLEN + 1 byte
    st.out_bytes = g.final ? (uint32_t)((total + 7) / 8) : (uint32_t)((total + 3 + 7) / 8) + 5u;
}

// ------------------------------------------------------------------------------------------------
// K-OFFS : exclusive scan of chunk sizes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_offsets(Job job)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    __shared__ unsigned long long cnt[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = job.total[0]; cnt[0] = 0; cnt[1] = 0; }
    __syncthreads();
    unsigned long long matches = 0, stored = 0;
    for (unsigned base = 0; base < job.nchunks; base += 1024) {
        const unsigned slot = base + tid;
        unsigned long long v = 0;
        if (slot < job.nchunks) {
            v = job.state[slot].out_bytes;
            matches += job.state[slot].ntok;
            stored += job.state[slot].block_type == 0 ? 1 : 0;
        }
        unsigned long long inc = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long s = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
            wsum[lane] = s;
        }
        __syncthreads();
        const unsigned long long before = carry + (warp ? wsum[warp - 1] : 0) + inc - v;
        if (slot < job.nchunks) job.state[slot].out_off = before;
        __syncthreads();
        if (tid == 0) carry += wsum[31];
        __syncthreads();
    }
    atomicAdd(&cnt[0], matches); atomicAdd(&cnt[1], stored);
    __syncthreads();
    if (tid == 0) { job.total[0] = carry; job.total[2] += cnt[0]; job.total[3] += cnt[1]; }
}

// ------------------------------------------------------------------------------------------------
// K-EMIT : one CTA per chunk
// ------------------------------------------------------------------------------------------------
constexpr int kOutWords = (kMaxChunk + 64) / 4;

__device__ __forceinline__ void copy_out(uint8_t* D, const unsigned* out32, unsigned bytes)
{
    const uint8_t* out8 = reinterpret_cast<const uint8_t*>(out32);
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(D) & 3);
    unsigned head = mis ? 4 - mis : 0; if (head > bytes) head = bytes;
    if ((unsigned)tid < head) D[tid] = out8[tid];
    const unsigned nfull = (bytes - head) >> 2;
    unsigned* Dw = reinterpret_cast<unsigned*>(D + head);
    const int sh = (int)head * 8;
    for (unsigned i = tid; i < nfull; i += nt) Dw[i] = __funnelshift_r(out32[i], out32[i + 1], sh);
    const unsigned done = head + nfull * 4;
    if ((unsigned)tid < bytes - done) D[done + tid] = out8[done + tid];
}

// global -> global copy of `bytes` bytes at arbitrary alignments (stored blocks): whole destination words, each from two
// aligned source words (a source word that holds a needed byte never leaves the buffer's pages)
__device__ __forceinline__ void copy_g2g(uint8_t* D, const uint8_t* S, unsigned bytes)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(D) & 3);
    unsigned head = mis ? 4 - mis : 0; if (head > bytes) head = bytes;
    if ((unsigned)tid < head) D[tid] = S[tid];
    const unsigned nfull = (bytes - head) >> 2;
    unsigned* Dw = reinterpret_cast<unsigned*>(D + head);
    const uint8_t* S2 = S + head;
    const unsigned smis = (unsigned)(reinterpret_cast<uintptr_t>(S2) & 3);
    const unsigned* Sw = reinterpret_cast<const unsigned*>(S2 - smis);
    if (smis == 0) {
#pragma unroll 8
        for (unsigned i = tid; i < nfull; i += nt) Dw[i] = __ldg(Sw + i);
    } else {
        const int sh = (int)smis * 8;
#pragma unroll 8
        for (unsigned i = tid; i < nfull; i += nt) Dw[i] = __funnelshift_r(__ldg(Sw + i), __ldg(Sw + i + 1), sh);
    }
    const unsigned done = head + nfull * 4;
    if ((unsigned)tid < bytes - done) D[done + tid] = S[done + tid];
}

// ------------------------------------------------------------------------------------------------
// K-EMIT, symbol-parallel.  The bit stream interleaves two sequences that are each compact in memory: the
// literals (K-MATCH leaves the literal bytes of the block in order in the chunk's info row) and the matches (the token
// list).  With l_k = number of literals before match k, the write offsets are
//     literal i : header + (code bits of literals < i) + (bits of the matches k with l_k <= i)
//     match k   : header + (code bits of literals < l_k) + (bits of the matches < k)
// so both are prefix sums over compact arrays, and every lane of a warp does the same work: no walk over positions,
// no data-dependent mix of literals and matches per thread.
//   pass A (tokens, 1024 per step): match bits, l_k and the match-bit prefix M_k, kept in the chunk's candidate row
//   pass B (literals, 2048 per step): the matches that sit inside the step scatter their bits onto the literal index
//           they precede (shared-memory D), one block scan of (literal bits, D) gives all offsets; literals are written,
//           then the step's matches at header + literal-prefix(l_k) + M_k.
// The output image is assembled in shared memory; two CTAs per SM.
// ------------------------------------------------------------------------------------------------
constexpr int kEmit2Threads = 512;
constexpr int kLitStep = 4 * kEmit2Threads;           // literals per step of pass B (four per thread)
constexpr int kMaxLitSteps = kMaxChunk / kLitStep + 2;
constexpr int kEmit2Smem = kOutWords * 4 + (286 + 259 + 30 + 1) * 4 + (kLitStep + 4) * 4 * 2 + kMaxTokens + 16;
static_assert((kOutWords * 4 + (286 + 259 + 30 + 1) * 4) % 16 == 0, "D must be 16-byte aligned");

struct TokInfo { uint32_t lit; uint32_t mbits; };      // l_k, M_k

__device__ __forceinline__ void or_bits(unsigned* out, unsigned bitOff, unsigned bits, unsigned n)   // 1 <= n <= 32
{
    const unsigned w = bitOff >> 5, sh = bitOff & 31u;
    atomicOr(&out[w], bits << sh);
    if (sh + n > 32u) atomicOr(&out[w + 1], bits >> (32u - sh));
}

__global__ void __launch_bounds__(kEmit2Threads, 2) k_emit2(Job job)
{
    extern __shared__ __align__(16) uint8_t smem[];
    unsigned* out = reinterpret_cast<unsigned*>(smem);
    unsigned* litc = out + kOutWords;        // bits | len << 24
    unsigned* lenc = litc + 286;             // merged length codes (CreateMergedLengthCodes, encoder.cpp:126-133)
    unsigned* dstc = lenc + 259;
    unsigned* D = dstc + 30 + 1;             // match bits that precede each literal of the step (16-byte aligned: 65600 + 2304)
    unsigned* PL = D + kLitStep + 4;         // literal-bit prefix of each literal index of the step
    uint8_t* MBs = reinterpret_cast<uint8_t*>(PL + kLitStep + 4);      // bits of every match
    __shared__ unsigned wsumA[32], wsumB[32];
    __shared__ unsigned sCarryA, sCarryB;
    __shared__ unsigned tokStart[kMaxLitSteps + 1];

    const unsigned slot = blockIdx.x;
    const Geom g = chunk_geom(job, slot);
    const ChunkState st = job.state[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kEmit2Threads / 32;
    const uint8_t* chunk0 = job.src + g.off;
    if (st.out_off + st.out_bytes > job.cap) {                        // never write past the caller's buffer
        if (tid == 0) atomicOr(reinterpret_cast<unsigned long long*>(&job.total[1]), 1ull);
        return;
    }
    uint8_t* Dst = job.dst + st.out_off;

    if (st.block_type == 0) {
        // stored blocks of <= 65535 bytes (encoder.cpp:482-502), then the aligning block for non-final chunks
        int written = 0; unsigned o = 0;
        while (written < g.body) {
            const int len = min(g.body - written, 0xFFFF);
            if (tid == 0) {
                Dst[o] = (uint8_t)((g.final && written + len == g.body) ? 1 : 0);
                Dst[o + 1] = (uint8_t)len; Dst[o + 2] = (uint8_t)(len >> 8);
                Dst[o + 3] = (uint8_t)~len; Dst[o + 4] = (uint8_t)((~len) >> 8);
            }
            copy_g2g(Dst + o + 5, chunk0 + written, (unsigned)len);
            o += 5 + len; written += len;
        }
        if (!g.final && tid == 0) {
            Dst[o] = 0; Dst[o + 1] = 1; Dst[o + 2] = 0; Dst[o + 3] = 0xFE; Dst[o + 4] = 0xFF; Dst[o + 5] = chunk0[g.n - 1];
        }
        return;
    }

    for (int i = tid; i < kOutWords; i += kEmit2Threads) out[i] = 0;
    const ChunkCodes& cc = job.codes[slot];
    for (int i = tid; i < 286; i += kEmit2Threads) { uint32_t c = cc.lit[i]; litc[i] = (c & 0xFFFF) | ((c >> 16) << 24); }
    for (int i = tid; i < 30; i += kEmit2Threads) { uint32_t c = cc.dist[i]; dstc[i] = (c & 0xFFFF) | ((c >> 16) << 24); }
    if (tid == 0) { sCarryA = 0; sCarryB = 0; }
    __syncthreads();
    for (int i = tid; i < 259; i += kEmit2Threads) {
        unsigned v = 0;
        if (i >= 3) {
            int eb, ev; const int sym = len_symbol(i, eb, ev);
            const unsigned c = litc[sym]; const unsigned cl = c >> 24;
            v = ((c & 0xFFFFFF) | ((unsigned)ev << cl)) | ((cl + eb) << 24);
        }
        lenc[i] = v;
    }
    // header bit string
    for (unsigned i = tid; i < (st.hdr_bits + 31) / 32; i += kEmit2Threads) {
        unsigned v = reinterpret_cast<const uint32_t*>(cc.hdr)[i];
        const unsigned rem = st.hdr_bits - i * 32;
        if (rem < 32) v &= (1u << rem) - 1u;
        atomicOr(&out[i], v);
    }
    __syncthreads();
    const uint32_t* tokA = job.tokA + (size_t)slot * kMaxTokens;
    const uint16_t* tokD = job.tokD + (size_t)slot * kMaxTokens;
    const int ntok = (int)st.ntok;
    const int nlit = (int)st.nlit;
    const uint8_t* lits = job.info + (size_t)slot * job.chunk;
    TokInfo* tinfo = reinterpret_cast<TokInfo*>(job.cand + (size_t)slot * job.chunk);    // 8 bytes per token <= 2 * chunk bytes
    const unsigned hdr = st.hdr_bits;

    auto matchCode = [&](int len, unsigned dist, unsigned& lo, unsigned& loN, unsigned& hi, unsigned& hiN) {
        const unsigned lc = lenc[len];
        lo = lc & 0xFFFFFFu; loN = lc >> 24;
        int eb, ev; const int ds = dist_symbol((int)dist, eb, ev);
        const unsigned dc = dstc[ds]; const unsigned dl = dc >> 24;
        hi = (dc & 0xFFFFFFu) | ((unsigned)ev << dl); hiN = dl + (unsigned)eb;
    };

    // ---- pass A: per token l_k = start - (match bytes before) and M_k = match bits before ----
    // two consecutive tokens per thread and step (1024 per step): half as many block scans
    auto load2 = [&](int k, uint2& t2, unsigned& d2) {
        t2 = make_uint2(0u, 0u); d2 = 0;
        if (k + 1 < ntok) { t2 = __ldg(reinterpret_cast<const uint2*>(tokA + k)); d2 = __ldg(reinterpret_cast<const unsigned*>(tokD + k)); }
        else if (k < ntok) { t2.x = __ldg(tokA + k); d2 = __ldg(tokD + k); }
    };
    uint2 tNext; unsigned dNext;
    load2(2 * tid, tNext, dNext);
    for (int k0 = 0; k0 < ntok; k0 += 2 * kEmit2Threads) {
        const int k = k0 + 2 * tid;
        const uint2 t2 = tNext; const unsigned d2 = dNext;
        load2(k + 2 * kEmit2Threads, tNext, dNext);                       // next step's tokens
        unsigned mb[2] = { 0, 0 }, len[2] = { 0, 0 }, ms[2] = { 0, 0 };
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (k + u < ntok) {
                const uint32_t t = u ? t2.y : t2.x;
                ms[u] = t & 0xFFFFu; len[u] = t >> 16;
                unsigned lo, loN, hi, hiN; matchCode((int)len[u], u ? d2 >> 16 : d2 & 0xFFFFu, lo, loN, hi, hiN);
                mb[u] = loN + hiN;
                MBs[k + u] = (uint8_t)mb[u];
            }
        }
        const unsigned sumA = mb[0] + mb[1], sumB = len[0] + len[1];
        unsigned ia = sumA, ib = sumB;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ta; ib += tb; }
        }
        if (lane == 31) { wsumA[warp] = ia; wsumB[warp] = ib; }
        __syncthreads();
        unsigned beforeA = sCarryA, beforeB = sCarryB, totA, totB;
        {   // every warp scans the 16 warp totals with shuffles
            unsigned va = lane < nwarps ? wsumA[lane] : 0u, vb = lane < nwarps ? wsumB[lane] : 0u;
            for (int o = 1; o < nwarps; o <<= 1) {
                const unsigned ta = __shfl_up_sync(0xffffffffu, va, o), tb = __shfl_up_sync(0xffffffffu, vb, o);
                if (lane >= o) { va += ta; vb += tb; }
            }
            totA = __shfl_sync(0xffffffffu, va, nwarps - 1); totB = __shfl_sync(0xffffffffu, vb, nwarps - 1);
            const unsigned pa = __shfl_sync(0xffffffffu, va, (warp + 31) & 31), pb = __shfl_sync(0xffffffffu, vb, (warp + 31) & 31);
            if (warp) { beforeA += pa; beforeB += pb; }
        }
        unsigned exA = beforeA + ia - sumA, exB = beforeB + ib - sumB;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (k + u < ntok) { TokInfo ti; ti.lit = ms[u] - exB; ti.mbits = exA; tinfo[k + u] = ti; }
            exA += mb[u]; exB += len[u];
        }
        __syncthreads();
        if (tid == 0) { sCarryA += totA; sCarryB += totB; }
    }
    __syncthreads();
    // first token of every literal step: tokStart[s] = first k with l_k >= s * kLitStep (l_k never decreases)
    const int nsteps = nlit / kLitStep + 1;                            // the last step also holds the virtual index nlit
    if (tid <= nsteps) {
        int lo = 0, hi = ntok;
        const unsigned want = (unsigned)tid * kLitStep;
        if (tid == nsteps) lo = ntok;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (tinfo[mid].lit >= want) hi = mid; else lo = mid + 1; }
        tokStart[tid] = (unsigned)lo;
    }
    if (tid == 0) { sCarryA = 0; sCarryB = 0; }                        // now: literal bits / match bits before the step
    for (int i = tid; i < kLitStep + 4; i += kEmit2Threads) D[i] = 0;
    __syncthreads();

    // ---- pass B: literal steps ----
    for (int s = 0; s < nsteps; ++s) {
        const int i0 = s * kLitStep;
        const int ka = (int)tokStart[s], kb = (int)tokStart[s + 1];
        // the step's global loads are issued before its first barrier: this thread's literals and its first token
        const int j0 = 4 * tid, i = i0 + j0;
        unsigned v = 0;
        if (i < nlit) v = __ldg(reinterpret_cast<const unsigned*>(lits + i));
        const int k1 = ka + tid;
        TokInfo ti1; ti1.lit = 0; ti1.mbits = 0; uint32_t t1 = 0; unsigned d1 = 0;
        if (k1 < kb) { ti1 = tinfo[k1]; t1 = __ldg(tokA + k1); d1 = __ldg(tokD + k1); }
        // D is all zero here: every thread clears the four entries it has read (below), and the scatter of a step
        // starts after the previous step's barrier in front of the match phase
        if (k1 < kb) atomicAdd(&D[ti1.lit - (unsigned)i0], (unsigned)MBs[k1]);
        for (int k = k1 + kEmit2Threads; k < kb; k += kEmit2Threads) atomicAdd(&D[tinfo[k].lit - (unsigned)i0], (unsigned)MBs[k]);
        __syncthreads();
        const uint4 d4 = *reinterpret_cast<const uint4*>(D + j0);
        *reinterpret_cast<uint4*>(D + j0) = make_uint4(0u, 0u, 0u, 0u);
        unsigned code[4], n[4];
        const unsigned dd[4] = { d4.x, d4.y, d4.z, d4.w };
        unsigned sumN = 0, sumD = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned c = i + q < nlit ? litc[(v >> (8 * q)) & 0xFFu] : 0u;
            code[q] = c & 0xFFFFFFu; n[q] = c >> 24;
            sumN += n[q]; sumD += dd[q];
        }
        unsigned ia = sumN, ib = sumD;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ta; ib += tb; }
        }
        if (lane == 31) { wsumA[warp] = ia; wsumB[warp] = ib; }
        __syncthreads();
        unsigned pn = sCarryA, pd = sCarryB, totA, totB;
        {
            unsigned va = lane < nwarps ? wsumA[lane] : 0u, vb = lane < nwarps ? wsumB[lane] : 0u;
            for (int o = 1; o < nwarps; o <<= 1) {
                const unsigned ta = __shfl_up_sync(0xffffffffu, va, o), tb = __shfl_up_sync(0xffffffffu, vb, o);
                if (lane >= o) { va += ta; vb += tb; }
            }
            totA = __shfl_sync(0xffffffffu, va, nwarps - 1); totB = __shfl_sync(0xffffffffu, vb, nwarps - 1);
            const unsigned pa = __shfl_sync(0xffffffffu, va, (warp + 31) & 31), pb = __shfl_sync(0xffffffffu, vb, (warp + 31) & 31);
            if (warp) { pn += pa; pd += pb; }
        }
        pn += ia - sumN; pd += ib - sumD;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pd += dd[q];
            PL[j0 + q] = pn;
            if (n[q]) or_bits(out, hdr + pn + pd, code[q], n[q]);
            pn += n[q];
        }
        __syncthreads();
        if (tid == 0) { sCarryA += totA; sCarryB += totB; }
        for (int k = k1; k < kb; k += kEmit2Threads) {
            TokInfo ti = ti1; uint32_t t = t1; unsigned dist = d1;
            if (k != k1) { ti = tinfo[k]; t = __ldg(tokA + k); dist = __ldg(tokD + k); }
            unsigned lo, loN, hi, hiN; matchCode((int)(t >> 16), dist, lo, loN, hi, hiN);
            const unsigned off = hdr + PL[ti.lit - (unsigned)i0] + ti.mbits;
            or_bits(out, off, lo, loN); or_bits(out, off + loN, hi, hiN);
        }
        // no barrier here: the next step's PL is written after two more barriers, its D entries are already clear
    }
    __syncthreads();

    if (tid == 0) {
        unsigned q = hdr + sCarryA + sCarryB;
        const unsigned eob = litc[256];
        or_bits(out, q, eob & 0xFFFFFFu, eob >> 24);
        q += eob >> 24;
        if ((unsigned long long)q != st.total_bits)                 // K-HUFF's exact size and K-EMIT must agree
            atomicOr(reinterpret_cast<unsigned long long*>(&job.total[1]), 2ull);
        unsigned bytes = (q + 7) >> 3;
        if (!g.final) {
            const unsigned bp = (q + 3 + 7) >> 3;                   // 3 header bits of the stored block, then pad
            uint8_t* o8 = reinterpret_cast<uint8_t*>(out);
            o8[bp] = 1; o8[bp + 1] = 0; o8[bp + 2] = 0xFE; o8[bp + 3] = 0xFF; o8[bp + 4] = chunk0[g.n - 1];
            bytes = bp + 5;
        }
        if (bytes != st.out_bytes) atomicOr(reinterpret_cast<unsigned long long*>(&job.total[1]), 4ull);
    }
    __syncthreads();
    copy_out(Dst, out, st.out_bytes);
}

// ------------------------------------------------------------------------------------------------
// K-FIXED (level 1) : WriteBlockFixedHuff, encoder.cpp:329-373.  Only the positions the walk visits enter the hash table
// (table[h(i+1)] = i), so the candidates depend on the parse and the chunk is one sequential chain.  A pair of warps
// takes a chunk: the *walker* (warp 0) advances in steps of 32 positions and keeps everything the next step depends
// on -- hash table, candidates, the step's matches; the *emitter* (warp 1) turns what the walker decided into codes,
// scans the bit counts, packs and flushes.  Bits go to the chunk's scratch slot; K-OFFS + k_gather place them in the
// stream.
//
// A step.  The walker presumes that all 32 positions are visited.  Every lane knows two match lengths before the step
// is settled: mOld against the table's entry (its candidate if no lower lane of the step with the same hash is visited)
// and mLow against the nearest lower lane of its hash group (its candidate if that lane is visited).  The matches of
// the step are then settled without any further memory access.  A lane without a lower lane of its hash group has one
// possible candidate (the table's), so the window in front of the first lane that does have one is settled in parallel
// (the real match starts are an orbit, collected by pointer doubling: two REDUX / SHFL rounds on average); behind it
// the walk takes one match per turn (one shuffle), and the lanes that depend on the visited set are evaluated against
// it when the walk reaches them (0.5 turns per step on text).  A lane whose true candidate is neither (a lower lane of
// its group that is not the nearest one) ends the step in front of it; a match of >= 8 bytes (exact length needed: a
// warp-wide gather) ends the step behind it.  On the text workload a step covers 32.7 positions (the first
// formulation, which ended a step at its first match: 8.6).  tools/model/l1_model.c states the step logic in plain C
// and is fuzzed against the sequential walk.
//
// Whether two lanes of a step share a hash is found out on the table itself: every lane reads its slot, then writes
// its own position, then reads the slot back -- a lane that finds another position there shares the slot.  Only then is
// the same-hash mask computed (13 ballots; __match_any_sync held the warp for ~300 cycles per step) and the slots put
// back; otherwise (94 % of the steps on random bytes, 57 % on text) every lane owns its slot and the lanes that turn out
// not to be visited restore what they found.  The bytes of the table's candidates are requested before that exchange
// and used after it.
//
// The input is staged through a 1 KiB shared-memory ring in 256-byte tiles (cp.async, 16 bytes per lane, two tiles
// ahead of the walk), the walker leaves one word per lane (nothing / literal / match / end of block) in a two-slot
// queue, two named barriers per slot (full / empty: barrier.sync on one side, barrier.arrive on the other) order it.
// 17.3 KiB of shared memory and five barriers per chunk: twelve chunks per SM (the hash table sets that limit).
// ncu (profiles/r02m_fixed.md): the walker is one dependent chain at one warp per scheduler -- the gather of the
// candidates' bytes is 11-16 % of it, the rest is fixed-latency dependencies spread over the whole step; at twelve
// chunks per SM the issue slots are ~70 % busy, so every instruction taken out of a step shows in the throughput.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned fixed_lit_code(unsigned v, int& len)     // fixedhuffmanluts.cpp:5 (RFC 1951 3.2.6)
{
    if (v < 144) { len = 8; return __brev(0x30u + v) >> 24; }
    if (v < 256) { len = 9; return __brev(0x190u + (v - 144)) >> 23; }
    if (v < 280) { len = 7; return __brev(v - 256) >> 25; }
    len = 8; return __brev(0xC0u + (v - 280)) >> 24;
}

__device__ __forceinline__ unsigned fixed_literal_code(unsigned v, int& len)     // the same for v < 256 (two ranges only)
{
    const bool low = v < 144;
    len = low ? 8 : 9;
    return __brev(v + (low ? 0x30u : 0x100u)) >> (low ? 24 : 23);
}

constexpr int kFxTile = 256, kFxRing = 4 * kFxTile, kFxAhead = 2;

__device__ __forceinline__ void cp_async16(void* smemDst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct FxRing {
    uint8_t* ring;            // kFxRing bytes, 16-byte aligned; byte u of the chunk's staging coordinate lives at ring[u & (kFxRing-1)]
    long long gb;             // global address of u = 0 (16-byte aligned; may lie before the readable stream)
    long long lo, hi;         // readable stream as addresses
    int issued;               // next tile to stage
    int ready;                // tiles <= ready are visible to the warp
};

__device__ __forceinline__ void fx_issue_tile(const FxRing& r, int t, int lane)
{
    const int u0 = t * kFxTile + 16 * lane;
    const long long ga = r.gb + u0;
    uint8_t* dst = r.ring + (u0 & (kFxRing - 1));
    if (lane >= kFxTile / 16) {
    } else if (ga >= r.lo && ga + 16 <= r.hi) {
        cp_async16(dst, reinterpret_cast<const void*>(ga));
    } else {
        unsigned w[4] = { 0, 0, 0, 0 };
        for (int k = 0; k < 16; ++k)
            if (ga + k >= r.lo && ga + k < r.hi) w[k >> 2] |= (unsigned)*reinterpret_cast<const uint8_t*>(ga + k) << (8 * (k & 3));
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    cp_async_commit();
}

// makes the bytes up to staging coordinate uHi readable (a step looks at most 38 bytes ahead of its first position, so it
// touches two tiles at most: the four slots hold those and the two tiles being fetched; a long match may carry the walk
// over a whole tile, the tiles it skips are fetched all the same)
__device__ __forceinline__ void fx_ensure(FxRing& r, int uHi, int lane)
{
    const int tNeed = uHi / kFxTile;
    if (tNeed > r.ready) {
        while (r.issued <= tNeed + kFxAhead) { fx_issue_tile(r, r.issued, lane); ++r.issued; }
        cp_async_wait_group<kFxAhead>();
        __syncwarp();
        r.ready = tNeed;
    }
}

__device__ __forceinline__ unsigned long long fx_load8(const FxRing& r, int u)
{
    const unsigned* w = reinterpret_cast<const unsigned*>(r.ring);
    const int i0 = (u >> 2) & (kFxRing / 4 - 1), i1 = (i0 + 1) & (kFxRing / 4 - 1), i2 = (i0 + 2) & (kFxRing / 4 - 1);
    const unsigned a = w[i0], b = w[i1], c = w[i2];
    const int sh = (u & 3) * 8;
    return (unsigned long long)__funnelshift_r(a, b, sh) | ((unsigned long long)__funnelshift_r(b, c, sh) << 32);
}

// number of equal low bytes of two 8-byte words, from their XOR (two 32-bit halves: the 64-bit __ffsll costs twice as much)
__device__ __forceinline__ int equal_bytes8(unsigned long long x)
{
    const unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
    if (lo) return (__ffs((int)lo) - 1) >> 3;
    if (hi) return 4 + ((__ffs((int)hi) - 1) >> 3);
    return 8;
}

constexpr int kFxQ = 2;
constexpr unsigned kFxLit = 0x80000000u, kFxMatch = 0x40000000u, kFxEob = 0x20000000u;

// barrier ids are immediates: with ids in registers ptxas reserves all 16 named barriers for the CTA, and the SM's barrier
// pool then admits four CTAs instead of twelve (measured: 24.8 instead of 32.7 GB/s on random bytes)
template <int ID> __device__ __forceinline__ void fx_bar_sync_c() { asm volatile("barrier.sync %0, 64;" ::"n"(ID) : "memory"); }
template <int ID> __device__ __forceinline__ void fx_bar_arrive_c() { asm volatile("barrier.arrive %0, 64;" ::"n"(ID) : "memory"); }
// slot s of the two-slot queue: barrier 1 + s = "full", barrier 3 + s = "empty" (s is warp-uniform).  barrier.sync is the
// non-aligned form (it tolerates a split warp), and it orders the memory accesses of the threads that take part in it --
// the walker's 32 lanes among them, so it also stands between a step's table commit and the next step's table reads.
__device__ __forceinline__ void fx_wait_full(unsigned s) { if (s == 0) fx_bar_sync_c<1>(); else fx_bar_sync_c<2>(); }
__device__ __forceinline__ void fx_wait_empty(unsigned s) { if (s == 0) fx_bar_sync_c<3>(); else fx_bar_sync_c<4>(); }
__device__ __forceinline__ void fx_post_full(unsigned s) { if (s == 0) fx_bar_arrive_c<1>(); else fx_bar_arrive_c<2>(); }
__device__ __forceinline__ void fx_post_empty(unsigned s) { if (s == 0) fx_bar_arrive_c<3>(); else fx_bar_arrive_c<4>(); }

// gload8 in two halves, so that the loads are in flight while the step does other things: the two aligned words now, the value later
struct Raw8 { unsigned long long x, y; int sh; };
__device__ __forceinline__ Raw8 gload8_issue(const uint8_t* p, const uint8_t* lo, const uint8_t* hi)
{
    Raw8 r;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint8_t* al = reinterpret_cast<const uint8_t*>(a & ~(uintptr_t)7);
    if (al >= lo && al + 16 <= hi) {
        r.x = __ldg(reinterpret_cast<const unsigned long long*>(al));
        r.y = __ldg(reinterpret_cast<const unsigned long long*>(al) + 1);
        r.sh = (int)(a & 7) * 8;
    } else {
        r.x = gload8(p, lo, hi); r.y = 0; r.sh = 0;
    }
    return r;
}
// the same for an address whose two aligned words are known to lie inside the stream
__device__ __forceinline__ Raw8 gload8_issue_inner(const uint8_t* p)
{
    Raw8 r;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned long long* al = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
    r.x = __ldg(al); r.y = __ldg(al + 1); r.sh = (int)(a & 7) * 8;
    return r;
}
__device__ __forceinline__ unsigned long long gload8_value(const Raw8& r) { return r.sh ? ((r.x >> r.sh) | (r.y << (64 - r.sh))) : r.x; }

// lanes whose 13-bit value equals mine (the result of __match_any_sync among the valid lanes), from 13 ballots
__device__ __forceinline__ unsigned fx_same_hash(unsigned h, bool valid, int lane)
{
    unsigned m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < kHashBits; ++b) {
        const unsigned bal = __ballot_sync(0xffffffffu, (h >> b) & 1u);
        m &= ((h >> b) & 1u) ? bal : ~bal;
    }
    return valid ? m : (1u << lane);
}

__global__ void __launch_bounds__(64) k_fixed(Job job)
{
    // The table holds positions modulo 65536 in 16 bits (16 KiB per chunk).  age = (i - entry) & 0xFFFF is exact while it
    // stays below 65536: a sweep every 8192 positions turns the entries that are more than 32768 behind (invalid from then
    // on, encoder.cpp:347) into "empty" ones of age 40000, so no entry ever gets older than ~48500 positions.
    __shared__ unsigned short table[kHashSize];
    __shared__ __align__(16) uint8_t ringMem[kFxRing];
    __shared__ unsigned queue[kFxQ * 32];
    __shared__ unsigned obuf[32];                 // a step emits <= 341 bits from bit (bitpos & 31) on: words 0..11 are used
    const unsigned slot = blockIdx.x;
    const Geom g = chunk_geom(job, slot);
    const int lane = threadIdx.x & 31;
    const bool walker = threadIdx.x < 32;
    const unsigned ltMask = (1u << lane) - 1u;
    constexpr int kEmptyAge = 40000, kSweep = 8192;
    for (int i = threadIdx.x; i < kHashSize; i += 64) table[i] = (unsigned short)(0 - kEmptyAge);
    if (threadIdx.x < 32) obuf[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* base = job.src + g.off;
    const uint8_t* lo = job.src - job.history;
    const uint8_t* hi = job.src + job.n;
    ChunkState& st = job.state[slot];
    const int n = g.body;

    if (walker) {
        // staging coordinate of chunk position q: u = q + uOff (>= 0 for the whole dictionary, 16-byte phase of the global address kept)
        const int uOff = (int)(reinterpret_cast<uintptr_t>(base) & 15) + kMaxDict;
        FxRing r;
        r.ring = ringMem;
        r.gb = (long long)reinterpret_cast<uintptr_t>(base) - uOff;
        r.lo = (long long)reinterpret_cast<uintptr_t>(lo); r.hi = (long long)reinterpret_cast<uintptr_t>(hi);
        r.issued = (uOff - g.dict) / kFxTile; r.ready = r.issued - 1;
        // every candidate of the chunk (and the aligned words around it) inside the stream: all chunks but the first and the last
        const bool inner = (long long)reinterpret_cast<uintptr_t>(base) - (kMaxDict + 8) >= r.lo && (long long)reinterpret_cast<uintptr_t>(base) + g.n + 24 <= r.hi;
        // level-1 priming convention: table[h(i+1)] = i for every dictionary position
        for (int i0 = -g.dict; i0 < 0; i0 += 32) {
            fx_ensure(r, i0 + uOff + 31 + 3, lane);
            const int i = i0 + lane;
            const bool v = i < 0;
            const unsigned h = hash3((unsigned)(fx_load8(r, i + uOff) >> 8) & 0xFFFFFFu);
            // the highest position of a hash group owns the slot: everybody writes; a lane that then finds a lower position of
            // this step in its slot writes again (the holder of a contested slot moves up every round)
            if (v) table[h] = (unsigned short)i;
            __syncwarp();
            bool again = v && (unsigned)((i - (int)table[h]) & 0xFFFF) - 1u < 31u;
            while (__any_sync(0xffffffffu, again)) {
                __syncwarp();
                if (again) table[h] = (unsigned short)i;
                __syncwarp();
                again = v && (unsigned)((i - (int)table[h]) & 0xFFFF) - 1u < 31u;
            }
            __syncwarp();
        }
        unsigned matches = 0, step = 0;
        if (n > 0) {
            int i0 = 0, nextSweep = kSweep;
            while (i0 < n) {
                if (i0 >= nextSweep) {
#pragma unroll 8
                    for (int k = lane; k < kHashSize; k += 32)
                        if (((i0 - (int)table[k]) & 0xFFFF) > kMaxDistance) table[k] = (unsigned short)(i0 - kEmptyAge);
                    nextSweep = i0 + kSweep;
                    __syncwarp();
                }
                fx_ensure(r, i0 + uOff + 31 + 7, lane);
                const int i = i0 + lane;
                const bool valid = i < n;
                const int remaining = n - i;
                const unsigned long long v8 = fx_load8(r, i + uOff);
                const unsigned h = hash3((unsigned)(v8 >> 8) & 0xFFFFFFu);
                const unsigned short old16 = table[h];
                const int dOld = valid ? ((i - (int)old16) & 0xFFFF) : 0xFFFF;
                const bool hasOld = dOld <= kMaxDistance;                     // unsigned(distance) <= maxDistance, encoder.cpp:347
                Raw8 cb; cb.x = cb.y = 0; cb.sh = 0;
                if (hasOld) cb = inner ? gload8_issue_inner(base + i - dOld) : gload8_issue(base + i - dOld, lo, hi);   // the candidates' bytes are on their way ...
                // ... while the lanes find out whether any two of them share a slot
                __syncwarp();
                if (valid) table[h] = (unsigned short)i;                      // as if every lane were visited
                __syncwarp();
                const bool shared = __any_sync(0xffffffffu, valid && table[h] != (unsigned short)i);
                unsigned grp = 1u << lane, lower = 0;
                int lowN = -1, mOld = 0, mLow = 0;
                if (shared) {
                    if (valid) table[h] = old16;                              // lanes of one slot found the same value there
                    grp = fx_same_hash(h, valid, lane);
                    lower = grp & ltMask;
                    lowN = lower ? 31 - __clz(lower) : -1;
                    const unsigned long long vl = __shfl_sync(0xffffffffu, v8, lowN < 0 ? lane : lowN);
                    if (valid && lowN >= 0) { mLow = equal_bytes8(v8 ^ vl); if (mLow > remaining) mLow = remaining; }
                }
                if (hasOld) {
                    mOld = equal_bytes8(v8 ^ gload8_value(cb));
                    if (mOld > remaining) mOld = remaining;                   // R2 clamp (short) / remain() clamp (long)
                }
                const unsigned packOld = (unsigned)mOld | ((unsigned)dOld << 8);
                const unsigned packLow = (unsigned)mLow | ((unsigned)(lane - lowN) << 8);
                // lanes whose outcome does not depend on the visited set and that start a match / lanes that have to be looked at
                const unsigned fixedAcc = __ballot_sync(0xffffffffu, valid && lower == 0 && mOld > 3);
                const unsigned depends = __ballot_sync(0xffffffffu, valid && lower != 0 && (mOld > 3 || mLow > 3 || (lower & (lower - 1)) != 0));
                // exact length of a match of >= 8 bytes that starts at lane f (remain(a, b, 8, n - i), encoder.cpp:352)
                auto exactLength = [&](int f, int D) {
                    const int fi = i0 + f;
                    const int maxLen = min(n - fi, kMaxMatch);
                    const unsigned long long xa = gload8(base + fi + 8 + lane * 8, lo, hi) ^ gload8(base + fi - D + 8 + lane * 8, lo, hi);
                    const unsigned mm = __ballot_sync(0xffffffffu, xa != 0);
                    int ext = 256;
                    if (mm) { const int src = __ffs(mm) - 1; const unsigned long long xs = __shfl_sync(0xffffffffu, xa, src); ext = src * 8 + ((__ffsll((long long)xs) - 1) >> 3); }
                    return min(8 + ext, maxLen);
                };
                // settle the step: p = first lane not decided yet, V = visited lanes, starts = lanes that start a match
                unsigned V = 0xffffffffu, starts = 0;
                int p = 32, adv = 32, myLen = 0, myDist = 0;
                bool done = true;
                if (fixedAcc | depends) {                                     // (no lane starts a match or has to be looked at: 32 literals)
                    // (1) In front of the first lane that has to be looked at (q) every outcome is known, so that part is settled in
                    // parallel: G = the first match start at or behind the end of my own match, the real match starts are the orbit of
                    // G from the first match start, collected by pointer doubling (three rounds cover the eight matches of a window).
                    const int q = depends ? __ffs(depends) - 1 : 32;
                    const unsigned belowQ = q >= 32 ? 0xffffffffu : ((1u << q) - 1u);
                    const unsigned accP = fixedAcc & belowQ;
                    if (accP) {
                        const int x = lane + mOld;
                        const unsigned behind = x >= 32 ? 0u : (accP & (0xffffffffu << x));
                        int G = (mOld == 8 || !behind) ? 32 : __ffs(behind) - 1;     // a match of >= 8 bytes ends the step
                        unsigned R = accP & (0u - accP);
                        for (;;) {
                            const unsigned add = __reduce_or_sync(0xffffffffu, (((R >> lane) & 1u) && G < 32) ? (1u << G) : 0u);
                            if ((add & ~R) == 0) break;
                            R |= add;
                            const int G2 = __shfl_sync(0xffffffffu, G, G & 31);
                            G = G >= 32 ? 32 : G2;
                        }
                        const bool inR = (R >> lane) & 1u;
                        const unsigned covered = __reduce_or_sync(0xffffffffu, inR ? (((1u << mOld) - 2u) << lane) : 0u);
                        if (inR) { myLen = mOld; myDist = dOld; }
                        starts = R;
                        const int last = 31 - __clz(R);
                        const unsigned pkLast = __shfl_sync(0xffffffffu, packOld, last);
                        int L = (int)(pkLast & 0xFFu);
                        if (L == 8) {
                            L = exactLength(last, (int)(pkLast >> 8));
                            if (lane == last) myLen = L;
                            adv = last + L; done = true; p = 32;
                            V = ~covered & ((2u << last) - 1u);
                        } else if (last + L >= 32) {
                            adv = last + L; done = true; p = 32;
                            V = ~covered;
                        } else {
                            p = max(last + L, q);
                            V = ~covered & (p >= 32 ? 0xffffffffu : ((1u << p) - 1u));
                            done = p >= 32;
                        }
                    } else {
                        p = q; V = belowQ; done = p >= 32;
                    }
                }
                // (2) the rest of the window, one match (or one lane that has to be looked at) per turn
                while (!done) {
                    const unsigned fromP = 0xffffffffu << p;
                    const unsigned stop = (fixedAcc | depends) & fromP;
                    if (!stop) { V |= fromP; break; }
                    const int f = __ffs(stop) - 1;
                    V |= fromP & ((1u << f) - 1u);                          // the lanes in between are literals
                    unsigned pk;
                    if ((depends >> f) & 1u) {
                        const unsigned elig = lower & V;                     // visited lower lanes of my group (V is complete below f)
                        const unsigned mine = !elig ? packOld : (31 - __clz(elig) == lowN ? packLow : 0x80000000u);
                        pk = __shfl_sync(0xffffffffu, mine, f);
                    } else {
                        pk = __shfl_sync(0xffffffffu, packOld, f);
                    }
                    if (pk & 0x80000000u) { adv = f; break; }                // its candidate is a lower lane that is not the nearest: next step
                    V |= 1u << f;
                    int L = (int)(pk & 0xFFu);
                    const int D = (int)(pk >> 8);
                    if (L <= 3) { p = f + 1; if (p >= 32) break; continue; }  // a literal after all
                    starts |= 1u << f;
                    const bool isLong = L == 8;
                    if (isLong) L = exactLength(f, D);
                    if (lane == f) { myLen = L; myDist = D; }
                    p = f + L;
                    if (isLong || p >= 32) { adv = p; break; }
                }
                // commit: the highest visited lane of a hash group owns the slot
                __syncwarp();
                const bool vis = valid && ((V >> lane) & 1u);
                if (shared) { if (vis && (grp & ~ltMask & ~(1u << lane) & V) == 0) table[h] = (unsigned short)i; }
                else if (valid && !vis) table[h] = old16;
                // hand the step to the emitter
                unsigned w = 0;
                if (vis) w = ((starts >> lane) & 1u) ? (kFxMatch | ((unsigned)myLen << 15) | (unsigned)(myDist - 1)) : (kFxLit | (unsigned)(v8 & 0xFF));
                const unsigned qs = step & (kFxQ - 1);
                fx_wait_empty(qs);
                queue[qs * 32 + lane] = w;
                fx_post_full(qs);
                ++step;
                matches += __popc(starts);
                i0 += adv;
            }
            const unsigned qs = step & (kFxQ - 1);
            fx_wait_empty(qs);
            queue[qs * 32 + lane] = kFxEob;
            fx_post_full(qs);
        }
        if (lane == 0) st.ntok = matches;
        return;
    }

    // ---- emitter ----
    unsigned* out32 = reinterpret_cast<unsigned*>(job.cand + (size_t)slot * job.chunk);
    unsigned long long bitpos = 0;           // bits of the chunk emitted so far (warp-uniform)
    // the step's bits sit in obuf from bit (bitpos & 31) on: full words leave for the scratch slot, the rest is carried
    auto stepFlush = [&](unsigned stepBits) {
        __syncwarp();
        const unsigned startBit = (unsigned)(bitpos & 31);
        const unsigned full = (startBit + stepBits) >> 5;
        const unsigned wbase = (unsigned)(bitpos >> 5);
        const unsigned mine = obuf[lane];
        if ((unsigned)lane < full) out32[wbase + lane] = mine;
        const unsigned carry = __shfl_sync(0xffffffffu, mine, full);
        obuf[lane] = lane == 0 ? carry : 0u;                          // every lane resets the word it has just read
        bitpos += stepBits;
        __syncwarp();
    };
    auto putAt = [&](unsigned off, unsigned bits, int nb) {          // off relative to bitpos, nb <= 32
        const unsigned o = (unsigned)(bitpos & 31) + off;
        const unsigned long long v = (unsigned long long)bits << (o & 31);
        atomicOr(&obuf[o >> 5], (unsigned)v);
        if ((o & 31) + nb > 32) atomicOr(&obuf[(o >> 5) + 1], (unsigned)(v >> 32));
    };
    if (n > 0) {
        if (lane == 0) putAt(0, (g.final ? 1u : 0u) | (1u << 1), 3);     // StartBlock(FixedHuffman, final)
        stepFlush(3);
        fx_post_empty(0); fx_post_empty(1);                              // both slots start empty
        for (unsigned step = 0;; ++step) {
            const unsigned qs = step & (kFxQ - 1);
            fx_wait_full(qs);
            const unsigned w = queue[qs * 32 + lane];
            fx_post_empty(qs);
            unsigned bits1 = 0, bits2 = 0; int n1 = 0, n2 = 0;
            if (w & kFxLit) {
                bits1 = fixed_literal_code(w & 0xFFu, n1);
            } else if (w & kFxMatch) {
                int eb, ev, cl;
                const int ls = len_symbol((int)((w >> 15) & 0x1FFu), eb, ev);
                const unsigned lc = fixed_lit_code((unsigned)ls, cl);
                bits1 = lc | ((unsigned)ev << cl); n1 = cl + eb;
                const int ds = dist_symbol((int)(w & 0x7FFFu) + 1, eb, ev);
                bits2 = (__brev((unsigned)ds) >> 27) | ((unsigned)ev << 5); n2 = 5 + eb;
            } else if ((w & kFxEob) && lane == 0) {
                bits1 = fixed_lit_code(256u, n1);
            }
            unsigned off, stepBits;
            if (__ballot_sync(0xffffffffu, (w & (kFxMatch | kFxEob)) != 0) == 0) {
                // literals only (8 or 9 bits each): the offsets come from two ballots
                const unsigned lit = __ballot_sync(0xffffffffu, n1 != 0), nine = __ballot_sync(0xffffffffu, n1 == 9);
                off = 8u * __popc(lit & ltMask) + __popc(nine & ltMask);
                stepBits = 8u * __popc(lit) + __popc(nine);
            } else {
                unsigned incl = (unsigned)(n1 + n2);
                for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                off = incl - (unsigned)(n1 + n2);
                stepBits = __shfl_sync(0xffffffffu, incl, 31);
            }
            if (n1) putAt(off, bits1, n1);
            if (n2) putAt(off + n1, bits2, n2);
            stepFlush(stepBits);
            if (w & kFxEob) break;                                       // every lane of the last step carries the flag
        }
    }
    const unsigned long long q = bitpos;
    unsigned bytes = (unsigned)((q + 7) >> 3);
    if (!g.final) {
        const unsigned pad = (unsigned)((8 - ((q + 3) & 7)) & 7);
        stepFlush(3 + pad);                                              // stored block header (not final) + padding: zeros
        if (lane == 0) putAt(0, 0xFFFE0001u, 32);                        // LEN = 1, NLEN = 0xFFFE
        stepFlush(32);
        const unsigned lastByte = (unsigned)(gload8(base + g.n - 1, lo, hi) & 0xFF);
        if (lane == 0) putAt(0, lastByte, 8);
        stepFlush(8);
        bytes = (unsigned)(bitpos >> 3);
    }
    {   // final partial word
        __syncwarp();
        if (lane == 0 && (bitpos & 31)) out32[bitpos >> 5] = obuf[0];
    }
    if (lane == 0) { st.block_type = 1; st.hdr_bits = 0; st.total_bits = q; st.out_bytes = bytes; }
}

__global__ void __launch_bounds__(256) k_gather(Job job)
{
    const unsigned slot = blockIdx.x;
    const ChunkState st = job.state[slot];
    if (st.out_off + st.out_bytes > job.cap) {
        if (threadIdx.x == 0) atomicOr(reinterpret_cast<unsigned long long*>(&job.total[1]), 1ull);
        return;
    }
    const unsigned* src32 = reinterpret_cast<const unsigned*>(job.cand + (size_t)slot * job.chunk);
    copy_out(job.dst + st.out_off, src32, st.out_bytes);
}

// ------------------------------------------------------------------------------------------------
// K-CKSUM : per-chunk Adler-32 (start 0) and CRC-32
// ------------------------------------------------------------------------------------------------
__constant__ uint32_t c_crcTable[4][256];   // slicing-by-4 tables of the reflected polynomial 0xEDB88320 (crc.cpp:5-22 is table 0)
__constant__ uint32_t c_powL[257];      // x^(8*256*q) mod P, reflected (q <= 256: a full 64 KiB chunk)
__constant__ uint32_t c_pow1[256];      // x^(8*r) mod P, reflected

__host__ __device__ inline uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1) ? 0xEDB88320u : 0);
    }
    return p;
}

constexpr int kCkThreads = 256;
constexpr int kCkSlice = 256;

// Each thread owns a 256-byte slice: Adler partial sums with dp4a (s1 = sum d, s2 = sum (len-i) d_i), raw CRC
// with slicing-by-4, then the slice CRC is multiplied by x^(8 * bytes after the slice) and everything is
// XOR-/sum-reduced.  The chunk's standard CRC adds the propagated 0xFFFFFFFF preset and the final inversion.
// kWant: bit 0 Adler-32, bit 1 CRC-32 (the Zlib trailer needs only the first, the Gzip trailer only the second; the CRC's
// table look-ups are the expensive half)
// Scratch of checksum_chunk: 4 KiB + 160 B of shared memory (static in K-CKSUM / K-STORED, a dead part of the dynamic
// allocation in K-LZ).  Works with any CTA of >= 256 threads: threads from 256 on have no slice.
struct CkShared { uint32_t tab[4][256]; unsigned long long redA[8], redB[8]; uint32_t redC[8]; };

template <int kWant>
__device__ __forceinline__ void checksum_chunk(const Job& job, unsigned slot, const Geom& g, CkShared& cs)
{
    uint32_t (*tab)[256] = cs.tab;
    unsigned long long* redA = cs.redA; unsigned long long* redB = cs.redB; uint32_t* redC = cs.redC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (kWant & 2) {
        if (tid < 256) for (int k = 0; k < 4; ++k) tab[k][tid] = c_crcTable[k][tid];
        __syncthreads();
    }
    const uint8_t* p = job.src + g.off;
    const int lo = tid * kCkSlice;
    int hi = lo + kCkSlice; if (hi > g.n) hi = g.n;
    unsigned long long s1 = 0, s2 = 0; uint32_t crc = 0;
    if (lo < hi) {
        const int len = hi - lo;
        unsigned a = 0, b = 0;                                       // slice-local: a = sum d, b = sum (len - i) d_i  (< 2^24)
        int i = 0;
        auto word = [&](unsigned w, int remaining) {                  // 4 bytes at offset i, `remaining` = len - i
            // sum_{k<4} (remaining - k) d_k = remaining * sum d_k - (0 d0 + 1 d1 + 2 d2 + 3 d3)
            if (kWant & 1) {
                const unsigned sum = __dp4a(w, 0x01010101u, 0u);
                b += (unsigned)remaining * sum - __dp4a(w, 0x03020100u, 0u);
                a += sum;
            }
            if (kWant & 2) {
                crc ^= w;
                crc = tab[3][crc & 0xFF] ^ tab[2][(crc >> 8) & 0xFF] ^ tab[1][(crc >> 16) & 0xFF] ^ tab[0][crc >> 24];
            }
        };
        const uint8_t* q = p + lo;
        if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
            for (; i + 16 <= len; i += 16) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(q + i));
                word(v.x, len - i); word(v.y, len - i - 4); word(v.z, len - i - 8); word(v.w, len - i - 12);
            }
        } else if ((reinterpret_cast<uintptr_t>(q) & 3) == 0) {
            for (; i + 4 <= len; i += 4) word(__ldg(reinterpret_cast<const unsigned*>(q + i)), len - i);
        }
        for (; i < len; ++i) {
            const unsigned v = q[i];
            if (kWant & 1) { a += v; b += (unsigned)(len - i) * v; }
            if (kWant & 2) crc = (crc >> 8) ^ tab[0][(crc ^ v) & 0xFF];
        }
        const int after = g.n - hi;
        s1 = a; s2 = (unsigned long long)b + (unsigned long long)after * a;      // b_chunk = sum (n - i) d[i]
        if (kWant & 2) crc = gf2_mulmod(gf2_mulmod(crc, c_powL[after >> 8]), c_pow1[after & 255]);
    }
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
    }
    if (lane == 0 && warp < kCkThreads / 32) { redA[warp] = s1; redB[warp] = s2; redC[warp] = crc; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0, b = 0; uint32_t c = 0;
        for (int w = 0; w < kCkThreads / 32; ++w) { a += redA[w]; b += redB[w]; c ^= redC[w]; }
        uint32_t* ck = job.ck + 2 * (job.first_chunk + slot);
        ck[0] = (uint32_t)(((b % 65521ull) << 16) | (a % 65521ull));
        // standard CRC = raw(M) ^ (0xFFFFFFFF * x^(8n)) ^ 0xFFFFFFFF
        if (kWant & 2) {
            const uint32_t init = gf2_mulmod(gf2_mulmod(0xFFFFFFFFu, c_powL[g.n >> 8]), c_pow1[g.n & 255]);
            ck[1] = c ^ init ^ 0xFFFFFFFFu;
        } else ck[1] = 0;
    }
}

// Adler-32 partial (start 0) of the chunk by all threads of a K-LZ CTA, in the tail of the kernel: coalesced 16-byte
// loads, several in flight per thread (the slice layout of checksum_chunk needs ~19 us of latency per chunk, far too
// long for a tail).  With b = sum (n - i) d_i:  a vector at offset o contributes (n - o) S - W, S = sum of its bytes,
// W = sum j d_j over its 16 bytes.  Only the Adler sum is fused (zlib framing); a wanted CRC keeps its own launch.
__device__ __forceinline__ void lz_checksums(const uint8_t* p, int n, uint32_t* ck, uint8_t* scratch)
{
    // (inlined, plain values: as a called function it cost K-LZ 0.9 ms per GiB, and a reference to the Job moved the kernel's
    // parameters to the stack)
    unsigned long long* redA = reinterpret_cast<unsigned long long*>(scratch);
    unsigned long long* redB = redA + 32;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    int head = (int)((16u - (unsigned)(reinterpret_cast<uintptr_t>(p) & 15)) & 15u); if (head > n) head = n;
    const int nvec = (n - head) >> 4;
    const int tail0 = head + nvec * 16;
    unsigned long long A = 0, B = 0;
    if (tid < head) { const unsigned v = p[tid]; A += v; B += (unsigned long long)(n - tid) * v; }
    if (tid < n - tail0) { const int i = tail0 + tid; const unsigned v = p[i]; A += v; B += (unsigned long long)(n - i) * v; }
    const uint4* pv = reinterpret_cast<const uint4*>(p + head);
    unsigned a32 = 0; unsigned long long b64 = 0;
#pragma unroll 4
    for (int v = tid; v < nvec; v += nthr) {
        const uint4 q = __ldg(pv + v);
        const unsigned s0 = __dp4a(q.x, 0x01010101u, 0u), s1 = __dp4a(q.y, 0x01010101u, 0u), s2 = __dp4a(q.z, 0x01010101u, 0u), s3 = __dp4a(q.w, 0x01010101u, 0u);
        const unsigned S = s0 + s1 + s2 + s3;
        const unsigned W = __dp4a(q.x, 0x03020100u, 0u) + __dp4a(q.y, 0x03020100u, 0u) + __dp4a(q.z, 0x03020100u, 0u) + __dp4a(q.w, 0x03020100u, 0u)
                           + 4u * s1 + 8u * s2 + 12u * s3;
        a32 += S;
        b64 += (unsigned long long)(unsigned)(n - head - 16 * v) * S - W;
    }
    A += a32; B += b64;
    for (int o = 16; o; o >>= 1) { A += __shfl_xor_sync(0xffffffffu, A, o); B += __shfl_xor_sync(0xffffffffu, B, o); }
    if (lane == 0) { redA[warp] = A; redB[warp] = B; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0, b = 0;
        for (int w = 0; w < nthr / 32; ++w) { a += redA[w]; b += redB[w]; }
        ck[0] = (uint32_t)(((b % 65521ull) << 16) | (a % 65521ull));
        ck[1] = 0;
    }
}

template <int kWant>
__global__ void __launch_bounds__(kCkThreads) k_checksums(Job job)
{
    __shared__ CkShared cs;
    const unsigned slot = blockIdx.x;
    checksum_chunk<kWant>(job, slot, chunk_geom(job, slot), cs);
}

// ------------------------------------------------------------------------------------------------
// K-STORED (level 0) : WriteUncompressedBlock (encoder.cpp:482-502) for every chunk plus, in the same kernel, the
// chunk's checksum partials.  At level 0 every size is known beforehand (5 bytes per block of <= 65535 bytes, the
// aligning 1-byte block behind every non-final chunk), so chunk c starts at c * size(full chunk): no
// Huffman stage, no offset scan, and the input is read from HBM once (the checksum pass finds it in cache).
// ------------------------------------------------------------------------------------------------
template <int kWant>
__global__ void __launch_bounds__(kCkThreads) k_stored(Job job)
{
    const unsigned slot = blockIdx.x;
    const Geom g = chunk_geom(job, slot);
    const int tid = threadIdx.x;
    const unsigned long long perFull = (unsigned long long)stored_size((int)job.chunk - 1) + 6ull;
    const unsigned long long off = (job.first_chunk + slot) * perFull;
    const uint32_t bytes = stored_size(g.body) + (g.final ? 0u : 6u);
    if (off + bytes > job.cap) {                                      // never write past the caller's buffer
        if (tid == 0) atomicOr(reinterpret_cast<unsigned long long*>(&job.total[1]), 1ull);
        return;
    }
    const uint8_t* chunk0 = job.src + g.off;
    uint8_t* Dst = job.dst + off;
    int written = 0; unsigned o = 0;
    while (written < g.body) {
        const int len = min(g.body - written, 0xFFFF);
        if (tid == 0) {
            Dst[o] = (uint8_t)((g.final && written + len == g.body) ? 1 : 0);
            Dst[o + 1] = (uint8_t)len; Dst[o + 2] = (uint8_t)(len >> 8);
            Dst[o + 3] = (uint8_t)~len; Dst[o + 4] = (uint8_t)((~len) >> 8);
        }
        copy_g2g(Dst + o + 5, chunk0 + written, (unsigned)len);
        o += 5 + len; written += len;
    }
    if (!g.final && tid == 0) {
        Dst[o] = 0; Dst[o + 1] = 1; Dst[o + 2] = 0; Dst[o + 3] = 0xFE; Dst[o + 4] = 0xFF; Dst[o + 5] = chunk0[g.n - 1];
    }
    if (tid == 0) {
        ChunkState& st = job.state[slot];
        st.ntok = 0; st.block_type = 0; st.hdr_bits = 0; st.total_bits = 0; st.out_bytes = bytes; st.out_off = off;
        if (slot == job.nchunks - 1) { job.total[0] = off + bytes; job.total[3] += job.nchunks; }
    }
    if (kWant) { __shared__ CkShared cs; checksum_chunk<(kWant ? kWant : 3)>(job, slot, g, cs); }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side: launch wrappers and checksum folds
// ------------------------------------------------------------------------------------------------
cudaError_t configure_kernels()
{
    cudaError_t e;
    e = cudaFuncSetAttribute(k_huffman_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, kHuffLSmem); if (e) return e;
    e = cudaFuncSetAttribute(k_emit2, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmit2Smem); if (e) return e;
    e = cudaFuncSetAttribute(k_lz, cudaFuncAttributeMaxDynamicSharedMemorySize, kLzSmem); if (e) return e;
    static uint32_t tab[4][256];
    uint32_t powL[257], pow1[256];
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int j = 0; j < 8; ++j) c = (c >> 1) ^ ((c & 1) * 0xEDB88320u);
        tab[0][i] = c;
    }
    for (int k = 1; k < 4; ++k)
        for (uint32_t i = 0; i < 256; ++i) tab[k][i] = (tab[k - 1][i] >> 8) ^ tab[0][tab[k - 1][i] & 0xFF];
    const uint32_t x8 = 0x00800000u;                                  // x^8
    uint32_t xL = 0x80000000u;                                        // x^(8*256) by repeated multiplication
    for (int i = 0; i < 256; ++i) xL = gf2_mulmod(xL, x8);
    pow1[0] = 0x80000000u; powL[0] = 0x80000000u;
    for (int i = 1; i < 256; ++i) pow1[i] = gf2_mulmod(pow1[i - 1], x8);
    for (int i = 1; i < 257; ++i) powL[i] = gf2_mulmod(powL[i - 1], xL);
    e = cudaMemcpyToSymbol(c_crcTable, tab, sizeof tab); if (e) return e;
    e = cudaMemcpyToSymbol(c_powL, powL, sizeof powL); if (e) return e;
    e = cudaMemcpyToSymbol(c_pow1, pow1, sizeof pow1); if (e) return e;
    return cudaSuccess;
}

int launch_candidates(const Job& job, cudaStream_t s)
{
    // one priming pass per run of chunks is exact only for the default geometry (see K-CAND)
    int run = 1;
    if (job.chunk == 65536 && job.dict == 32768) {
        const unsigned target = 148u * 6u;                        // resident warps: 32 KiB table + staging per CTA
        run = (int)((job.nchunks + target - 1) / target);
        if (run < 1) run = 1;
        if (run > 64) run = 64;
    }
    k_candidates<<<(job.nchunks + run - 1) / run, 32, 0, s>>>(job, run);
    return 1;
}

int launch_huffman(const Job& job, cudaStream_t s)
{
    // resident warps of the warp-per-chunk kernel: 13 CTAs of 4 warps per SM (17 KiB of shared memory each)
    if (job.nchunks <= 148u * 52u && !job.mode) k_huffman<<<(job.nchunks + kHuffWarps - 1) / kHuffWarps, kHuffThreads, 0, s>>>(job);
    else k_huffman_lanes<<<(job.nchunks + kHuffLThreads - 1) / kHuffLThreads, kHuffLThreads, kHuffLSmem, s>>>(job);
    return 1;
}

int launch_offsets(const Job& job, cudaStream_t s)
{
    k_offsets<<<1, 1024, 0, s>>>(job);
    return 1;
}

// A/B switches of kernel variants (zzgpu_set_option); none changes the produced bytes.
static int g_optTma = 1;         // K-LZ window: 1 = cp.async.bulk (TMA) + mbarrier, 0 = LDG.128 -> STS
static int g_optSpec = 1;        // K-LZ walk: 1 = speculative per-lane chains merged by the true walk, 0 = the true walk alone
static int g_optRegion = 1;      // K-LZ histograms: 1 = counted beside the walks region by region, 0 = all at the end of the chunk
bool set_kernel_option(const char* name, int value)
{
    if (!strcmp(name, "tma")) { g_optTma = value ? 1 : 0; return true; }
    if (!strcmp(name, "spec")) { g_optSpec = value ? 1 : 0; return true; }
    if (!strcmp(name, "region")) { g_optRegion = value ? 1 : 0; return true; }
    return false;
}
int launch_lz(const Job& job, cudaStream_t s)
{
    k_lz<<<job.nchunks, kLzThreads, kLzSmem, s>>>(job, g_optTma, g_optSpec, g_optRegion);
#ifdef ZZ_PHASE_TIMING
    {
        cudaStreamSynchronize(s);
        unsigned long long h[12] = { 0 }; cudaMemcpyFromSymbol(h, g_lzPhase, sizeof h);
        double tot = 0; for (int k = 0; k < 12; ++k) tot += (double)h[k];
        static const char* names[12] = { "window", "firstprobe", "A", "preC", "chains", "links", "walk", "waitC", "D", "flush", "", "" };
        fprintf(stderr, "K-LZ phases (share of CTA time, tid 0):");
        for (int k = 0; k < 10; ++k) fprintf(stderr, " %s %.1f%%", names[k], 100.0 * (double)h[k] / (tot > 0 ? tot : 1));
        fprintf(stderr, "  | cycles per chunk %.0f\n", tot / (double)job.nchunks);
        unsigned long long z[12] = { 0 }; cudaMemcpyToSymbol(g_lzPhase, z, sizeof z);
    }
#endif
#ifdef ZZ_LZ_CHECKS
    {   // diagnostic build: bounds and loop guards inside K-LZ report the first violation
        const cudaError_t e = cudaStreamSynchronize(s);
        unsigned h[8] = { 0 };
        cudaMemcpyFromSymbol(h, g_lzDebug, sizeof h);
        if (e != cudaSuccess || h[0]) fprintf(stderr, "K-LZ check: cuda=%d code=%u a=%d b=%d block=%u thread=%u\n", (int)e, h[0], (int)h[1], (int)h[2], h[3], h[4]);
        unsigned z[8] = { 0 }; cudaMemcpyToSymbol(g_lzDebug, z, sizeof z);
    }
#endif
    return 1;
}

int launch_emit(const Job& job, cudaStream_t s)
{
    k_emit2<<<job.nchunks, kEmit2Threads, kEmit2Smem, s>>>(job);
    return 1;
}

int launch_checksums(const Job& job, cudaStream_t s)
{
    const int want = job.want_checksums & 3;
    if (want == 1) k_checksums<1><<<job.nchunks, kCkThreads, 0, s>>>(job);
    else if (want == 2) k_checksums<2><<<job.nchunks, kCkThreads, 0, s>>>(job);
    else k_checksums<3><<<job.nchunks, kCkThreads, 0, s>>>(job);
    return 1;
}

int launch_stored(const Job& job, cudaStream_t s)
{
    switch (job.want_checksums & 3) {
    case 0: k_stored<0><<<job.nchunks, kCkThreads, 0, s>>>(job); break;
    case 1: k_stored<1><<<job.nchunks, kCkThreads, 0, s>>>(job); break;
    case 2: k_stored<2><<<job.nchunks, kCkThreads, 0, s>>>(job); break;
    default: k_stored<3><<<job.nchunks, kCkThreads, 0, s>>>(job); break;
    }
    return 1;
}

int launch_fixed(const Job& job, cudaStream_t s)
{
    k_fixed<<<job.nchunks, 64, 0, s>>>(job);
    return 1;
}

int launch_gather(const Job& job, cudaStream_t s)
{
    k_gather<<<job.nchunks, 256, 0, s>>>(job);
    return 1;
}

uint32_t adler32_combine(uint32_t first, uint32_t second, size_t lenSecond)      // adler.cpp:5-15
{
    const uint64_t MOD = 65521;
    uint64_t a = (first & 0xFFFF) + (second & 0xFFFF);
    uint64_t b = (first >> 16) + (second >> 16);
    b += (uint64_t)(lenSecond % MOD) * (first & 0xFFFF);
    return (uint32_t)(((b % MOD) << 16) | (a % MOD));
}

uint32_t crc32_shift_operator(uint64_t len)                        // x^(8*len) mod P, reflected
{
    uint32_t xp = 0x80000000u, sq = 0x00800000u;
    for (uint64_t k = len; k; k >>= 1) {
        if (k & 1) xp = gf2_mulmod(xp, sq);
        sq = gf2_mulmod(sq, sq);
    }
    return xp;
}

uint32_t crc32_apply_shift(uint32_t crc, uint32_t op) { return gf2_mulmod(crc, op); }

uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
    if (len2 == 0) return crc1;
    return gf2_mulmod(crc1, crc32_shift_operator(len2)) ^ crc2;
}

}  // namespace zz
