/* TEST INFRASTRUCTURE ONLY -- see zz_oracle.h for scope, parity status and defect policies.
 *
 * Restatement of the zzflate encode path in plain C.  References are to /root/reference/zzflate/.
 */
#include "zz_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define HASH_BITS 13
#define HASH_SIZE (1 << HASH_BITS)
#define MAX_DISTANCE 0x8000          /* encoder.h:46 */
#define MAX_LENGTH 258               /* encoder.h:47 */
#define MAX_RECORDS 20000            /* encoder.h:44 */
#define BATCH 16384                  /* encoder.cpp:227 */
#define BLOCK_CAP 500000             /* encoder.cpp:518 */
#define EMPTY_SLOT (-100000)         /* encoder.cpp:533-536 */

/* ------------------------------------------------------------------------------------------------
 * static tables (luts.cpp:5-116, fixedhuffmanluts.cpp:5-55), regenerated from RFC 1951 3.2.5/3.2.6
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t length; uint32_t bits; } zcode;          /* outputbitstream.h:14-24 */

static int16_t  g_lenCode[259];
static int8_t   g_lenExtra[259];
static int8_t   g_lenExtraBits[259];
static const uint8_t g_order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
static uint8_t  g_extraDist[30];
static uint8_t  g_extraLen[286];
static uint16_t g_distBase[30];
static uint8_t  g_distLut[32769];
static zcode    g_codesF[286], g_lcodesF[259], g_dcodesF[30];
static uint32_t g_crcTable[256];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* huffman.cpp:11-33 */
unsigned zzo_reverse(unsigned value, int len)
{
    unsigned r = 0;
    for (int i = 0; i < len; ++i)
        r |= ((value >> i) & 1u) << (len - 1 - i);
    return r;
}

/* huffman.cpp:35-51 */
void zzo_default_table_lengths(int* out)
{
    for (int i = 0; i < 288; ++i)
        out[i] = (i <= 143 || i >= 280) ? 8 : (i <= 255 ? 9 : 7);
}

/* huffman.h:49-81 : canonical codes, stored bit-reversed; zero-length entries untouched */
static void generate(const int* lengths, int n, zcode* codes)
{
    int bl_count[16] = { 0 };
    for (int i = 0; i < n; ++i) bl_count[lengths[i]]++;
    unsigned next_code[16] = { 0 };
    unsigned bits = 0;
    bl_count[0] = 0;
    for (int b = 1; b < 16; ++b) {
        bits = (bits + (unsigned)bl_count[b - 1]) << 1;
        next_code[b] = bits;
    }
    for (int i = 0; i < n; ++i) {
        int len = lengths[i];
        if (len <= 0) continue;
        codes[i].length = len;
        codes[i].bits = zzo_reverse(next_code[len], len);
        next_code[len]++;
    }
}

static zcode merge(zcode first, zcode second)                       /* encoder.cpp:121-124 */
{
    zcode r;
    r.length = first.length + second.length;
    r.bits = (second.bits << first.length) | first.bits;
    return r;
}

static void create_merged_length_codes(zcode* lcodes, const zcode* symbolCodes)   /* encoder.cpp:126-133 */
{
    for (int i = 0; i < 259; ++i) {
        zcode extra = { g_lenExtraBits[i], (uint32_t)g_lenExtra[i] };
        lcodes[i] = merge(symbolCodes[g_lenCode[i]], extra);
    }
}

static void init_tables(void)
{
    /* length symbols 257..285 */
    static const int lbase[29] = { 3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258 };
    static const int lext[29]  = { 0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0 };
    for (int i = 0; i < 3; ++i) { g_lenCode[i] = 0; g_lenExtra[i] = 0; g_lenExtraBits[i] = 0; }
    for (int len = 3; len <= 258; ++len) {
        int s = 28;
        if (len < 258) { s = 0; while (s < 27 && lbase[s + 1] <= len) ++s; }
        g_lenCode[len] = (int16_t)(257 + s);
        g_lenExtra[len] = (int8_t)(len - lbase[s]);
        g_lenExtraBits[len] = (int8_t)lext[s];
    }
    for (int i = 0; i < 286; ++i) g_extraLen[i] = (uint8_t)(i < 257 ? 0 : lext[i - 257]);
    /* distance symbols 0..29 */
    int base = 1;
    for (int s = 0; s < 30; ++s) {
        int eb = s < 4 ? 0 : (s - 2) / 2;
        g_extraDist[s] = (uint8_t)eb;
        g_distBase[s] = (uint16_t)base;
        base += 1 << eb;
    }
    g_distLut[0] = 255;
    for (int d = 1; d <= 32768; ++d) {
        int s = 29;
        while (g_distBase[s] > d) --s;
        g_distLut[d] = (uint8_t)s;
    }
    /* fixed Huffman tables */
    int l288[288];
    zcode c288[288];
    zzo_default_table_lengths(l288);
    memset(c288, 0, sizeof c288);
    generate(l288, 288, c288);
    memcpy(g_codesF, c288, sizeof g_codesF);
    create_merged_length_codes(g_lcodesF, g_codesF);
    for (int i = 0; i < 30; ++i) { g_dcodesF[i].length = 5; g_dcodesF[i].bits = zzo_reverse((unsigned)i, 5); }
    /* crc.cpp:5-22 */
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t crc = i;
        for (int j = 0; j < 8; ++j) crc = (crc >> 1) ^ ((crc & 1) * 0xEDB88320u);
        g_crcTable[i] = crc;
    }
}

static void ensure_tables(void) { pthread_once(&g_once, init_tables); }

void zzo_tables(int16_t* lengthCode259, int8_t* lengthExtra259, int8_t* lengthExtraBits259,
                uint8_t* order19, uint8_t* extraDist30, uint8_t* extraLen286, uint16_t* distBase30,
                int32_t* codesF286x2, int32_t* lcodesF259x2, int32_t* dcodesF30x2)
{
    ensure_tables();
    for (int i = 0; i < 259; ++i) {
        lengthCode259[i] = g_lenCode[i]; lengthExtra259[i] = g_lenExtra[i]; lengthExtraBits259[i] = g_lenExtraBits[i];
        lcodesF259x2[2 * i] = g_lcodesF[i].length; lcodesF259x2[2 * i + 1] = (int32_t)g_lcodesF[i].bits;
    }
    for (int i = 0; i < 19; ++i) order19[i] = g_order[i];
    for (int i = 0; i < 30; ++i) {
        extraDist30[i] = g_extraDist[i]; distBase30[i] = g_distBase[i];
        dcodesF30x2[2 * i] = g_dcodesF[i].length; dcodesF30x2[2 * i + 1] = (int32_t)g_dcodesF[i].bits;
    }
    for (int i = 0; i < 286; ++i) {
        extraLen286[i] = g_extraLen[i];
        codesF286x2[2 * i] = g_codesF[i].length; codesF286x2[2 * i + 1] = (int32_t)g_codesF[i].bits;
    }
}

int zzo_find_distance(int offset)                                   /* encoder.cpp:51-61 */
{
    ensure_tables();
    for (int n = 1; n < 30; ++n)
        if (offset < g_distBase[n]) return n - 1;
    return offset <= 32768 ? 29 : -1;
}

int zzo_read_lut(int offset) { ensure_tables(); return g_distLut[offset]; }

void zzo_merged_length_codes(const int32_t* symbolCodes286x2, int32_t* lcodes259x2)
{
    ensure_tables();
    zcode sym[286], l[259];
    for (int i = 0; i < 286; ++i) { sym[i].length = symbolCodes286x2[2 * i]; sym[i].bits = (uint32_t)symbolCodes286x2[2 * i + 1]; }
    create_merged_length_codes(l, sym);
    for (int i = 0; i < 259; ++i) { lcodes259x2[2 * i] = l[i].length; lcodes259x2[2 * i + 1] = (int32_t)l[i].bits; }
}

void zzo_generate(const int* lengths, int n, int32_t* codesOut)
{
    zcode* c = (zcode*)calloc((size_t)n, sizeof(zcode));
    generate(lengths, n, c);
    for (int i = 0; i < n; ++i) { codesOut[2 * i] = c[i].length; codesOut[2 * i + 1] = (int32_t)c[i].bits; }
    free(c);
}

/* encoder.cpp:11-17 : 3 bytes -> 13 bits */
unsigned zzo_hash(const uint8_t* p)
{
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    return (v * 0x00d68664u) >> (32 - HASH_BITS);
}

/* ------------------------------------------------------------------------------------------------
 * bit writer (outputbitstream.h).  LSB-first; the reference ORs into a 64-bit accumulator and stores
 * whole little-endian words, which is the same byte sequence as this byte-granular writer.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t* buf; size_t cap; size_t pos;      /* whole bytes written */
    uint64_t acc; int used;                    /* pending bits (< 8 after every append) */
    int overflow;
} bitw;

static void bw_init(bitw* w, uint8_t* buf, size_t cap) { w->buf = buf; w->cap = cap; w->pos = 0; w->acc = 0; w->used = 0; w->overflow = 0; }

static void bw_append(bitw* w, uint64_t bits, int count)            /* outputbitstream.h:83-98 */
{
    while (count > 0) {
        int take = count > 32 ? 32 : count;
        uint64_t part = take == 64 ? bits : (bits & ((1ull << take) - 1));
        w->acc |= part << w->used;
        w->used += take;
        while (w->used >= 8) {
            if (w->pos < w->cap) w->buf[w->pos] = (uint8_t)w->acc; else w->overflow = 1;
            w->pos++;
            w->acc >>= 8;
            w->used -= 8;
        }
        bits >>= take;
        count -= take;
    }
}

static void bw_code(bitw* w, zcode c) { bw_append(w, c.bits, c.length); }
static void bw_pad(bitw* w) { bw_append(w, 0, (-w->used) & 7); }                 /* outputbitstream.h:100-103 */
static void bw_flush(bitw* w) { bw_pad(w); }                                     /* outputbitstream.h:105-124 */
static void bw_u16(bitw* w, unsigned v) { bw_pad(w); bw_append(w, v & 0xFFFF, 16); }
static void bw_u32(bitw* w, uint32_t v) { bw_pad(w); bw_append(w, v, 32); }
static void bw_be32(bitw* w, uint32_t v)                                         /* outputbitstream.h:145-152 */
{
    bw_pad(w);
    bw_append(w, (v >> 24) & 0xFF, 8); bw_append(w, (v >> 16) & 0xFF, 8);
    bw_append(w, (v >> 8) & 0xFF, 8);  bw_append(w, v & 0xFF, 8);
}
static void bw_bytes(bitw* w, const uint8_t* src, size_t n)                      /* outputbitstream.h:155-160 */
{
    bw_flush(w);
    if (w->pos + n <= w->cap) memcpy(w->buf + w->pos, src, n); else w->overflow = 1;
    w->pos += n;
}

size_t zzo_bitstream_kat(const uint64_t* bits, const int* counts, int n, int flush, uint8_t* buf, size_t cap)
{
    /* The reference stores nothing before Flush for < 64 pending bits (TestBitOutput.cpp:16); the KAT
     * only looks at the bytes after Flush, so the byte-granular writer is compared after flushing. */
    bitw w; bw_init(&w, buf, cap);
    uint8_t* scratch = (uint8_t*)calloc(cap ? cap : 1, 1);
    bitw s; bw_init(&s, scratch, cap);
    for (int i = 0; i < n; ++i) bw_append(&s, bits[i], counts[i]);
    if (!flush) { free(scratch); (void)w; return 0; }
    bw_flush(&s);
    memcpy(buf, scratch, s.pos <= cap ? s.pos : cap);
    free(scratch);
    return s.pos;
}

/* ------------------------------------------------------------------------------------------------
 * Huffman code lengths: heap-based tree with libstdc++ heap layout (huffman.cpp:55-154).
 * The heap primitives restate GCC 13 bits/stl_heap.h (__push_heap, __adjust_heap, make_heap,
 * pop_heap) because tie-breaks between equal frequencies are decided by the heap layout alone.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int frequency; int id; } hrec;                      /* huffman.h:10-14 */
typedef struct { int frequency; int left; int right; int bits; } titem;   /* huffman.h:16-22 */

#define HCOMP(a, b) ((a).frequency > (b).frequency)                 /* huffman.cpp:55-63 */

static void heap_push_(hrec* h, int hole, int top, hrec v)
{
    int parent = (hole - 1) / 2;
    while (hole > top && HCOMP(h[parent], v)) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = v;
}

static void heap_adjust(hrec* h, int hole, int len, hrec v)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (HCOMP(h[child], h[child - 1])) child--;
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    heap_push_(h, hole, top, v);
}

static void heap_make(hrec* h, int len)
{
    if (len < 2) return;
    for (int parent = (len - 2) / 2; ; --parent) {
        heap_adjust(h, parent, len, h[parent]);
        if (parent == 0) return;
    }
}

static void heap_pop(hrec* h, int len)      /* std::pop_heap: moves the minimum to h[len-1] */
{
    if (len > 1) {
        hrec v = h[len - 1];
        h[len - 1] = h[0];
        heap_adjust(h, 0, len - 1, v);
    }
}

/* huffman.cpp:67-120 ; tree must hold 2*n items, recs n items.  Returns max leaf depth. */
static int calculate_tree(const int* freqs, int n, int minFreq, titem* tree, int* treeSize, hrec* recs)
{
    int tn = 0, rn = 0;
    for (int i = 0; i < n; ++i) {
        if (freqs[i] == 0) { titem t = { 0, i, -1, 0 }; tree[tn++] = t; continue; }
        int f = freqs[i] > minFreq ? freqs[i] : minFreq;
        titem t = { f, i, -1, 0 }; tree[tn++] = t;
        hrec r = { f, i }; recs[rn++] = r;
    }
    heap_make(recs, rn);
    while (rn >= 2) {
        heap_pop(recs, rn); hrec a = recs[--rn];
        heap_pop(recs, rn); hrec b = recs[--rn];
        titem t = { a.frequency + b.frequency, a.id, b.id, 0 };
        tree[tn++] = t;
        hrec r = { t.frequency, tn - 1 };
        recs[rn++] = r;
        heap_push_(recs, rn - 1, 0, r);
    }
    int maxLength = 0;
    for (int i = tn - 1; i != 0; --i) {                              /* index 0 is not visited */
        titem item = tree[i];
        if (item.right == -1) { if (item.bits > maxLength) maxLength = item.bits; continue; }
        tree[item.left].bits = item.bits + 1;
        tree[item.right].bits = item.bits + 1;
    }
    *treeSize = tn;
    return maxLength;
}

/* huffman.cpp:122-154 */
int zzo_calc_lengths_iters(const int* freqs, int n, int maxLength, int* lengths)
{
    titem* tree = (titem*)malloc(sizeof(titem) * 2 * (size_t)(n > 0 ? n : 1));
    hrec* recs = (hrec*)malloc(sizeof(hrec) * (size_t)(n > 0 ? n : 1));
    int minFreq = 0, iters = 0;
    for (;;) {
        int tn = 0;
        int max = calculate_tree(freqs, n, minFreq, tree, &tn, recs);
        iters++;
        if (max <= maxLength) {
            for (int i = 0; i < n; ++i) lengths[i] = 0;
            for (int i = 0; i < tn; ++i) {
                if (tree[i].right != -1) break;
                lengths[tree[i].left] = tree[i].frequency == 0 ? 0 : (tree[i].bits > 1 ? tree[i].bits : 1);
            }
            break;
        }
        int total = 0;
        for (int i = 0; i < n; ++i) total += freqs[i];
        int step = total / (1 << maxLength);
        minFreq += step > 1 ? step : 1;
    }
    free(tree); free(recs);
    return iters;
}

void zzo_calc_lengths(const int* freqs, int n, int maxLength, int* lengths)
{
    (void)zzo_calc_lengths_iters(freqs, n, maxLength, lengths);
}

/* huffman.cpp:158-216 : RLE of a code-length sequence into symbols 0..18 */
typedef struct { uint8_t value; uint8_t payLoad; } lrec;

static int add_records(lrec* v, int vn, int value, int count)       /* huffman.cpp:158-189 */
{
    if (count == 0) return vn;
    if (value == 0) {
        while (count >= 3) {
            int w = count < 138 ? count : 138;
            count -= w;
            v[vn].value = (uint8_t)(w < 11 ? 17 : 18); v[vn].payLoad = (uint8_t)w; vn++;
        }
    } else {
        v[vn].value = (uint8_t)value; v[vn].payLoad = 0; vn++;
        count--;
        while (count >= 3) {
            int w = count < 6 ? count : 6;
            count -= w;
            v[vn].value = 16; v[vn].payLoad = (uint8_t)w; vn++;
        }
    }
    for (int i = 0; i < count; ++i) { v[vn].value = (uint8_t)value; v[vn].payLoad = 0; vn++; }
    return vn;
}

static int from_lengths(const int* lengths, int n, int* freqs19, lrec* out)   /* huffman.cpp:191-216 */
{
    int vn = 0, current = -1, count = 0;
    for (int i = 0; i < n; ++i) {
        if (lengths[i] == current) { count++; continue; }
        vn = add_records(out, vn, current, count);
        current = lengths[i];
        count = 1;
    }
    vn = add_records(out, vn, current, count);
    for (int i = 0; i < vn; ++i) freqs19[out[i].value]++;
    return vn;
}

int zzo_from_lengths(const int* lengths, int n, int* freqs19, uint8_t* recordsOut, int maxRecords)
{
    lrec* r = (lrec*)malloc(sizeof(lrec) * (size_t)(n + 1));
    int c = from_lengths(lengths, n, freqs19, r);
    for (int i = 0; i < c && i < maxRecords; ++i) { recordsOut[2 * i] = r[i].value; recordsOut[2 * i + 1] = r[i].payLoad; }
    free(r);
    return c;
}

/* ------------------------------------------------------------------------------------------------
 * checksums
 * ---------------------------------------------------------------------------------------------- */
#define MOD_ADLER 65521u

uint32_t zzo_adler32(uint32_t start, const uint8_t* d, size_t n)
{
    uint32_t a = start & 0xFFFF, b = start >> 16;
    while (n > 0) {
        size_t k = n < 5552 ? n : 5552;
        for (size_t i = 0; i < k; ++i) { a += d[i]; b += a; }
        a %= MOD_ADLER; b %= MOD_ADLER;
        d += k; n -= k;
    }
    return (b << 16) | a;
}

uint32_t zzo_adler32x_literal(uint32_t start, const uint8_t* data, size_t len)  /* adler.cpp:17-43 */
{
    uint64_t a = (start & 0xFFFF), b = start >> 16;
    size_t index = 0;
    for (; index + 4 <= len; index += 4) {
        a += data[index];     b += a;
        a += data[index + 1]; b += a;
        a += data[index + 2]; b += a;
        a += data[index + 3]; b += a;
    }
    for (; index < len; ++index) {
        a += data[index]; b += a;
        a %= MOD_ADLER; b %= MOD_ADLER;
    }
    a %= MOD_ADLER; b %= MOD_ADLER;
    return (uint32_t)((b << 16) | a);
}

uint32_t zzo_combine(uint32_t first, uint32_t second, size_t lenSecond)     /* adler.cpp:5-15 */
{
    uint64_t a = (first & 0xFFFF) + (second & 0xFFFF);
    uint64_t b = (first >> 16) + (second >> 16);
    b += (uint64_t)lenSecond * (first & 0xFFFF);
    a %= MOD_ADLER; b %= MOD_ADLER;
    return (uint32_t)((b << 16) | a);
}

uint32_t zzo_crc32(const uint8_t* buffer, size_t length, uint32_t startValue)   /* crc.cpp:24-33 */
{
    ensure_tables();
    uint32_t crc = ~startValue;
    for (size_t i = 0; i < length; ++i)
        crc = (crc >> 8) ^ g_crcTable[(crc & 0xFF) ^ buffer[i]];
    return ~crc;
}

/* CRC-32 of A||B from crc(A), crc(B), len(B): multiply crc(A) by x^(8*len2) modulo the reflected
 * polynomial, by square-and-multiply in GF(2)[x].  Not in the reference (it can only chain through
 * startValue, crc.cpp:24-26). */
static uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
    /* reflected representation: bit 31 is x^0 */
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1) ? 0xEDB88320u : 0);
    }
    return p;
}

uint32_t zzo_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
    if (len2 == 0) return crc1;
    uint32_t xp = 0x80000000u;          /* x^0 */
    uint32_t sq = 0x00800000u;          /* x^8 : one byte */
    uint64_t k = len2;
    while (k) {
        if (k & 1) xp = gf2_mulmod(xp, sq);
        sq = gf2_mulmod(sq, sq);
        k >>= 1;
    }
    return gf2_mulmod(crc1, xp) ^ crc2;
}

/* ------------------------------------------------------------------------------------------------
 * block encoder (encoder.h:37-116, encoder.cpp)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t literals; uint16_t backoffset; uint16_t length; } crec;   /* encoder.h:29-34 */

typedef struct {
    int level;
    int table[HASH_SIZE];
    bitw bw;
    crec* recs; int nrec;
    zcode codes[286], lcodes[259], dcodes[30];
    int defects;
    /* last dynamic-block statistics */
    int block_type; int64_t block_bits;
    int lit_freq[286], dist_freq[30], lit_len[286], dist_len[30], meta_len[19];
} enc;

static void enc_init(enc* e, int level, uint8_t* out, size_t cap)   /* encoder.cpp:527-537 */
{
    memset(e, 0, sizeof *e);
    e->level = level;
    for (int i = 0; i < HASH_SIZE; ++i) e->table[i] = EMPTY_SLOT;
    bw_init(&e->bw, out, cap);
    e->recs = (crec*)malloc(sizeof(crec) * MAX_RECORDS);
    e->block_type = -1;
}

static void enc_free(enc* e) { free(e->recs); e->recs = NULL; }

static void start_block(enc* e, int type, int final)                /* encoder.cpp:143-147 */
{
    bw_append(&e->bw, (uint64_t)(final ? 1 : 0), 1);
    bw_append(&e->bw, (uint64_t)type, 2);
}

static void write_distance(enc* e, const zcode* dist, int offset)    /* encoder.cpp:135-141 */
{
    int bucket = g_distLut[offset];
    bw_code(&e->bw, dist[bucket]);
    bw_append(&e->bw, (uint64_t)(offset - g_distBase[bucket]), g_extraDist[bucket]);
}

static void add_hash_entries(enc* e, const uint8_t* src, int i, int extra)     /* encoder.cpp:474-480 */
{
    for (int n = i; n < i + extra; ++n) e->table[zzo_hash(src + n)] = n;
}

static void fix_hash_table(enc* e, int offset)                       /* encoder.cpp:320-327 */
{
    for (int i = 0; i < HASH_SIZE; ++i) e->table[i] -= offset;
}

/* encoder.cpp:482-502 (output-space checks dropped: the caller sizes the buffer with zzo_bound) */
static int write_uncompressed_block(enc* e, const uint8_t* src, int byteCount, int final)
{
    int length = byteCount < 0xFFFF ? byteCount : 0xFFFF;
    start_block(e, ZZO_BLOCK_STORED, final && length == byteCount);
    bw_pad(&e->bw);
    bw_u16(&e->bw, (unsigned)length);
    bw_u16(&e->bw, (unsigned)(~length) & 0xFFFF);
    bw_bytes(&e->bw, src, (size_t)length);
    return length;
}

static int uncompressed_fallback(enc* e, int length, const uint8_t* src, int final)    /* encoder.cpp:305-317 */
{
    int written = 0;
    while (written < length) written += write_uncompressed_block(e, src + written, length - written, final);
    return written;
}

/* encoder.cpp:329-373.  lo = lowest readable position relative to src (<= 0). */
static int write_block_fixed_huff(enc* e, const uint8_t* src, int byteCount, int final)
{
    const int n = byteCount;
    start_block(e, ZZO_BLOCK_FIXED, final);
    for (int i = 0; i < n; ++i) {
        unsigned h = zzo_hash(src + i + 1);
        int distance = i - e->table[h];
        e->table[h] = i;
        if ((unsigned)distance <= MAX_DISTANCE) {
            const uint8_t* a = src + i; const uint8_t* b = a - distance;
            int m = 0;
            while (m < 8 && a[m] == b[m]) ++m;                        /* 8-byte XOR + ZeroCount */
            if (m == 8) {                                           /* remain(a,b,8,n-i) */
                int maxLength = n - i < MAX_LENGTH ? n - i : MAX_LENGTH;
                while (m < maxLength && a[m] == b[m]) ++m;
                if (m > maxLength) m = maxLength;
            } else if (m > n - i) {                                 /* R2: reference does not clamp here */
                if (m > 3) e->defects |= ZZO_DEFECT_R2;
                m = n - i;
            }
            if (m > 3) {
                bw_code(&e->bw, g_lcodesF[m]);
                write_distance(e, g_dcodesF, distance);
                i += m - 1;
                continue;
            }
        }
        bw_code(&e->bw, g_codesF[src[i]]);
    }
    fix_hash_table(e, n);
    bw_code(&e->bw, g_codesF[256]);
    e->block_type = ZZO_BLOCK_FIXED;
    return n;
}

/* encoder.cpp:375-440.  lo = lowest readable position relative to src (global offset 0). */
static int first_pass(enc* e, const uint8_t* src, int startPos, int end, long lo)
{
    if (startPos == end) return startPos;
    int bre = startPos + 1;
    const int firstRec = e->nrec;
    int j = startPos + 1;
    while (j < end) {
        const uint8_t* s = src + j;
        unsigned h = zzo_hash(s);
        long p = e->table[h];
        long distance = (long)j - p;
        e->table[h] = j;
        if (distance >= MAX_DISTANCE) { j++; continue; }
        int fwd = 0;
        while (fwd < MAX_LENGTH && s[fwd] == src[p + fwd]) ++fwd;   /* XOR/ZeroCount + remain(s,s-d,8) */
        int maxBack = j - bre, lb = 0;                              /* countMatchBackward, encoder.cpp:92-102 */
        while (lb < maxBack && lb < 259) {
            if (p - 1 - lb < lo) { e->defects |= ZZO_DEFECT_R4; break; }
            if (src[j - 1 - lb] != src[p - 1 - lb]) break;
            ++lb;
        }
        if (lb >= 259) { e->defects |= ZZO_DEFECT_R6; lb = 258; }
        int m = fwd + lb;
        if (m < 4) { j++; continue; }
        if (m > MAX_LENGTH) m = MAX_LENGTH;
        int ms = j - lb;
        add_hash_entries(e, src, ms + 1, m);
        crec r = { (uint32_t)(ms - bre), (uint16_t)distance, (uint16_t)m };
        e->recs[e->nrec++] = r;
        bre = ms + m;
        j = bre + 1;
        if (e->nrec == MAX_RECORDS) { end = 0; break; }
    }
    int lost = 0;
    if (e->nrec > firstRec) e->recs[firstRec].literals += 1;
    else { e->defects |= ZZO_DEFECT_R1; lost = 1; }                 /* R1: keep the byte */
    if (bre > end) return bre;
    crec r = { (uint32_t)(end - bre + lost), 0, 0 };
    e->recs[e->nrec++] = r;
    return end;
}

static void get_frequencies(enc* e, const uint8_t* src, int* sym, int* dist)     /* encoder.cpp:442-471 */
{
    int index = 0;
    for (int n = 0; n < e->nrec; ++n) {
        crec r = e->recs[n];
        for (unsigned i = 0; i < r.literals; ++i) sym[src[index + (int)i]]++;
        index += (int)r.literals;
        if (r.length == 0) continue;
        sym[g_lenCode[r.length]]++;
        dist[g_distLut[r.backoffset]]++;
        index += r.length;
    }
    sym[256]++;
}

static int64_t count_bits(const int* freqs, int n, const int* lengths, const uint8_t* extra)   /* encoder.cpp:178-187 */
{
    int64_t total = 0;
    for (int i = 0; i < n; ++i) total += (int64_t)freqs[i] * (lengths[i] + extra[i]);
    return total;
}

static void write_lengths(bitw* w, int64_t* counter, const lrec* recs, int n, const zcode* table)   /* encoder.cpp:20-46 */
{
    for (int i = 0; i < n; ++i) {
        zcode c = table[recs[i].value];
        int eb = 0; unsigned ev = 0;
        switch (recs[i].value) {
        case 16: eb = 2; ev = (unsigned)recs[i].payLoad - 3; break;
        case 17: eb = 3; ev = (unsigned)recs[i].payLoad - 3; break;
        case 18: eb = 7; ev = (unsigned)recs[i].payLoad - 11; break;
        default: break;
        }
        if (w) { bw_code(w, c); if (eb) bw_append(w, ev, eb); }
        else *counter += c.length + eb;
    }
}

static void write_records(enc* e, const uint8_t* src)               /* encoder.cpp:149-169 */
{
    int offset = 0;
    for (int i = 0; i < e->nrec; ++i) {
        crec r = e->recs[i];
        for (unsigned n = 0; n < r.literals; ++n) bw_code(&e->bw, e->codes[src[offset + (int)n]]);
        if (r.length != 0) {
            bw_code(&e->bw, e->lcodes[r.length]);
            write_distance(e, e->dcodes, r.backoffset);
        }
        offset += (int)r.literals + r.length;
    }
}

/* encoder.cpp:217-303 */
static int write_block_2pass(enc* e, const uint8_t* src, int byteCount, int final, long lo)
{
    e->nrec = 0;
    int target = byteCount - MAX_LENGTH > 0 ? byteCount - MAX_LENGTH : 0;
    int length = 0;
    while (target > 0 && e->nrec < MAX_RECORDS) {
        int batch = target < BATCH ? target : BATCH;
        int newEnd = first_pass(e, src, length, length + batch, lo);
        target -= newEnd - length;
        length = newEnd;
    }
    if (target <= 0 && e->nrec < MAX_RECORDS) {
        crec r = { (uint32_t)(byteCount - length), 0, 0 };
        e->recs[e->nrec++] = r;
        length = byteCount;
    }
    fix_hash_table(e, length);

    int symF[286] = { 0 }, distF[30] = { 0 }, metaF[19] = { 0 };
    int symL[286], distL[30], metaL[19];
    lrec symRecs[287], distRecs[31];
    get_frequencies(e, src, symF, distF);

    zzo_calc_lengths(symF, 286, 15, symL);                          /* ComputeCodes, encoder.cpp:171-176 */
    generate(symL, 286, e->codes);
    int nSymRecs = from_lengths(symL, 286, metaF, symRecs);
    int64_t bits = count_bits(symF, 286, symL, g_extraLen);

    zzo_calc_lengths(distF, 30, 15, distL);
    generate(distL, 30, e->dcodes);
    int nDistRecs = from_lengths(distL, 30, metaF, distRecs);
    bits += count_bits(distF, 30, distL, g_extraDist);

    zcode meta[19];
    memset(meta, 0, sizeof meta);
    zzo_calc_lengths(metaF, 19, 7, metaL);
    generate(metaL, 19, meta);

    int64_t total = 3 + 5 + 5 + 4 + 3 * 19 + bits;
    write_lengths(NULL, &total, symRecs, nSymRecs, meta);
    write_lengths(NULL, &total, distRecs, nDistRecs, meta);

    memcpy(e->lit_freq, symF, sizeof symF);  memcpy(e->dist_freq, distF, sizeof distF);
    memcpy(e->lit_len, symL, sizeof symL);   memcpy(e->dist_len, distL, sizeof distL);
    memcpy(e->meta_len, metaL, sizeof metaL);
    e->block_bits = total;

    int64_t required = (total + 8) / 8;
    if (required >= length) {
        e->block_type = ZZO_BLOCK_STORED;
        return uncompressed_fallback(e, length, src, final);
    }
    e->block_type = ZZO_BLOCK_DYNAMIC;
    start_block(e, ZZO_BLOCK_DYNAMIC, length < byteCount ? 0 : final);
    bw_append(&e->bw, 286 - 257, 5);
    bw_append(&e->bw, 30 - 1, 5);
    bw_append(&e->bw, 19 - 4, 4);
    for (int i = 0; i < 19; ++i) bw_append(&e->bw, (uint64_t)metaL[g_order[i]], 3);
    write_lengths(&e->bw, NULL, symRecs, nSymRecs, meta);
    write_lengths(&e->bw, NULL, distRecs, nDistRecs, meta);
    create_merged_length_codes(e->lcodes, e->codes);
    write_records(e, src);
    bw_code(&e->bw, e->codes[256]);
    return length;
}

static int write_deflate_block(enc* e, const uint8_t* src, int inputLength, int final, long lo)   /* encoder.cpp:506-525 */
{
    if (e->level == 0) return write_uncompressed_block(e, src, inputLength, final);
    if (e->level == 1) return write_block_fixed_huff(e, src, inputLength, final);
    if (inputLength > BLOCK_CAP) { inputLength = BLOCK_CAP; final = 0; }
    return write_block_2pass(e, src, inputLength, final, lo);
}

/* encoder.cpp:539-551.  off = global offset of start (for the R4 clamp). */
static void add_data(enc* e, const uint8_t* start, const uint8_t* end, int final, uint64_t off)
{
    while (start != end) {
        int n = write_deflate_block(e, start, (int)(end - start), final, -(long)off);
        if (n <= 0) return;
        start += n; off += (uint64_t)n;
    }
}

static void fill_info(const enc* e, zzo_chunk_info* info)
{
    if (!info) return;
    info->defects = e->defects;
    info->block_type = e->block_type;
    info->block_bits = e->block_bits;
    info->n_records = e->nrec;
    memcpy(info->lit_freq, e->lit_freq, sizeof e->lit_freq);   memcpy(info->dist_freq, e->dist_freq, sizeof e->dist_freq);
    memcpy(info->lit_len, e->lit_len, sizeof e->lit_len);      memcpy(info->dist_len, e->dist_len, sizeof e->dist_len);
    memcpy(info->meta_len, e->meta_len, sizeof e->meta_len);
    int nm = 0, pos = 0;
    for (int i = 0; i < e->nrec; ++i) {
        crec r = e->recs[i];
        if (info->records && i < info->max_records) {
            info->records[3 * i] = r.literals; info->records[3 * i + 1] = r.backoffset; info->records[3 * i + 2] = r.length;
        }
        pos += (int)r.literals;
        if (r.length) {
            if (info->matches && nm < info->max_matches) {
                info->matches[3 * nm] = (uint32_t)pos; info->matches[3 * nm + 1] = r.length; info->matches[3 * nm + 2] = r.backoffset;
            }
            nm++;
            pos += r.length;
        }
    }
    info->n_matches = nm;
}

size_t zzo_chunk_encode(const uint8_t* chunk, size_t n, size_t dict, uint64_t global_off,
                        int level, int final, uint8_t* out, size_t cap, zzo_chunk_info* info)
{
    ensure_tables();
    enc* e = (enc*)malloc(sizeof(enc));
    enc_init(e, level, out, cap);
    if (dict > 0) {
        if (level >= 2) add_hash_entries(e, chunk, -(int)dict, (int)dict);            /* A.7 */
        else if (level == 1) for (int i = -(int)dict; i < 0; ++i) e->table[zzo_hash(chunk + i + 1)] = i;
    }
    if (final) {
        add_data(e, chunk, chunk + n, 1, global_off);
    } else {
        add_data(e, chunk, chunk + n - 1, 0, global_off);                             /* zzflate.cpp:116 */
        e->level = 0;                                                                   /* zzflate.cpp:119 */
        add_data(e, chunk + n - 1, chunk + n, 0, global_off + n - 1);                 /* zzflate.cpp:120 */
    }
    bw_flush(&e->bw);
    fill_info(e, info);
    size_t written = e->bw.overflow ? ~(size_t)0 : e->bw.pos;
    enc_free(e); free(e);
    return written;
}

void zzo_chunk_candidates(const uint8_t* chunk, size_t n, size_t dict, uint16_t* cand)
{
    int* table = (int*)malloc(sizeof(int) * HASH_SIZE);
    for (int i = 0; i < HASH_SIZE; ++i) table[i] = -(1 << 30);
    for (long k = -(long)dict; k < 0; ++k) table[zzo_hash(chunk + k)] = (int)k;
    if (n > 0) cand[0] = 0;
    for (long j = 1; j < (long)n; ++j) {
        unsigned h = zzo_hash(chunk + j);
        long d = j - table[h];
        table[h] = (int)j;
        cand[j] = (uint16_t)(d < MAX_DISTANCE ? d : 0);
    }
    free(table);
}

/* ------------------------------------------------------------------------------------------------
 * stream framing (zzflate.cpp:10-63,170-192)
 * ---------------------------------------------------------------------------------------------- */
static size_t header_bytes(int format, uint8_t* h)
{
    static const uint8_t gz[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xFF };          /* zzflate.cpp:28 */
    switch (format) {
    case ZZO_ZLIB: {                                                                    /* zzflate.cpp:30-36 */
        unsigned cmf = 8 | (7 << 4), flg = 0;
        unsigned rem = (cmf * 0x100 + flg) % 31;
        flg |= (31 - rem) & 0xF;
        h[0] = (uint8_t)cmf; h[1] = (uint8_t)flg;
        return 2;
    }
    case ZZO_GZIP: memcpy(h, gz, 10); return 10;
    default: return 0;
    }
}

static size_t trailer_bytes(int format, uint32_t adler, uint32_t crc, uint64_t n, uint8_t* t)   /* zzflate.cpp:170-192 */
{
    bitw w; bw_init(&w, t, 8);
    if (format == ZZO_ZLIB) { bw_be32(&w, adler); }
    else if (format == ZZO_GZIP) { bw_u32(&w, crc); bw_u32(&w, (uint32_t)n); }
    bw_flush(&w);
    return w.pos;
}

size_t zzo_bound(size_t n, int level, size_t chunk)
{
    size_t chunks = n ? (n + chunk - 1) / chunk : 1;
    size_t per = level == 1 ? (chunk * 9 + 7) / 8 + 16 : chunk + 16;    /* A.6: 65 546 for 65 536 */
    return 10 + chunks * per + 8;
}

static const uint8_t g_emptyFinalStored[5] = { 0x01, 0x00, 0x00, 0xFF, 0xFF };         /* R7 policy */

size_t zzo_stream_chunked(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                          int format, int level, size_t chunk, size_t dict, int* defects)
{
    ensure_tables();
    if (level < 0 || level > 3) return ~(size_t)0;
    uint8_t h[10], t[8];
    size_t hl = header_bytes(format, h);
    if (cap < hl) return ~(size_t)0;
    memcpy(dest, h, hl);
    size_t pos = hl;
    int def = 0;
    if (n == 0) {
        if (pos + 5 > cap) return ~(size_t)0;
        memcpy(dest + pos, g_emptyFinalStored, 5); pos += 5;
    }
    for (size_t off = 0; off < n; off += chunk) {
        size_t len = n - off < chunk ? n - off : chunk;
        size_t d = off < dict ? off : dict;
        zzo_chunk_info info; memset(&info, 0, sizeof info);
        size_t w = zzo_chunk_encode(src + off, len, d, off, level, off + len == n, dest + pos, cap - pos, &info);
        if (w == ~(size_t)0) return w;
        def |= info.defects;
        pos += w;
    }
    size_t tl = trailer_bytes(format, format == ZZO_ZLIB ? zzo_adler32(1, src, n) : 0,
                              format == ZZO_GZIP ? zzo_crc32(src, n, 0) : 0, n, t);
    if (pos + tl > cap) return ~(size_t)0;
    memcpy(dest + pos, t, tl);
    if (defects) *defects = def;
    return pos + tl;
}

size_t zzo_stream_reference(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                            int format, int level, int* defects)
{
    ensure_tables();
    if (level < 0 || level > 3) return ~(size_t)0;
    uint8_t h[10], t[8];
    size_t hl = header_bytes(format, h);
    if (cap < hl) return ~(size_t)0;
    memcpy(dest, h, hl);
    enc* e = (enc*)malloc(sizeof(enc));
    enc_init(e, level, dest + hl, cap - hl);
    add_data(e, src, src + n, 1, 0);                                /* zzflate.cpp:86-88 */
    bw_flush(&e->bw);
    size_t pos = hl + e->bw.pos;
    int overflow = e->bw.overflow;
    if (defects) *defects = e->defects;
    enc_free(e); free(e);
    if (overflow) return ~(size_t)0;
    /* the reference trailer uses adler32x (R3 above ~362 MiB); below that it equals the true Adler-32 */
    size_t tl = trailer_bytes(format, format == ZZO_ZLIB ? zzo_adler32(1, src, n) : 0,
                              format == ZZO_GZIP ? zzo_crc32(src, n, 0) : 0, n, t);
    if (pos + tl > cap) return ~(size_t)0;
    memcpy(dest + pos, t, tl);
    return pos + tl;
}

/* ------------------------------------------------------------------------------------------------
 * multi-threaded chunked stream: CPU baseline ("port") for bench.py
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* src; size_t n; int level; size_t chunk, dict;
    size_t nchunks; uint8_t* slots; size_t slotSize; size_t* sizes;
    size_t next; pthread_mutex_t mu;
} mtjob;

static void* mt_worker(void* arg)
{
    mtjob* j = (mtjob*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t c = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (c >= j->nchunks) break;
        size_t off = c * j->chunk;
        size_t len = j->n - off < j->chunk ? j->n - off : j->chunk;
        size_t d = off < j->dict ? off : j->dict;
        j->sizes[c] = zzo_chunk_encode(j->src + off, len, d, off, j->level, off + len == j->n,
                                       j->slots + c * j->slotSize, j->slotSize, NULL);
    }
    return NULL;
}

size_t zzo_stream_chunked_mt(uint8_t* dest, size_t cap, const uint8_t* src, size_t n,
                             int format, int level, size_t chunk, size_t dict, int threads)
{
    ensure_tables();
    if (n == 0 || threads <= 1) return zzo_stream_chunked(dest, cap, src, n, format, level, chunk, dict, NULL);
    mtjob j; memset(&j, 0, sizeof j);
    j.src = src; j.n = n; j.level = level; j.chunk = chunk; j.dict = dict;
    j.nchunks = (n + chunk - 1) / chunk;
    j.slotSize = (level == 1 ? (chunk * 9 + 7) / 8 : chunk) + 32;
    j.slots = (uint8_t*)malloc(j.nchunks * j.slotSize);
    j.sizes = (size_t*)calloc(j.nchunks, sizeof(size_t));
    pthread_mutex_init(&j.mu, NULL);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, mt_worker, &j);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    uint8_t h[10], t[8];
    size_t pos = header_bytes(format, h);
    size_t result = ~(size_t)0;
    if (pos <= cap) {
        memcpy(dest, h, pos);
        int ok = 1;
        for (size_t c = 0; c < j.nchunks && ok; ++c) {
            if (j.sizes[c] == ~(size_t)0 || pos + j.sizes[c] > cap) { ok = 0; break; }
            memcpy(dest + pos, j.slots + c * j.slotSize, j.sizes[c]);
            pos += j.sizes[c];
        }
        if (ok) {
            size_t tl = trailer_bytes(format, format == ZZO_ZLIB ? zzo_adler32(1, src, n) : 0,
                                      format == ZZO_GZIP ? zzo_crc32(src, n, 0) : 0, n, t);
            if (pos + tl <= cap) { memcpy(dest + pos, t, tl); result = pos + tl; }
        }
    }
    free(th); free(j.slots); free(j.sizes); pthread_mutex_destroy(&j.mu);
    return result;
}
