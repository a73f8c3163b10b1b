/* CPU model of the level-1 kernel's step logic (development aid, no GPU; see k_fixed in zz_kernels.cu).
 *
 * seq_l1  : the level-1 walk as the oracle states it (oracle/zz_oracle.c write_block_fixed_huff, which follows
 *           encoder.cpp:329-373): only visited positions enter the table, the hash is taken one byte ahead.
 * warp_l1 : the same walk in steps of 32 positions.  All 32 positions are presumed visited; every lane knows its
 *           match length against the table's entry (mOld) and against the nearest lower lane of its hash group
 *           (mLow).  The matches of a step are then resolved one after the other on these two numbers only.  A lane
 *           whose true candidate is neither (a lower lane of its group, but not the nearest one) ends the step.
 *
 *   gcc -O2 -shared -fPIC -o tools/model/libl1model.so tools/model/l1_model.c
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define HASH_BITS 13
#define HASH_SIZE (1 << HASH_BITS)
#define MAX_DISTANCE 32768
#define MAX_LENGTH 258

static unsigned hash3(const uint8_t* p) { unsigned v = p[0] | (p[1] << 8) | (p[2] << 16); return (v * 0x00d68664u) >> (32 - HASH_BITS); }

static long g_steps, g_amb, g_long;
void l1m_stats(long* o) { o[0] = g_steps; o[1] = g_amb; o[2] = g_long; g_steps = g_amb = g_long = 0; }

static void prime(int* table, const uint8_t* c, int dict)
{
    for (int i = 0; i < HASH_SIZE; ++i) table[i] = -(1 << 30);
    for (int i = -dict; i < 0; ++i) table[hash3(c + i + 1)] = i;
}

static int match8(const uint8_t* a, const uint8_t* b) { int m = 0; while (m < 8 && a[m] == b[m]) ++m; return m; }

/* tokens: (pos, len, dist) triples */
int l1m_seq(const uint8_t* c, int n, int dict, uint32_t* tok, int maxTok)
{
    int* table = malloc(sizeof(int) * HASH_SIZE); prime(table, c, dict);
    int k = 0;
    for (int i = 0; i < n; ++i) {
        unsigned h = hash3(c + i + 1);
        int d = i - table[h]; table[h] = i;
        if ((unsigned)d <= MAX_DISTANCE) {
            int m = match8(c + i, c + i - d);
            if (m == 8) { int mx = n - i < MAX_LENGTH ? n - i : MAX_LENGTH; while (m < mx && c[i + m] == c[i - d + m]) ++m; if (m > mx) m = mx; }
            else if (m > n - i) m = n - i;
            if (m > 3) { if (k < maxTok) { tok[3 * k] = i; tok[3 * k + 1] = m; tok[3 * k + 2] = d; } ++k; i += m - 1; }
        }
    }
    free(table);
    return k;
}

int l1m_warp(const uint8_t* c, int n, int dict, uint32_t* tok, int maxTok)
{
    int* table = malloc(sizeof(int) * HASH_SIZE); prime(table, c, dict);
    int k = 0, i0 = 0;
    while (i0 < n) {
        ++g_steps;
        unsigned h[32]; int valid[32], old[32], lowN[32], mOld[32], mLow[32];
        uint32_t grp[32];
        for (int j = 0; j < 32; ++j) { valid[j] = i0 + j < n; h[j] = valid[j] ? hash3(c + i0 + j + 1) : 0x10000u + j; }
        for (int j = 0; j < 32; ++j) {
            grp[j] = 0; for (int t = 0; t < 32; ++t) if (h[t] == h[j]) grp[j] |= 1u << t;
            const uint32_t lower = grp[j] & ((1u << j) - 1u);
            lowN[j] = lower ? 31 - __builtin_clz(lower) : -1;
            const int i = i0 + j, rem = n - i;
            mOld[j] = mLow[j] = 0; old[j] = 0;
            if (!valid[j]) continue;
            old[j] = table[h[j]];
            const int d = i - old[j];
            if ((unsigned)d <= MAX_DISTANCE) { int m = match8(c + i, c + old[j]); mOld[j] = m > rem ? rem : m; }
            if (lowN[j] >= 0) { int m = match8(c + i, c + i0 + lowN[j]); mLow[j] = m > rem ? rem : m; }
        }
        uint32_t V = 0; int p = 0, next = i0 + 32;
        for (;;) {
            uint32_t acc = 0, amb = 0; int m[32], d[32];
            const uint32_t fromP = p >= 32 ? 0u : ~((1u << p) - 1u);
            for (int j = p; j < 32; ++j) {
                if (!valid[j]) continue;
                const uint32_t elig = grp[j] & ((1u << j) - 1u) & (V | fromP);
                if (!elig) { m[j] = mOld[j]; d[j] = i0 + j - old[j]; }
                else if (31 - __builtin_clz(elig) == lowN[j]) { m[j] = mLow[j]; d[j] = j - lowN[j]; }
                else { amb |= 1u << j; m[j] = 0; d[j] = 0; }
                if (m[j] > 3) acc |= 1u << j;
            }
            const uint32_t stop = acc | amb;
            if (!stop) { V |= fromP; break; }
            const int f = __builtin_ctz(stop);
            V |= fromP & ((1u << f) - 1u);
            if ((amb >> f) & 1u) { ++g_amb; next = i0 + f; break; }
            V |= 1u << f;
            int L = m[f];
            if (L == 8) {
                ++g_long;
                const int fi = i0 + f, mx = n - fi < MAX_LENGTH ? n - fi : MAX_LENGTH;
                while (L < mx && c[fi + L] == c[fi - d[f] + L]) ++L;
                if (L > mx) L = mx;
            }
            if (k < maxTok) { tok[3 * k] = i0 + f; tok[3 * k + 1] = L; tok[3 * k + 2] = d[f]; } ++k;
            p = f + L;
            if (m[f] == 8 || p >= 32) { next = i0 + p; break; }
        }
        for (int j = 0; j < 32; ++j)
            if (valid[j] && ((V >> j) & 1u) && (grp[j] & ~((2u << j) - 1u) & V) == 0) table[h[j]] = i0 + j;
        i0 = next;
    }
    free(table);
    return k;
}
