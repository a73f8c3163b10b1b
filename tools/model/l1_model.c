/* CPU model of the level-1 kernel's step logic (development aid, no GPU; see k_fixed in zz_kernels.cu).
 *
 * seq_l1  : the level-1 walk as the oracle states it (oracle/zz_oracle.c write_block_fixed_huff, which follows
 *           encoder.cpp:329-373): only visited positions enter the table, the hash is taken one byte ahead.
 * warp_l1 : the same walk in steps of 32 positions.  All 32 positions are presumed visited; every lane knows its
 *           match length against the table's entry (mOld) and against the nearest lower lane of its hash group
 *           (mLow).  The matches of a step are then resolved one after the other on these two numbers only.  A lane
 *           whose true candidate is neither (a lower lane of its group, but not the nearest one) ends the step.
 * warp2_l1: the step as k_fixed settles it: the part of the window whose outcomes do not depend on the visited set in
 *           parallel (pointer doubling over the match starts), the rest one match per turn.
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

/* l1m_warp2: the step as the kernel settles it now.  Lanes without a lower lane of their hash group ("fixed" lanes) have
 * one possible candidate, so the part of the window in front of the first lane that has to be looked at (q) is settled
 * in parallel: T[j] = first fixed match start at or behind the end of j's match, the real match starts are the orbit of T
 * from the first match start, found by pointer doubling (three rounds cover the eight matches a window can hold);
 * the rest of the window is settled sequentially as in l1m_warp. */
static long g_par_rounds, g_par_steps, g_seq_iters;
void l1m_stats2(long* o) { o[0] = g_par_steps; o[1] = g_par_rounds; o[2] = g_seq_iters; g_par_rounds = g_par_steps = g_seq_iters = 0; }

int l1m_warp2(const uint8_t* c, int n, int dict, uint32_t* tok, int maxTok)
{
    int* table = malloc(sizeof(int) * HASH_SIZE); prime(table, c, dict);
    int k = 0, i0 = 0;
    while (i0 < n) {
        ++g_steps;
        unsigned h[32]; int valid[32], old[32], lowN[32], mOld[32], mLow[32], dOld[32];
        uint32_t grp[32], lower[32];
        for (int j = 0; j < 32; ++j) { valid[j] = i0 + j < n; h[j] = valid[j] ? hash3(c + i0 + j + 1) : 0x10000u + j; }
        uint32_t fixedAcc = 0, depends = 0;
        for (int j = 0; j < 32; ++j) {
            grp[j] = 0; for (int t = 0; t < 32; ++t) if (h[t] == h[j]) grp[j] |= 1u << t;
            lower[j] = grp[j] & ((1u << j) - 1u);
            lowN[j] = lower[j] ? 31 - __builtin_clz(lower[j]) : -1;
            const int i = i0 + j, rem = n - i;
            mOld[j] = mLow[j] = 0; old[j] = 0; dOld[j] = 0xFFFF;
            if (!valid[j]) continue;
            old[j] = table[h[j]];
            const long d = (long)i - old[j];
            if (d >= 0 && d <= MAX_DISTANCE) { dOld[j] = (int)d; int m = match8(c + i, c + old[j]); mOld[j] = m > rem ? rem : m; }
            if (lowN[j] >= 0) { int m = match8(c + i, c + i0 + lowN[j]); mLow[j] = m > rem ? rem : m; }
            if (!lower[j] && mOld[j] > 3) fixedAcc |= 1u << j;
            if (lower[j] && (mOld[j] > 3 || mLow[j] > 3 || (lower[j] & (lower[j] - 1)))) depends |= 1u << j;
        }
        uint32_t V = 0, starts = 0; int p = 0, adv = 32, done = 0;
        int myLen[32] = { 0 }, myDist[32] = { 0 };
        /* ---- parallel part: lanes below q ---- */
        const int q = depends ? __builtin_ctz(depends) : 32;
        const uint32_t belowQ = q >= 32 ? 0xffffffffu : ((1u << q) - 1u);
        const uint32_t accP = fixedAcc & belowQ;
        if (accP) {
            ++g_par_steps;
            int G[32];
            for (int j = 0; j < 32; ++j) {
                const int x = j + mOld[j];
                const uint32_t m2 = x >= 32 ? 0u : (accP & (0xffffffffu << x));
                G[j] = (mOld[j] == 8 || !m2) ? 32 : __builtin_ctz(m2);       /* a long match ends the step */
            }
            uint32_t R = 1u << __builtin_ctz(accP);
            for (;;) {
                uint32_t add = 0;
                for (int j = 0; j < 32; ++j) if (((R >> j) & 1u) && G[j] < 32) add |= 1u << G[j];   /* REDUX.OR */
                if ((add & ~R) == 0) break;
                R |= add; ++g_par_rounds;
                int G2[32];
                for (int j = 0; j < 32; ++j) G2[j] = G[j] >= 32 ? 32 : G[G[j]];                         /* SHFL */
                memcpy(G, G2, sizeof G);
            }
            uint32_t covered = 0;
            for (int j = 0; j < 32; ++j) if ((R >> j) & 1u) { covered |= (((1u << mOld[j]) - 2u) << j); myLen[j] = mOld[j]; myDist[j] = dOld[j]; }
            const int last = 31 - __builtin_clz(R);
            int L = mOld[last];
            starts = R;
            if (L == 8) {
                ++g_long;
                const int fi = i0 + last, mx = n - fi < MAX_LENGTH ? n - fi : MAX_LENGTH;
                while (L < mx && c[fi + L] == c[fi - dOld[last] + L]) ++L;
                if (L > mx) L = mx;
                myLen[last] = L;
                adv = last + L; done = 1;
                V = ~covered & ((2u << last) - 1u);
            } else {
                const int x = last + L;
                if (x >= 32) { adv = x; done = 1; V = ~covered; }
                else { p = x > q ? x : q; V = ~covered & (p >= 32 ? 0xffffffffu : ((1u << p) - 1u)); if (p >= 32) { done = 1; } }
            }
        } else {
            p = q; V = belowQ; if (p >= 32) done = 1;
        }
        /* ---- sequential part ---- */
        while (!done) {
            ++g_seq_iters;
            const uint32_t fromP = 0xffffffffu << p;
            const uint32_t stop = (fixedAcc | depends) & fromP;
            if (!stop) { V |= fromP; break; }
            const int f = __builtin_ctz(stop);
            V |= fromP & ((1u << f) - 1u);
            int m, d, amb = 0;
            if ((depends >> f) & 1u) {
                const uint32_t elig = lower[f] & V;
                if (!elig) { m = mOld[f]; d = dOld[f]; }
                else if (31 - __builtin_clz(elig) == lowN[f]) { m = mLow[f]; d = f - lowN[f]; }
                else { amb = 1; m = d = 0; }
            } else { m = mOld[f]; d = dOld[f]; }
            if (amb) { ++g_amb; adv = f; break; }
            V |= 1u << f;
            int L = m;
            if (L <= 3) { p = f + 1; if (p >= 32) break; continue; }
            starts |= 1u << f;
            if (L == 8) {
                ++g_long;
                const int fi = i0 + f, mx = n - fi < MAX_LENGTH ? n - fi : MAX_LENGTH;
                while (L < mx && c[fi + L] == c[fi - d + L]) ++L;
                if (L > mx) L = mx;
            }
            myLen[f] = L; myDist[f] = d;
            p = f + L;
            if (m == 8 || p >= 32) { adv = p; break; }
        }
        for (int j = 0; j < 32; ++j) {
            if (!valid[j] || !((V >> j) & 1u)) continue;
            if ((starts >> j) & 1u) { if (k < maxTok) { tok[3 * k] = i0 + j; tok[3 * k + 1] = myLen[j]; tok[3 * k + 2] = myDist[j]; } ++k; }
            if ((grp[j] & ~((2u << j) - 1u) & V) == 0) table[h[j]] = i0 + j;
        }
        i0 += adv;
    }
    free(table);
    return k;
}
