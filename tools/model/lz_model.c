/* Sequential model of the fused K-LZ kernel's orchestration (sub-batches, state-space successor function,
 * speculative per-segment chains merged by one true walk).  Test infrastructure for development: validates the
 * algorithm against the oracle's tokens on the CPU before it is written in CUDA.  Not part of the product.
 *   gcc -O2 -shared -fPIC -o tools/model/liblzmodel.so tools/model/lz_model.c */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define MAXM 258
#define MAXD 32768
#define BATCH 16384
#define SUB 8192
#define CAPLEN 32
#define LONGGAP (MAXM - CAPLEN + 1)
#define SUBCAP (SUB + 128 + 96)

typedef struct { int nexcl, excl[4], npatch, patchJ[4], patchD[4]; } Patches;

static int g_spec = 1;
void lzm_set_spec(int v) { g_spec = v; }
static long g_stat_truewalk, g_stat_merged, g_stat_spec;
void lzm_stats(long* o) { o[0] = g_stat_truewalk; o[1] = g_stat_merged; o[2] = g_stat_spec; g_stat_truewalk = g_stat_merged = g_stat_spec = 0; }

static int cand_of(const uint16_t* cand, const Patches* ps, int j)
{
    int d = cand[j];
    for (int k = 0; k < ps->npatch; ++k) if (ps->patchJ[k] == j) d = ps->patchD[k];
    return d;
}

static int effective_cand(const uint16_t* cand, const Patches* ps, int j)
{
    int d = cand[j];
    for (;;) {
        if (d == 0) return 0;
        int p = j - d, hit = 0;
        for (int k = 0; k < ps->nexcl; ++k) hit |= ps->excl[k] == p;
        if (!hit) return d;
        int dd = p > 0 ? cand[p] : 0;
        if (dd == 0) return 0;
        d += dd;
        if (d >= MAXD) return 0;
    }
}

/* 0 = unusable, else 1 + min(fwd, 32) */
static int info_of(const uint8_t* c, int j, int d, int pre)
{
    if (d == 0) return 0;
    const uint8_t* a = c + j; const uint8_t* b = a - d;
    int fwd = 0;
    while (fwd < CAPLEN && a[fwd] == b[fwd]) ++fwd;
    if (fwd >= 4) return fwd + 1;
    int room = j - d + pre, back = 0;
    while (back < 4 && back < room && a[-1 - back] == b[-1 - back]) ++back;
    return fwd + back >= 4 ? fwd + 1 : 0;
}

typedef struct {
    int base, s0, s1, lim;          /* arrays cover [base, base + lim) */
    uint8_t info[SUBCAP + 64];
    uint16_t F[SUBCAP];
    uint16_t dist[SUBCAP + 64];
    uint8_t S[SUBCAP], T[SUBCAP], L[SUBCAP];
} Sub;

/* first position the walk takes from state b, or -1 (positions outside [s0,s1) are masked) */
static int probe_next(const Sub* s, int b)
{
    for (int k = 1; k <= 3; ++k) { int j = b + k; if (j - s->base < s->lim + 64 && j >= s->base && s->info[j - s->base] >= 5 - k) return j; }
    for (int j = b + 4; j < s->s1; ++j) if (j >= s->base && s->info[j - s->base]) return j;
    return -1;
}

static int needs_exact(int fwd, int gap) { return fwd >= CAPLEN || gap >= LONGGAP; }

/* exact token from state x taking position j; returns the new state */
static int exact_token(const uint8_t* c, int pre, int x, int j, int d, int* ms_, int* m_)
{
    const uint8_t* a = c + j; const uint8_t* b = a - d;
    int fwd = 0; while (fwd < MAXM && a[fwd] == b[fwd]) ++fwd;
    int maxBack = j - x; { int room = j - d + pre; if (room < maxBack) maxBack = room; }
    if (maxBack > MAXM) maxBack = MAXM;
    int lb = 0; while (lb < maxBack && a[-1 - lb] == b[-1 - lb]) ++lb;
    int m = fwd + lb; if (m > MAXM) m = MAXM;
    *ms_ = j - lb; *m_ = m;
    return j - lb + m;
}

/* tokens: triples (start, length, distance); returns the count */
int lzm_chunk(const uint8_t* c, int n, int body, int pre, const uint16_t* cand, uint32_t* tok, int maxTok)
{
    (void)n;
    static Sub sub;
    Patches ps; memset(&ps, 0, sizeof ps);
    const int t0 = body > MAXM ? body - MAXM : 0;
    int pos = 0, ntok = 0, fixS = -1;
#define EMIT(ms, m, d) do { if (ntok < maxTok) { tok[3 * ntok] = (uint32_t)(ms); tok[3 * ntok + 1] = (uint32_t)(m); tok[3 * ntok + 2] = (uint32_t)(d); } ++ntok; } while (0)
    while (pos < t0) {
        if (fixS >= 0) {
            int hiJ = fixS + MAXD; if (hiJ > t0) hiJ = t0;
            for (int j = fixS + 1; j < hiJ; ++j)
                if ((int)cand[j] == j - fixS) { int k = ps.npatch++; ps.patchJ[k] = j; ps.patchD[k] = effective_cand(cand, &ps, j); break; }
        }
        int E = pos + BATCH; if (E > t0) E = t0;
        const int B0 = pos + 1;
        int b = B0;
        if (B0 < E) {                                      /* first probe of the batch: j == backRefEnd */
            int d = cand_of(cand, &ps, B0);
            int inf = info_of(c, B0, d, pre);
            if (inf >= 5) { int ms, m; b = exact_token(c, pre, B0, B0, d, &ms, &m); EMIT(ms, m, d); }
        }
        for (int s0 = B0; s0 < E; ) {
            int s1 = s0 + SUB; if (s1 > E) s1 = E;
            if (b >= s1) { s0 = s1; continue; }
            Sub* s = &sub;
            int lowb = b > s0 - 64 ? b : s0 - 64; if (lowb > s0) lowb = s0;
            s->base = lowb & ~31; s->s0 = s0; s->s1 = s1;
            s->lim = ((s1 - s->base + 31) >> 5) * 32;
            memset(s->info, 0, sizeof s->info); memset(s->S, 0, sizeof s->S); memset(s->T, 0, sizeof s->T); memset(s->L, 0, sizeof s->L);
            for (int j = s0; j < s1; ++j) {
                int d = cand_of(cand, &ps, j);
                s->dist[j - s->base] = (uint16_t)d;
                s->info[j - s->base] = (uint8_t)info_of(c, j, d, pre);
            }
            if (b < s->base) {                             /* entry state far behind the arrays: every usable position is acceptable */
                int j = -1;
                for (int q = s0; q < s1; ++q) if (s->info[q - s->base]) { j = q; break; }
                if (j < 0) { s0 = s1; continue; }
                int ms, m; int d = s->dist[j - s->base];
                b = exact_token(c, pre, b, j, d, &ms, &m); EMIT(ms, m, d);
                if (b >= s1) { s0 = s1; continue; }
            }
            /* exact next states of long matches measured so far in this sub-batch */
            int exN = 0; static int exX[4096], exNb[4096];
#define SUCC(x, out) do { unsigned f_ = 0; if ((x) >= B0 && (x) < s1) { int j_ = probe_next(s, (x)); if (j_ >= 0) { int fw_ = s->info[j_ - s->base] - 1; f_ = needs_exact(fw_, j_ - (x)) ? 1u : (unsigned)(j_ + fw_); } } (out) = f_; } while (0)
#define LOOKUP(x, out) do { (out) = 0; for (int q_ = 0; q_ < exN; ++q_) if (exX[q_] == (x)) (out) = (unsigned)exNb[q_]; } while (0)
#define EXACT(x, out) do { int j_ = probe_next(s, (x)); int d_ = s->dist[j_ - s->base], ms_, m_; int nb_ = exact_token(c, pre, (x), j_, d_, &ms_, &m_); \
                           exX[exN] = (x); exNb[exN] = nb_; ++exN; (out) = nb_; } while (0)
            /* speculative chains: 128 lanes (4 warps), lane i owns states [base + 64 i, base + 64 (i+1)) (the last lane also the tail) */
            enum { NL = 128, SEG = 64 };
            int stopState[NL], stopKind[NL], stopTgt[NL], chX[NL];      /* kind: 0 none, 1 ends, 2 breaker, 3 leaves the segment */
            for (int i = 0; i < NL; ++i) { stopKind[i] = 0; chX[i] = s->base + SEG * i; }
            for (int round = 0; g_spec; ++round) {
                int any = 0;
                for (int i = 0; i < NL; ++i) {
                    int segLo = s->base + SEG * i, segHi = i == NL - 1 ? s->base + s->lim : segLo + SEG;
                    if (segLo >= s1 || stopKind[i] != 0) continue;
                    int x = chX[i];
                    for (;;) {
                        s->S[x - s->base] = 1; g_stat_spec++;
                        unsigned f; SUCC(x, f);
                        if (f == 0) { stopKind[i] = 1; stopState[i] = x; break; }
                        if (f == 1) { stopKind[i] = 2; stopState[i] = x; break; }
                        if ((int)f >= segHi || (int)f >= s1) { stopKind[i] = 3; stopState[i] = x; stopTgt[i] = (int)f; break; }
                        x = (int)f;
                    }
                }
                /* per warp: breakers are measured speculatively only while they are few */
                for (int w = 0; w < NL / 32; ++w) {
                    int cnt = 0;
                    for (int l = 0; l < 32; ++l) cnt += stopKind[32 * w + l] == 2 && chX[32 * w + l] >= 0;
                    if (cnt == 0 || cnt > 8) { for (int l = 0; l < 32; ++l) if (stopKind[32 * w + l] == 2) chX[32 * w + l] = -1; continue; }
                    for (int l = 0; l < 32; ++l) {
                        int i = 32 * w + l;
                        if (stopKind[i] != 2 || chX[i] < 0) continue;
                        int segLo = s->base + SEG * i, segHi = i == NL - 1 ? s->base + s->lim : segLo + SEG;
                        unsigned nb; LOOKUP(stopState[i], nb);
                        if (!nb) EXACT(stopState[i], nb);
                        if ((int)nb >= segHi || (int)nb >= s1) { stopKind[i] = 3; stopTgt[i] = (int)nb; }
                        else { stopKind[i] = 0; chX[i] = (int)nb; any = 1; }
                    }
                }
                if (!any) break;
            }
            /* links */
            int linkMp[NL];
            for (int i = 0; i < NL; ++i) {
                linkMp[i] = -1;
                if (!g_spec || i == 0 || stopKind[i - 1] != 3 || stopKind[i] == 0) continue;
                int segLo = s->base + SEG * i, segHi = i == NL - 1 ? s->base + s->lim : segLo + SEG;
                int x = stopTgt[i - 1];
                if (x < segLo || x >= segHi || x >= s1) continue;
                for (;;) {
                    if (s->S[x - s->base]) { linkMp[i] = x; break; }
                    s->L[x - s->base] = 1;
                    unsigned f; SUCC(x, f);
                    if (f == 1) LOOKUP(x, f);
                    if (f == 0 || (int)f >= segHi || (int)f >= s1) break;
                    x = (int)f;
                }
            }
            /* the true walk */
            int mp[NL], linked[NL]; for (int i = 0; i < NL; ++i) { mp[i] = -1; linked[i] = 0; }
            int cur = b, bout = -1;
            for (;;) {
                if (cur >= s1) { bout = cur; break; }
                int i = (cur - s->base) / SEG; if (i > NL - 1) i = NL - 1;
                int x;
                if (s->S[cur - s->base] && stopKind[i] && mp[i] < 0) {
                    mp[i] = cur; g_stat_merged++;
                    while (stopKind[i] == 3 && i + 1 < NL && linkMp[i + 1] >= 0) { ++i; linked[i] = 1; mp[i] = linkMp[i]; }
                    if (stopKind[i] == 1) { bout = stopState[i]; break; }
                    if (stopKind[i] == 3) { cur = stopTgt[i]; continue; }
                    x = stopState[i];                       /* breaker left unmeasured: resolved below */
                } else {
                    x = cur; s->T[x - s->base] = 1; g_stat_truewalk++;
                    unsigned f; SUCC(x, f);
                    if (f == 0) { bout = x; break; }
                    if (f != 1) { cur = (int)f; continue; }
                }
                { unsigned nb; LOOKUP(x, nb); if (!nb) EXACT(x, nb); cur = (int)nb; }
            }
            for (int i = 0; i < NL; ++i) if (mp[i] >= 0) {
                int segLo = s->base + SEG * i, segHi = i == NL - 1 ? s->base + s->lim : segLo + SEG;
                for (int x = mp[i]; x < segHi; ++x) if (s->S[x - s->base]) s->T[x - s->base] = 1;
                if (linked[i]) for (int x = segLo; x < segHi; ++x) if (s->L[x - s->base]) s->T[x - s->base] = 1;
            }
            /* tokens of the sub-batch */
            for (int x = s->base; x < s->base + s->lim; ++x) if (s->T[x - s->base]) {
                unsigned f; SUCC(x, f);
                if (f == 0) continue;                      /* the orbit's last state: no token */
                if (f == 1) LOOKUP(x, f);
                int j = probe_next(s, x);
                int d = s->dist[j - s->base];
                int fwd = s->info[j - s->base] - 1;
                int limit = j - x; { int room = j - d + pre; if (room < limit) limit = room; } if (limit > MAXM) limit = MAXM;
                int lb = 0; while (lb < limit && c[j - 1 - lb] == c[j - d - 1 - lb]) ++lb;
                int ms = j - lb;
                int m = needs_exact(fwd, j - x) ? (int)f - ms : fwd + lb;
                if (m > MAXM) m = MAXM;
                EMIT(ms, m, d);
            }
            b = bout; s0 = s1;
        }
        const int finalB = b;
        const int newpos = finalB > E ? finalB : E;
        fixS = -1;
        if (finalB < E && newpos < t0 && ps.nexcl < 4) { ps.excl[ps.nexcl++] = newpos; fixS = newpos; }
        pos = newpos;
    }
    return ntok;
}
