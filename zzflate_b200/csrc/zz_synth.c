/* Synthetic workload generators of SURVEY 8(d) (bench / test support, not part of the encode path).
 *
 *   zz_synth_markov : order-2 Markov text.  Independent 1 MiB segments, segment s seeded 0x5EED0001+s,
 *                     start state = first two training bytes, next symbol = inverse CDF of
 *                     splitmix64() % rowTotal over the context's successor counts.
 *   zz_synth_random : splitmix64 byte stream.
 * Both are deterministic and thread-count independent (pthread pool over segments).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

static inline uint64_t splitmix64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct {
    uint8_t* dst; size_t n; uint64_t seg0; size_t segBytes;
    const uint32_t* rowOff; const uint8_t* syms; const uint32_t* cum;   /* cum: inclusive cumulative counts per row */
    uint8_t s0, s1;
    size_t next; size_t nseg; pthread_mutex_t mu;
    uint64_t seed; int kind;
} job_t;

static void gen_markov_segment(const job_t* j, size_t seg)
{
    size_t lo = seg * j->segBytes, hi = lo + j->segBytes;
    if (hi > j->n) hi = j->n;
    uint64_t st = 0x5EED0001ull + j->seg0 + seg;
    unsigned c1 = j->s0, c2 = j->s1;
    for (size_t i = lo; i < hi; ++i) {
        unsigned ctx = (c1 << 8) | c2;
        uint32_t a = j->rowOff[ctx], b = j->rowOff[ctx + 1];
        uint32_t total = j->cum[b - 1];
        uint32_t r = (uint32_t)(splitmix64(&st) % total);
        while (j->cum[a] <= r) ++a;                 /* rows are short: linear inverse CDF */
        uint8_t sym = j->syms[a];
        j->dst[i] = sym;
        c1 = c2; c2 = sym;
    }
}

static void gen_random_segment(const job_t* j, size_t seg)
{
    size_t lo = seg * j->segBytes, hi = lo + j->segBytes;
    if (hi > j->n) hi = j->n;
    /* word k of the stream is splitmix64 step k+1 from the seed: jump straight to the segment */
    uint64_t st = j->seed + (uint64_t)(lo / 8) * 0x9E3779B97F4A7C15ull;
    size_t i = lo;
    for (; i + 8 <= hi; i += 8) { uint64_t z = splitmix64(&st); memcpy(j->dst + i, &z, 8); }
    if (i < hi) { uint64_t z = splitmix64(&st); memcpy(j->dst + i, &z, hi - i); }
}

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t seg = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (seg >= j->nseg) break;
        if (j->kind == 0) gen_markov_segment(j, seg); else gen_random_segment(j, seg);
    }
    return NULL;
}

static void run(job_t* j, int threads)
{
    j->nseg = (j->n + j->segBytes - 1) / j->segBytes;
    j->next = 0;
    pthread_mutex_init(&j->mu, NULL);
    if (threads < 1) threads = 1;
    if ((size_t)threads > j->nseg) threads = (int)(j->nseg ? j->nseg : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, worker, j);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&j->mu);
}

/* n bytes starting at segment seg0 (so shards of one long stream can be generated independently) */
void zz_synth_markov(uint8_t* dst, size_t n, uint64_t seg0, const uint32_t* rowOff, const uint8_t* syms,
                     const uint32_t* cum, uint8_t s0, uint8_t s1, int threads)
{
    job_t j; memset(&j, 0, sizeof j);
    j.dst = dst; j.n = n; j.seg0 = seg0; j.segBytes = 1u << 20;
    j.rowOff = rowOff; j.syms = syms; j.cum = cum; j.s0 = s0; j.s1 = s1; j.kind = 0;
    run(&j, threads);
}

void zz_synth_random(uint8_t* dst, size_t n, uint64_t seed, int threads)
{
    job_t j; memset(&j, 0, sizeof j);
    j.dst = dst; j.n = n; j.segBytes = 1u << 20; j.seed = seed; j.kind = 1;
    run(&j, threads);
}
