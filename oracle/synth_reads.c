/*
 * synth_reads.c — the synthetic ONT-like read stream of the benchmark workloads, restated
 * in plain C for the CPU side: bench.py's reference arm and the full-size parity tests generate their
 * inputs with this file, so that those processes never load the product library.
 *
 * TEST INFRASTRUCTURE ONLY (like everything under oracle/).  The reference has no generator (it reads
 * FASTA/FASTQ files, /root/reference/approx_counter.cpp:819-825); this is a second, independent
 * implementation of the repo's own generator (approx_counter_b200/csrc/host/host_util.cpp: synth_read) and
 * tests/test_host.py::test_oracle_synth_equals_host_synth holds the two to byte equality.
 *
 * Read i of stream `seed` is a pure function of (seed, i, sl): xoshiro256** seeded by splitmix64, raw
 * outputs only; length 2*sl + 50 + (u mod 400); body uniform ACGT with N at 1e-4; 90 % of the reads carry the
 * ligation adapter at offset u mod 8 from the start and the bottom adapter ending u mod 8 before the end,
 * each copy through an error channel (sub 3 %, ins 2 %, del 3 % per base).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s[4]; } rng_t;

static uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static void rng_init(rng_t *g, uint64_t seed, uint64_t index) {
    uint64_t x = seed * 0xd1342543de82ef95ULL + index * 0x9e3779b97f4a7c15ULL + 0x2545f4914f6cdd1dULL;
    for (int i = 0; i < 4; i++) g->s[i] = splitmix(&x);
}

static uint64_t rng_next(rng_t *g) {
    uint64_t *s = g->s;
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}

static const char LETTERS[] = "ACGT";
static const char ADAPTER_TOP[] = "AATGTACTTCGTTCAGTTACGTATTGCT";
static const char ADAPTER_BOTTOM[] = "GCAATACGTAACTGAACGAAGT";

/* one noisy copy of `adapter` into out (at most 2 * strlen(adapter) letters); returns its length */
static size_t noisy_copy(rng_t *g, const char *adapter, char *out) {
    size_t n = 0;
    for (const char *p = adapter; *p; p++) {
        const uint64_t r = rng_next(g) % 100;
        if (r < 3) { /* substitution by one of the three other letters */
            const int b = (int)(strchr(LETTERS, *p) - LETTERS);
            out[n++] = LETTERS[(b + 1 + rng_next(g) % 3) & 3];
        } else if (r < 5) { /* insertion in front of the base */
            out[n++] = LETTERS[rng_next(g) & 3];
            out[n++] = *p;
        } else if (r >= 8) { /* 5..7: deleted */
            out[n++] = *p;
        }
    }
    return n;
}

/* read `index` of the stream into buf (capacity >= 2*sl + 450); returns its length */
static size_t synth_one(uint64_t seed, uint64_t index, uint32_t sl, char *buf) {
    rng_t g;
    rng_init(&g, seed, index);
    const size_t len = 2 * (size_t)sl + 50 + rng_next(&g) % 400;
    for (size_t i = 0; i < len; i++) {
        const uint64_t r = rng_next(&g);
        buf[i] = (r % 10000 == 0) ? 'N' : LETTERS[(r >> 20) & 3];
    }
    /* bit 40 of the seed: the "wide" variant of a stream, adapter offsets uniform in 0..sl/2 instead of 0..7 */
    const uint64_t off_range = (seed >> 40) & 1 ? (uint64_t)sl / 2 + 1 : 8;
    if (rng_next(&g) % 10 != 0) {
        char copy[64];
        const size_t off_top = rng_next(&g) % off_range;
        size_t n = noisy_copy(&g, ADAPTER_TOP, copy);
        if (off_top + n <= len) memcpy(buf + off_top, copy, n);
        const size_t off_bottom = rng_next(&g) % off_range;
        n = noisy_copy(&g, ADAPTER_BOTTOM, copy);
        if (off_bottom + n <= len) memcpy(buf + (len - off_bottom - n), copy, n);
    }
    return len;
}

/* sampled ends of reads [first, first+n): n rows of sl ASCII bytes (bot = 0: the first sl bases, :466) or
 * sl+1 bytes (bot = 1: the last sl+1 bases, :463) */
int orc_synth_ends(uint64_t seed, uint64_t first, uint64_t n, uint32_t sl, int bot, uint8_t *out) {
    const size_t row = (size_t)sl + (bot ? 1 : 0);
    int ok = 1;
#pragma omp parallel
    {
        char *buf = (char *)malloc(2 * (size_t)sl + 512);
        if (!buf) {
#pragma omp atomic write
            ok = 0;
        }
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; i++) {
            if (!buf) continue;
            const size_t len = synth_one(seed, first + (uint64_t)i, sl, buf);
            memcpy(out + (size_t)i * row, bot ? buf + (len - 1 - sl) : buf, row);
        }
        free(buf);
    }
    return ok ? 0 : -1;
}
