/*
 * apc_oracle.c — CPU restatement of approx_counter's counting path.
 *
 * TEST INFRASTRUCTURE ONLY (see apc_oracle.h).  PARITY UNPINNED: no reference
 * golden vectors exist and the reference cannot be built here (SeqAn absent).
 * Citations are /root/reference/approx_counter.cpp:line.
 */
#include "apc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const char ORC_DNA[4] = {'A', 'C', 'G', 'T'}; /* :22 */
#define ORC_MAXERR 2                                  /* :25 */

uint8_t orc_char2code(char c) {
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* :55-62 */
uint64_t orc_dna2int(const uint8_t *seq, int k) {
    uint64_t value = 0;
    for (int i = 0; i < k; i++) value = value << 2 | (uint8_t)seq[i];
    return value;
}

/* :70-78 (builds the string back to front) */
void orc_int2dna(uint64_t value, int k, char *out) {
    for (int i = k - 1; i >= 0; i--) {
        out[i] = ORC_DNA[value & 3];
        value >>= 2;
    }
    out[k] = '\0';
}

/* :183-186  double pow ratio -> float -> float multiply */
float orc_adjust_threshold(float c_old, uint8_t k_old, uint8_t k_new) {
    float c_new = c_old * (float)(pow((double)(k_new - 2 + 1), 2.0) /
                                  pow((double)(k_old - 2 + 1), 2.0));
    return c_new;
}

/* shared body of :216-231 / :249-264 */
uint64_t orc_dimer_sum(uint64_t kmer, uint8_t k) {
    uint64_t counts[16] = {0};
    for (int i = 0; i < k - 1; i++) {
        uint8_t c = kmer & 15;
        kmer >>= 2;
        counts[c]++;
    }
    uint64_t sum = 0;
    for (int v = 0; v < 16; v++) sum += counts[v] * (counts[v] - 1);
    return sum;
}

/* the same over an array (test-side accelerator for top-N selection over 1e8 distinct k-mers) */
void orc_dimer_sums(const uint64_t *kmers, uint64_t n, uint8_t k, uint32_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) out[i] = (uint32_t)orc_dimer_sum(kmers[i], k);
}

/* :214-234 */
int orc_have_low_complexity(uint64_t kmer, uint8_t k, float threshold) {
    float s = (float)orc_dimer_sum(kmer, k) / (float)(2 * (k - 2));
    return s >= threshold;
}

/* :247-267 */
float orc_get_complexity(uint64_t kmer, uint8_t k) {
    return (float)orc_dimer_sum(kmer, k) / (float)(2 * (k - 2));
}

/* :283-302 */
int orc_compare_count(uint64_t a_kmer, uint64_t a_count, uint64_t b_kmer,
                      uint64_t b_count, int k) {
    if (a_count == b_count) {
        float a_comp = orc_get_complexity(a_kmer, (uint8_t)k);
        float b_comp = orc_get_complexity(b_kmer, (uint8_t)k);
        if (a_comp == b_comp) return a_kmer > b_kmer;
        return a_comp < b_comp;
    }
    return a_count > b_count;
}

/* ---- tiny open-addressing u64->u64 map standing in for std::unordered_map :33 */
typedef struct {
    uint64_t *keys, *vals;
    uint8_t *used;
    uint64_t cap, n;
} orc_map;

static uint64_t orc_mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static void orc_map_init(orc_map *m, uint64_t cap) {
    m->cap = cap; m->n = 0;
    m->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    m->vals = (uint64_t *)malloc(cap * sizeof(uint64_t));
    m->used = (uint8_t *)calloc(cap, 1);
}
static void orc_map_free(orc_map *m) { free(m->keys); free(m->vals); free(m->used); }
static void orc_map_add(orc_map *m, uint64_t key, uint64_t delta);
static void orc_map_grow(orc_map *m) {
    orc_map old = *m;
    orc_map_init(m, old.cap * 2);
    for (uint64_t i = 0; i < old.cap; i++)
        if (old.used[i]) orc_map_add(m, old.keys[i], old.vals[i]);
    orc_map_free(&old);
}
static void orc_map_add(orc_map *m, uint64_t key, uint64_t delta) {
    if ((m->n + 1) * 10 > m->cap * 7) orc_map_grow(m);
    uint64_t i = orc_mix(key) & (m->cap - 1);
    while (m->used[i] && m->keys[i] != key) i = (i + 1) & (m->cap - 1);
    if (!m->used[i]) { m->used[i] = 1; m->keys[i] = key; m->vals[i] = 0; m->n++; }
    m->vals[i] += delta;
}

static int orc_cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* :487-519 */
uint64_t orc_count_kmers(const uint8_t *codes, const uint64_t *offs,
                         uint64_t n_reads, uint8_t k, float threshold,
                         const uint64_t *forbidden, uint64_t n_forbidden,
                         uint64_t **keys, uint64_t **counts, uint64_t *had_n_out) {
    orc_map count;
    orc_map_init(&count, 1u << 16);
    uint64_t had_n = 0;
    uint64_t *forb = NULL;
    if (n_forbidden) { /* std::set<uint64_t> :44 -> sorted array + bsearch */
        forb = (uint64_t *)malloc(n_forbidden * sizeof(uint64_t));
        memcpy(forb, forbidden, n_forbidden * sizeof(uint64_t));
        qsort(forb, n_forbidden, sizeof(uint64_t), orc_cmp_u64);
    }
    for (uint64_t r = 0; r < n_reads; r++) {
        const uint8_t *seq = codes + offs[r];
        uint64_t len = offs[r + 1] - offs[r];
        for (uint64_t i = 0; i + k <= len; i++) { /* :496 */
            int is_dna = 1; /* :313-321 */
            for (int j = 0; j < k; j++)
                if (seq[i + j] >= 4) { is_dna = 0; break; }
            if (is_dna) {
                uint64_t n = orc_dna2int(seq + i, k); /* :499 */
                int forbidden_hit =
                    forb && bsearch(&n, forb, n_forbidden, sizeof(uint64_t), orc_cmp_u64);
                if (!orc_have_low_complexity(n, k, threshold) && !forbidden_hit)
                    orc_map_add(&count, n, 1); /* :502 */
            } else {
                had_n++; /* :506 */
            }
        }
    }
    free(forb);
    uint64_t n = count.n, j = 0;
    *keys = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    *counts = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < count.cap; i++)
        if (count.used[i]) { (*keys)[j] = count.keys[i]; (*counts)[j] = count.vals[i]; j++; }
    orc_map_free(&count);
    if (had_n_out) *had_n_out = had_n;
    return n;
}

/* The same count with all host threads — set-up of bench.py's reference arm and of the full-size parity tests
 * only (the reference's count_kmers is single-threaded, :487-519, and so is orc_count_kmers above; at C3/C4
 * size that is minutes per end).  Every thread walks all reads with a rolling 2-bit window and keeps the
 * k-mers of its own hash shard in its own map, so no two threads ever count the same k-mer; the shards are
 * concatenated.  tests/test_oracle.py holds it equal to orc_count_kmers as a multiset. */
uint64_t orc_count_kmers_mt(const uint8_t *codes, const uint64_t *offs, uint64_t n_reads, uint8_t k,
                            float threshold, int nb_thread, uint64_t **keys, uint64_t **counts,
                            uint64_t *had_n_out) {
    int T = nb_thread > 0 ? nb_thread : omp_get_max_threads();
    if (T < 1) T = 1;
    orc_map *maps = (orc_map *)malloc((size_t)T * sizeof(orc_map));
    uint64_t had_n = 0;
    const uint64_t mask = k < 32 ? ((1ull << (2 * k)) - 1) : ~0ull;
#pragma omp parallel num_threads(T) reduction(+ : had_n)
    {
        const int t = omp_get_thread_num();
        const int nt = omp_get_num_threads();
        if (t < T) {
            orc_map_init(&maps[t], 1u << 16);
            for (uint64_t r = 0; r < n_reads; r++) {
                const uint8_t *seq = codes + offs[r];
                const uint64_t len = offs[r + 1] - offs[r];
                uint64_t w = 0, run = 0; /* run = letters since the last N */
                for (uint64_t i = 0; i < len; i++) {
                    if (seq[i] >= 4) { run = 0; w = 0; } else { w = ((w << 2) | seq[i]) & mask; run++; }
                    if (i + 1 < k) continue;
                    if (run < k) { if (t == 0) had_n++; continue; } /* window holds an N (:506) */
                    uint64_t h = w * 0x9E3779B97F4A7C15ull;
                    h ^= h >> 29;
                    if ((int)(h % (uint64_t)nt) != t) continue;
                    if (!orc_have_low_complexity(w, k, threshold)) orc_map_add(&maps[t], w, 1);
                }
            }
        }
        if (nt < T && t == 0) { /* fewer threads than asked for: the missing shards stay empty */
            for (int m = nt; m < T; m++) orc_map_init(&maps[m], 16);
        }
    }
    uint64_t n = 0;
    for (int t = 0; t < T; t++) n += maps[t].n;
    *keys = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    *counts = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    uint64_t j = 0;
    for (int t = 0; t < T; t++) {
        for (uint64_t i = 0; i < maps[t].cap; i++)
            if (maps[t].used[i]) { (*keys)[j] = maps[t].keys[i]; (*counts)[j] = maps[t].vals[i]; j++; }
        orc_map_free(&maps[t]);
    }
    free(maps);
    if (had_n_out) *had_n_out = had_n;
    return n;
}

/* ---- :396-405 */
typedef struct { uint64_t kmer, count; } orc_pair;
static int orc_sort_k;
static int orc_pair_cmp(const void *pa, const void *pb) {
    const orc_pair *a = (const orc_pair *)pa, *b = (const orc_pair *)pb;
    if (orc_compare_count(a->kmer, a->count, b->kmer, b->count, orc_sort_k)) return -1;
    if (orc_compare_count(b->kmer, b->count, a->kmer, a->count, orc_sort_k)) return 1;
    return 0;
}
static void orc_sort_pairs(uint64_t *keys, uint64_t *counts, uint64_t n, int k) {
    orc_pair *v = (orc_pair *)malloc((n ? n : 1) * sizeof(orc_pair));
    for (uint64_t i = 0; i < n; i++) { v[i].kmer = keys[i]; v[i].count = counts[i]; }
    orc_sort_k = k;
    qsort(v, n, sizeof(orc_pair), orc_pair_cmp); /* std::sort(CompareCount(k)) :400 */
    for (uint64_t i = 0; i < n; i++) { keys[i] = v[i].kmer; counts[i] = v[i].count; }
    free(v);
}
uint64_t orc_get_most_frequent(uint64_t *keys, uint64_t *counts, uint64_t n,
                               uint64_t limit, int k) {
    orc_sort_pairs(keys, counts, n, k);
    return n > limit ? limit : n; /* :401-403 */
}

/* :372-388 */
uint64_t orc_get_solid_kmers(uint64_t *keys, uint64_t *counts, uint64_t n,
                             uint64_t solid_km, int k) {
    orc_sort_pairs(keys, counts, n, k);
    uint64_t limit = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (counts[i] >= solid_km) limit++;
        else break;
    }
    return limit;
}

/* ---- approximate stage ------------------------------------------------- */

/* Minimum, over all substrings of `text` (including the empty one), of the
 * unit-cost Levenshtein distance to the k-mer; N in the text matches nothing.
 * Clamped to ORC_MAXERR+1.  This is d_r of SURVEY.md §0.1. */
int orc_min_infix_distance(const uint8_t *text, uint64_t len, uint64_t kmer,
                           uint8_t k) {
    uint8_t pat[32];
    int col[33];
    for (int i = 0; i < k; i++) pat[i] = (kmer >> (2 * (k - 1 - i))) & 3;
    for (int i = 0; i <= k; i++) col[i] = i; /* empty text prefix */
    int best = col[k];
    for (uint64_t j = 0; j < len; j++) {
        int diag = col[0]; /* D[0][j-1] = 0 */
        col[0] = 0;        /* free start anywhere in the read */
        for (int i = 1; i <= k; i++) {
            int sub = diag + ((text[j] >= 4 || text[j] != pat[i - 1]) ? 1 : 0);
            int del = col[i - 1] + 1; /* k-mer char unmatched */
            int ins = col[i] + 1;     /* text char unmatched  */
            diag = col[i];
            int v = sub < del ? sub : del;
            col[i] = v < ins ? v : ins;
        }
        if (col[k] < best) best = col[k];
    }
    return best > ORC_MAXERR + 1 ? ORC_MAXERR + 1 : best;
}

/* :531-601.  The reference keeps, per k-mer, three read bitsets tcount[e]
 * (:553, :580-582); SeqAn's delegate sets tcount[errors][read] for every
 * occurrence (:556-565) and the result is the sum of the three popcounts
 * (:590-596).  With d_r the minimum infix distance, read r ends up flagged at
 * every level e >= d_r (oracle/seqan_model.cpp). */
void orc_error_count(const uint8_t *codes, const uint64_t *offs,
                     uint64_t n_reads, const uint64_t *kmers, uint64_t n_kmers,
                     uint8_t k, uint64_t *counts_out) {
    uint8_t *tcount[ORC_MAXERR + 1];
    for (int e = 0; e <= ORC_MAXERR; e++) tcount[e] = (uint8_t *)malloc(n_reads ? n_reads : 1);
    for (uint64_t q = 0; q < n_kmers; q++) {
        for (int e = 0; e <= ORC_MAXERR; e++) memset(tcount[e], 0, n_reads); /* :580-582 */
        for (uint64_t r = 0; r < n_reads; r++) {
            int d = orc_min_infix_distance(codes + offs[r], offs[r + 1] - offs[r], kmers[q], k);
            for (int e = d; e <= ORC_MAXERR; e++) tcount[e][r] = 1;
        }
        uint64_t total = 0; /* :590-593 */
        for (int e = 0; e <= ORC_MAXERR; e++)
            for (uint64_t r = 0; r < n_reads; r++) total += tcount[e][r];
        counts_out[q] = total; /* :596 */
    }
    for (int e = 0; e <= ORC_MAXERR; e++) free(tcount[e]);
}

/* Myers 1999 bit-vector column update, semi-global (free text start). */
void orc_error_count_fast(const uint8_t *codes, const uint64_t *offs,
                          uint64_t n_reads, const uint64_t *kmers,
                          uint64_t n_kmers, uint8_t k, int nb_thread,
                          uint64_t *counts_out) {
#ifdef _OPENMP
    if (nb_thread > 0) omp_set_num_threads(nb_thread); /* :547 */
#else
    (void)nb_thread;
#endif
    const uint64_t top = 1ULL << (k - 1);
#pragma omp parallel for schedule(dynamic) /* :567 */
    for (uint64_t q = 0; q < n_kmers; q++) {
        uint64_t peq[5] = {0, 0, 0, 0, 0}; /* peq[4] (N) stays 0 */
        for (int i = 0; i < k; i++)
            peq[(kmers[q] >> (2 * (k - 1 - i))) & 3] |= 1ULL << i;
        uint64_t total = 0;
        for (uint64_t r = 0; r < n_reads; r++) {
            const uint8_t *t = codes + offs[r];
            uint64_t len = offs[r + 1] - offs[r];
            uint64_t pv = ~0ULL, mv = 0;
            int score = k, best = k;
            for (uint64_t j = 0; j < len; j++) {
                uint64_t eq = peq[t[j] > 4 ? 4 : t[j]];
                uint64_t xv = eq | mv;
                uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
                uint64_t ph = mv | ~(xh | pv);
                uint64_t mh = pv & xh;
                if (ph & top) score++;
                if (mh & top) score--;
                ph <<= 1; mh <<= 1;
                pv = mh | ~(xv | ph);
                mv = ph & xv;
                if (score < best) best = score;
            }
            if (best <= ORC_MAXERR) total += (uint64_t)(ORC_MAXERR + 1 - best);
        }
        counts_out[q] = total;
    }
}

/* :415-476 */
uint64_t orc_sample_sequences(const uint8_t *codes, const uint64_t *offs,
                              uint64_t n_reads, const uint64_t *perm,
                              uint64_t nb_sample, uint64_t cut_size, int bot,
                              uint8_t *out_codes, uint64_t *out_offs) {
    uint64_t nb_seq = 0, i = 0, w = 0;
    out_offs[0] = 0;
    while (nb_seq < nb_sample && i < n_reads) { /* :447 */
        uint64_t seq_id = perm[i];
        uint64_t len = offs[seq_id + 1] - offs[seq_id];
        const uint8_t *seq = codes + offs[seq_id];
        uint64_t cur = len < cut_size ? len : cut_size; /* :453 */
        if (len >= cut_size * 2) {                      /* :461 */
            uint64_t from, n;
            if (bot) { from = len - 1 - cur; n = len - from; } /* suffix(seq, len-1-cut) :463 -> cut+1 bases */
            else { from = 0; n = cur; }                        /* prefix(seq, cut) :466 */
            memcpy(out_codes + w, seq + from, n);
            w += n;
            nb_seq++;
            out_offs[nb_seq] = w;
        }
        i++;
    }
    return nb_seq;
}

/* :157-174 */
int orc_export_counter(const uint64_t *keys, const uint64_t *counts, uint64_t n,
                       uint8_t k, const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) return 0;
    char buf[40];
    for (uint64_t i = 0; i < n; i++) {
        orc_int2dna(keys[i], k, buf);
        fprintf(f, "%s\t%llu\n", buf, (unsigned long long)counts[i]); /* :165 */
    }
    fclose(f);
    return 1;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
