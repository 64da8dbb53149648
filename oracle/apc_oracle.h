/*
 * apc_oracle.h — CPU restatement of approx_counter's counting path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker / the timed CPU
 * baseline.  The product path (approx_counter_b200/csrc) never links it.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors,
 * and its approximate search is SeqAn's find<0,2>(…, EditDistance()), an
 * un-vendored header-only dependency (SeqAn >= 2.4.0, reference README.md:14)
 * that is absent from this image, so the reference binary cannot be built
 * here.  The approximate stage follows the closed form derived in SURVEY.md
 * §0.1 and is cross-checked against oracle/seqan_model.cpp (a literal
 * restatement of SeqAn's published search-scheme recursion).
 *
 * All file:line citations are into /root/reference/approx_counter.cpp.
 *
 * Sample representation: Dna5 ordinals (A=0 C=1 G=2 T=3 N=4), reads
 * concatenated in `codes`, read r occupies codes[offs[r] .. offs[r+1]).
 */
#ifndef APC_ORACLE_H
#define APC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* :55-62  2-bit pack, first base most significant.  seq holds ordinals 0..3. */
uint64_t orc_dna2int(const uint8_t *seq, int k);
/* :70-78  inverse; writes k ASCII letters + NUL. */
void orc_int2dna(uint64_t value, int k, char *out);
/* ASCII -> Dna5 ordinal (SeqAn Dna5 conversion: ACGT case-insensitive, else N). */
uint8_t orc_char2code(char c);

/* :183-186 */
float orc_adjust_threshold(float c_old, uint8_t k_old, uint8_t k_new);
/* :214-234  (float32 arithmetic, `>=`). */
int orc_have_low_complexity(uint64_t kmer, uint8_t k, float threshold);
/* :247-267 */
float orc_get_complexity(uint64_t kmer, uint8_t k);
/* integer numerator of the score: sum_v v*(v-1) over the 16 dimer bins. */
uint64_t orc_dimer_sum(uint64_t kmer, uint8_t k);
void orc_dimer_sums(const uint64_t *kmers, uint64_t n, uint8_t k, uint32_t *out);
/* :283-302  returns 1 iff a sorts before b. */
int orc_compare_count(uint64_t a_kmer, uint64_t a_count, uint64_t b_kmer,
                      uint64_t b_count, int k);

/* :487-519  exact k-mer count.  Returns number of distinct k-mers; *keys and
 * *counts are malloc'ed (caller frees), in unspecified order (like the
 * unordered_map).  forbidden may be NULL.  *had_n = windows skipped for N. */
uint64_t orc_count_kmers(const uint8_t *codes, const uint64_t *offs,
                         uint64_t n_reads, uint8_t k, float threshold,
                         const uint64_t *forbidden, uint64_t n_forbidden,
                         uint64_t **keys, uint64_t **counts, uint64_t *had_n);
/* same multiset with all host threads (set-up of the reference arm / full-size tests only) */
uint64_t orc_count_kmers_mt(const uint8_t *codes, const uint64_t *offs, uint64_t n_reads, uint8_t k,
                            float threshold, int nb_thread, uint64_t **keys, uint64_t **counts,
                            uint64_t *had_n_out);

/* :396-405  sort (kmer,count) pairs by CompareCount, keep first `limit`.
 * Sorts in place; returns the kept length. */
uint64_t orc_get_most_frequent(uint64_t *keys, uint64_t *counts, uint64_t n,
                               uint64_t limit, int k);

/* :372-388  solid k-mers: count >= solid_km, sorted by count desc (ties in the
 * reference are unspecified; here broken by CompareCount). */
uint64_t orc_get_solid_kmers(uint64_t *keys, uint64_t *counts, uint64_t n,
                             uint64_t solid_km, int k);

/* :531-601  approximate count, textbook O(k*L) DP per (k-mer, read).
 * counts_out[i] = sum_e popcount(tcount[e]) for kmers[i]. */
void orc_error_count(const uint8_t *codes, const uint64_t *offs,
                     uint64_t n_reads, const uint64_t *kmers, uint64_t n_kmers,
                     uint8_t k, uint64_t *counts_out);
/* Same result; Myers bit-vector inner loop + OpenMP over k-mers (mirrors the
 * reference's `omp for schedule(dynamic)` at :567).  Used as the timed CPU
 * baseline.  nb_thread <= 0 -> omp default. */
void orc_error_count_fast(const uint8_t *codes, const uint64_t *offs,
                          uint64_t n_reads, const uint64_t *kmers,
                          uint64_t n_kmers, uint8_t k, int nb_thread,
                          uint64_t *counts_out);
/* per-read minimum infix edit distance (clamped to 3), for property tests. */
int orc_min_infix_distance(const uint8_t *text, uint64_t len, uint64_t kmer,
                           uint8_t k);

/* :415-476 given the shuffled id order `perm` (the reference draws it from
 * mt19937(random_device), :427-429).  Reads as codes/offs; writes the sample
 * to out_codes/out_offs (caller-sized: nb_sample*(cut+1) codes, nb_sample+1
 * offsets).  Returns number of reads sampled. */
uint64_t orc_sample_sequences(const uint8_t *codes, const uint64_t *offs,
                              uint64_t n_reads, const uint64_t *perm,
                              uint64_t nb_sample, uint64_t cut_size, int bot,
                              uint8_t *out_codes, uint64_t *out_offs);

/* :157-174  "<KMER>\t<count>\n" per entry.  Returns 1 on success. */
int orc_export_counter(const uint64_t *keys, const uint64_t *counts, uint64_t n,
                       uint8_t k, const char *path);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
