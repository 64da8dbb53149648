/*
 * apc_host.h — C view of the host-side (CPU, no GPU needed) pieces of the
 * drop-in `approx_counter` binary: codec, threshold adjustment, CompareCount
 * ordering, FASTA/FASTQ reading, sampling, file export, the synthetic read
 * generator used by tests/bench, and the CLI entry itself.  These live in
 * libapc.so next to the CUDA path so that the parity tests can drive exactly
 * the code the binary runs.  Citations: /root/reference/approx_counter.cpp.
 */
#ifndef APC_HOST_H
#define APC_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* :55-62 — ASCII k-mer -> 2-bit value.  Returns 0, or -1 if seq holds a
 * character other than ACGTacgt or k > 32. */
int apch_dna2int(const char *seq, uint32_t k, uint64_t *out);
/* :70-78 — writes k letters and a NUL. */
void apch_int2dna(uint64_t value, uint32_t k, char *out);
/* :183-186 */
float apch_adjust_threshold(float c_old, uint8_t k_old, uint8_t k_new);
/* :247-267 / :214-234 */
float apch_get_complexity(uint64_t kmer, uint8_t k);
int apch_have_low_complexity(uint64_t kmer, uint8_t k, float threshold);
/* Smallest dimer sum that the filter (:232-233) rejects for this k and
 * threshold: the integer the device kernels compare against. */
uint32_t apch_lc_min_filtered_sum(uint8_t k, float threshold);
/* :396-405 — sort pairs by CompareCount (:275-305) in place, return
 * min(n, limit). */
uint64_t apch_get_most_frequent(uint64_t *kmers, uint64_t *counts, uint64_t n,
                                uint64_t limit, int k);
/* :157-174 — returns 1 on success, 0 if the file cannot be opened. */
int apch_export_counter(const uint64_t *kmers, const uint64_t *counts,
                        uint64_t n, uint8_t k, const char *path);
/* :340-364 — one k-mer per line, lines with non-ACGT characters skipped.
 * Returns the number parsed (writes at most `capacity`), or -1 if unreadable. */
int64_t apch_parse_kmer_list(const char *path, uint64_t *out, uint64_t capacity);

/* :819-825 — whole FASTA/FASTQ file into memory (format auto-detected,
 * multi-line records accepted, ids and qualities dropped). */
typedef struct apch_reads apch_reads;
int apch_reads_load(const char *path, apch_reads **out);
uint64_t apch_reads_count(const apch_reads *r);
uint64_t apch_reads_length(const apch_reads *r, uint64_t i);
const char *apch_reads_seq(const apch_reads *r, uint64_t i);
/* 1 when the reads are views into a read-only mapping of the file (large files whose records keep their sequence
 * on one line are parsed without copying), 0 when they were compacted into a buffer. */
int apch_reads_mapped(const apch_reads *r);
void apch_reads_free(apch_reads *r);

/* :415-476 — shuffle read ids (std::shuffle over std::mt19937; seeded from
 * std::random_device like the reference when seed < 0), walk them until
 * nb_sample reads of length >= 2*cut are taken; start = first `cut` bases,
 * end = last `cut`+1 bases (:463).  `out` receives n_sampled rows of
 * cut (+1 if bot) ASCII bytes. */
int apch_sample(const apch_reads *r, uint64_t nb_sample, uint64_t cut, int bot,
                int64_t seed, uint8_t *out, uint64_t *n_sampled);

/* :423-429 — the shuffled read ids alone: out[0..n) = 0..n-1 through
 * std::shuffle over std::mt19937 seeded like apch_sample (seed < 0:
 * std::random_device).  Feeds apc_sample_resident, which walks the ids on the
 * device. */
int apch_shuffle_order(uint64_t n, int64_t seed, uint32_t *out);

/* Synthetic ONT-like reads with planted adapters: read i of
 * stream `seed` is a pure function of (seed, i, sl).  apch_synth_ends writes
 * the sampled ends of reads [first, first+n) directly (n rows of sl, or sl+1
 * if bot); apch_synth_write writes the same reads as FASTA (or FASTQ).  Bit 40
 * of `seed` selects the wide variant of a stream: adapter offsets uniform in
 * 0..sl/2 instead of 0..7. */
int apch_synth_ends(uint64_t seed, uint64_t first, uint64_t n, uint32_t sl,
                    int bot, uint8_t *out);
int apch_synth_write(const char *path, uint64_t seed, uint64_t n, uint32_t sl,
                     int fastq);

/* The drop-in binary's main(): same flags and files as the reference
 * (:604-669, :691-958) plus --seed/--gpus/--device extensions. */
int apch_cli_main(int argc, const char **argv);

#ifdef __cplusplus
}
#endif
#endif /* APC_HOST_H */
