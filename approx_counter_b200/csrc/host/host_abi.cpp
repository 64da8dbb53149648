// host_abi.cpp — C exports of the host-side pieces (include/apc_host.h).
#include <cstring>
#include <new>
#include <string>

#include "apc_host.h"
#include "host_util.h"

struct apch_reads : public apch::Reads {};

// no C++ exception may cross the C ABI (apch_cli_main excepted: like the reference, an invalid
// -k / -sl aborts the process through an uncaught std::invalid_argument, :781-787)
#define APCH_GUARD(fail_value, ...)      \
    try {                                \
        __VA_ARGS__                      \
    } catch (...) {                      \
        return fail_value;               \
    }

extern "C" {

int apch_dna2int(const char *seq, uint32_t k, uint64_t *out) {
    APCH_GUARD(-1,
    uint64_t v = 0;
    if (!seq || !out || !apch::dna2int(seq, k, v)) return -1;
    *out = v;
    return 0;
    )
}

void apch_int2dna(uint64_t value, uint32_t k, char *out) {
    const std::string s = apch::int2dna(value, k);
    std::memcpy(out, s.c_str(), s.size() + 1);
}

float apch_adjust_threshold(float c_old, uint8_t k_old, uint8_t k_new) {
    return apch::adjust_threshold(c_old, k_old, k_new);
}
float apch_get_complexity(uint64_t kmer, uint8_t k) { return apch::get_complexity(kmer, k); }
int apch_have_low_complexity(uint64_t kmer, uint8_t k, float threshold) {
    return apch::have_low_complexity(kmer, k, threshold);
}
uint32_t apch_lc_min_filtered_sum(uint8_t k, float threshold) { return apch::lc_min_filtered_sum(k, threshold); }

uint64_t apch_get_most_frequent(uint64_t *kmers, uint64_t *counts, uint64_t n, uint64_t limit, int k) {
    APCH_GUARD(0,
    apch::pair_vector v(n);
    for (uint64_t i = 0; i < n; i++) v[i] = {kmers[i], counts[i]};
    apch::get_most_frequent(v, limit, k);
    for (uint64_t i = 0; i < v.size(); i++) { kmers[i] = v[i].first; counts[i] = v[i].second; }
    return v.size();
    )
}

int apch_export_counter(const uint64_t *kmers, const uint64_t *counts, uint64_t n, uint8_t k, const char *path) {
    APCH_GUARD(0,
    apch::pair_vector v(n);
    for (uint64_t i = 0; i < n; i++) v[i] = {kmers[i], counts[i]};
    return apch::export_counter(v, k, path) ? 1 : 0;
    )
}

int64_t apch_parse_kmer_list(const char *path, uint64_t *out, uint64_t capacity) {
    APCH_GUARD(-1,
    std::vector<uint64_t> v;
    if (!apch::parse_kmer_list(path, v)) return -1;
    for (uint64_t i = 0; i < v.size() && i < capacity; i++) out[i] = v[i];
    return (int64_t)v.size();
    )
}

int apch_reads_load(const char *path, apch_reads **out) {
    APCH_GUARD(-1,
    if (!path || !out) return -1;
    apch_reads *r = new (std::nothrow) apch_reads();
    if (!r) return -1;
    std::string err;
    if (!apch::read_fastx(path, *r, err)) {
        delete r;
        *out = nullptr;
        return -1;
    }
    *out = r;
    return 0;
    )
}
uint64_t apch_reads_count(const apch_reads *r) { return r ? r->size() : 0; }
uint64_t apch_reads_length(const apch_reads *r, uint64_t i) { return r->length(i); }
const char *apch_reads_seq(const apch_reads *r, uint64_t i) { return r->seq(i); }
int apch_reads_mapped(const apch_reads *r) { return r && r->mapping ? 1 : 0; }
void apch_reads_free(apch_reads *r) { delete r; }

int apch_sample(const apch_reads *r, uint64_t nb_sample, uint64_t cut, int bot, int64_t seed, uint8_t *out,
                uint64_t *n_sampled) {
    APCH_GUARD(-1,
    if (!r || !n_sampled) return -1;
    const std::vector<uint64_t> chosen = apch::sample_ids(*r, nb_sample, cut, seed);
    *n_sampled = chosen.size();
    if (out) apch::sample_gather(*r, chosen, cut, bot != 0, out); // straight into the caller's rows
    return 0;
    )
}

int apch_shuffle_order(uint64_t n, int64_t seed, uint32_t *out) {
    APCH_GUARD(-1,
    if (!out && n) return -1;
    const std::vector<int> v = apch::shuffle_order(n, seed);
    for (uint64_t i = 0; i < n; i++) out[i] = (uint32_t)v[i];
    return 0;
    )
}

int apch_synth_ends(uint64_t seed, uint64_t first, uint64_t n, uint32_t sl, int bot, uint8_t *out) {
    APCH_GUARD(-1,
    if (!out && n) return -1;
    apch::synth_ends(seed, first, n, sl, bot != 0, out);
    return 0;
    )
}
int apch_synth_write(const char *path, uint64_t seed, uint64_t n, uint32_t sl, int fastq) {
    APCH_GUARD(-1,
    return path && apch::synth_write(path, seed, n, sl, fastq != 0) ? 0 : -1;
    )
}

int apch_cli_main(int argc, const char **argv) { return apch::cli_main(argc, argv); }

} // extern "C"
