// host_util.h — host-side (CPU) pieces of the drop-in binary: codec, filter
// threshold, CompareCount order, exporters, FASTA/FASTQ reader, sampling,
// synthetic reads.  Citations: /root/reference/approx_counter.cpp:line.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace apch {

using pair_vector = std::vector<std::pair<uint64_t, uint64_t>>; // (k-mer, count) :35-36

// :55-62 / :70-78
bool dna2int(const char *seq, uint32_t k, uint64_t &out);
std::string int2dna(uint64_t value, uint32_t k);

// :183-186
float adjust_threshold(float c_old, uint8_t k_old, uint8_t k_new);
// numerator of the DUST-like score (:216-231): sum over dimer bins of v*(v-1)
uint32_t dimer_sum(uint64_t kmer, uint8_t k);
// :247-267 / :214-234
float get_complexity(uint64_t kmer, uint8_t k);
bool have_low_complexity(uint64_t kmer, uint8_t k, float threshold);
// smallest dimer sum s with float(s)/float(2(k-2)) >= threshold (0xFFFFFFFF if none)
uint32_t lc_min_filtered_sum(uint8_t k, float threshold);

// :275-305
struct CompareCount {
    explicit CompareCount(int k_) : k(k_) {}
    bool operator()(const std::pair<uint64_t, uint64_t> &a, const std::pair<uint64_t, uint64_t> &b) const;
    int k;
};
// :396-405 on an already materialised vector
void get_most_frequent(pair_vector &v, uint64_t limit, int k);

// :157-174
bool export_counter(const pair_vector &v, uint8_t k, const std::string &path);
// :340-364 (returns false if the file cannot be opened)
bool parse_kmer_list(const std::string &path, std::vector<uint64_t> &out);
// :103-135
bool parse_config(const std::string &path, std::vector<std::pair<std::string, std::string>> &out);

// :819-825 — all records of a FASTA/FASTQ file (ids and qualities dropped)
struct Reads {
    // Either the file's own buffer, sequences compacted in place (no second copy of the data) and read i at
    // storage + offsets[i]; or — large files whose records keep their sequence on one line, e.g. 4-line FASTQ — the
    // read-only mapping of the file itself, read i at mapping + starts[i] (no copy at all).
    std::unique_ptr<char[]> storage;
    std::shared_ptr<const char> mapping; // unmapped by its deleter
    uint64_t n_bases = 0;          // sequence letters (ASCII as read)
    std::vector<uint64_t> offsets; // n+1 prefix sums of the read lengths
    std::vector<uint64_t> starts;  // n file offsets (mapping mode only)
    uint64_t size() const { return offsets.empty() ? 0 : offsets.size() - 1; }
    uint64_t length(uint64_t i) const { return offsets[i + 1] - offsets[i]; }
    const char *seq(uint64_t i) const { return mapping ? mapping.get() + starts[i] : storage.get() + offsets[i]; }
};
bool read_fastx(const std::string &path, Reads &out, std::string &err);

// :415-476.  Returns the ASCII sample matrix (n_sampled rows of cut [+1 if bot]).
std::vector<int> shuffle_order(uint64_t n, int64_t seed); // :423-429
// The sample's bytes without the zero fill a std::vector would do first (150 MB at C3: the pages are touched by the
// threads that write them instead).
struct SampleBytes {
    std::unique_ptr<uint8_t[]> bytes;
    size_t n = 0;
    const uint8_t *data() const { return bytes.get(); }
    uint8_t *data() { return bytes.get(); }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
};
std::vector<uint64_t> sample_ids(const Reads &reads, uint64_t nb_sample, uint64_t cut, int64_t seed);               // :447-461
void sample_gather(const Reads &reads, const std::vector<uint64_t> &chosen, uint64_t cut, bool bot, uint8_t *out); // :463 / :466
SampleBytes sample_sequences(const Reads &reads, uint64_t nb_sample, uint64_t cut, bool bot, int64_t seed,
                             uint64_t &n_sampled, uint32_t &row_len);

// synthetic ONT-like reads
std::string synth_read(uint64_t seed, uint64_t index, uint32_t sl);
void synth_ends(uint64_t seed, uint64_t first, uint64_t n, uint32_t sl, bool bot, uint8_t *out);
bool synth_write(const std::string &path, uint64_t seed, uint64_t n, uint32_t sl, bool fastq);

int cli_main(int argc, const char **argv);
extern bool one_shot_process; // true: cli_main leaves the GPU contexts to the end of the process (the binary's main)

} // namespace apch
