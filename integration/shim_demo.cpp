// shim_demo.cpp — the per-end body of the reference's main loop (:858-952) written against
// apc_reference_shim.h: what approx_counter.cpp looks like after the switch.  Built and run by
// tests/test_gpu_cli.py::test_reference_shim; usage: shim_demo <fasta> <k> <sl> <lim> <lc> <out_prefix> [device]
// With "device" the file is read and sampled on the GPU (readRecordsResident / sampleSequencesResident) instead.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>

#include "apc_host.h"
#include "apc_reference_shim.h"

using namespace apc_shim;

int main(int argc, char **argv) {
    if (argc != 7 && argc != 8) return 2;
    const bool on_device = argc == 8 && std::string(argv[7]) == "device";
    const uint8_t k = (uint8_t)atoi(argv[2]);
    const uint64_t sl = (uint64_t)atoll(argv[3]), limit = (uint64_t)atoll(argv[4]);
    const float lc = apch_adjust_threshold((float)atof(argv[5]), 16, k); // :790
    const std::string out = argv[6];
    apch_reads *reads = nullptr;
    uint64_t n = 0;
    std::mt19937 g(12345);
    if (on_device) {
        std::ifstream f(argv[1], std::ios::binary);
        const std::string bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        if (!readRecordsResident(bytes.data(), bytes.size(), &n)) return 3; // :825 on the GPU
    } else {
        if (apch_reads_load(argv[1], &reads) != 0) return 1;
        n = apch_reads_count(reads);
    }
    const char *ends[2] = {"start", "end"};
    for (int bottom = 0; bottom < 2; bottom++) {
        // sampleSequences (:415-476) with sn >= #reads: every eligible read, ends cut to size
        sequence_set_type sample;
        if (on_device) {
            sampleSequencesResident(n, (unsigned)n, (unsigned)sl, bottom != 0, g); // :867 on the GPU; no upload()
        } else {
            for (uint64_t i = 0; i < n; i++) {
                const uint64_t len = apch_reads_length(reads, i);
                if (len < 2 * sl) continue; // :461
                const char *s = apch_reads_seq(reads, i);
                sample.emplace_back(bottom ? std::string(s + (len - 1 - sl), sl + 1) : std::string(s, sl)); // :463 / :466
            }
            upload(sample);
        }
        kmer_set_t forbidden;
        pair_vector first_n_vector = count_kmers_topn(k, lc, forbidden, limit);              // :874 + :898
        counter error_counter = errorCount(sample, first_n_vector, /*nb_thread*/ 4, k, 1);   // :922
        // get_most_frequent(error_counter, limit, k) (:923) and exportCounter (:928), host side
        std::vector<uint64_t> km, ct;
        for (const auto &p : error_counter) { km.push_back(p.first); ct.push_back(p.second); }
        const uint64_t kept = apch_get_most_frequent(km.data(), ct.data(), km.size(), limit, k);
        if (!apch_export_counter(km.data(), ct.data(), kept, k, (out + "_0." + ends[bottom]).c_str())) return 1;
    }
    if (reads) apch_reads_free(reads);
    apc_destroy(context());
    return 0;
}
