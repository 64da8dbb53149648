// apc_reference_shim.h — the reference's own function signatures on top of libapc's C ABI.
//
// A maintainer of qbonenfant/approx_counter who wants the GPU path includes this header in
// approx_counter.cpp and deletes the bodies of count_kmers (:487-519), get_most_frequent
// (:396-405, only where it follows count_kmers) and errorCount (:531-601).  The container
// typedefs are the reference's (:33-36, :44); the only difference is that the sampled reads
// arrive as std::vector<std::string> (convert a SeqAn StringSet<Dna5String> with one loop,
// see INTEGRATION.md) because this header must compile without SeqAn.
//
// C++14, header-only, no CUDA headers needed: it only calls the extern "C" functions of apc.h.
#pragma once

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "apc.h"

namespace apc_shim {

using counter = std::unordered_map<uint64_t, uint64_t>;       // :33
using int_pair = std::pair<uint64_t, uint64_t>;               // :35
using pair_vector = std::vector<int_pair>;                    // :36
using kmer_set_t = std::set<uint64_t>;                        // :44
using sequence_set_type = std::vector<std::string>;           // stands in for StringSet<Dna5String> (:38)

// One process-wide context, created on first use (the reference has no set-up call to hook).
inline apc_ctx *context(int device = 0) {
    static apc_ctx *ctx = nullptr;
    if (!ctx) {
        const int st = apc_create(device, &ctx);
        if (st != APC_OK) throw std::runtime_error(std::string("apc_create: ") + apc_strerror(st));
    }
    return ctx;
}

inline void check(apc_ctx *ctx, int st, const char *what) {
    if (st != APC_OK)
        throw std::runtime_error(std::string(what) + ": " + apc_strerror(st) + " (" + apc_last_error(ctx) + ")");
}

// Hands the sampled reads to the GPU; replaces building the FM index (:537-541).  Call once per
// sample, before count_kmers_topn / errorCount.
inline void upload(const sequence_set_type &sequences) {
    apc_ctx *ctx = context();
    std::vector<uint64_t> offs(sequences.size() + 1, 0);
    for (size_t i = 0; i < sequences.size(); i++) offs[i + 1] = offs[i] + sequences[i].size();
    std::string flat;
    flat.reserve(offs.back());
    for (const auto &s : sequences) flat += s;
    check(ctx, apc_upload_sample_ragged(ctx, reinterpret_cast<const uint8_t *>(flat.data()), offs.data(),
                                        sequences.size()),
          "apc_upload_sample_ragged");
}

// readRecords (:825) on the GPU: the bytes of the FASTA / FASTQ file (any host buffer or mapping) are copied to HBM and
// the records indexed there.  Returns false when the file is one the device parser leaves to SeqAn (wrapped FASTQ,
// blanks inside sequence lines): keep the readRecords route for it.  *n_seqs = length(seqs) (:831).
inline bool readRecordsResident(const void *file_bytes, uint64_t n_bytes, uint64_t *n_seqs) {
    apc_ctx *ctx = context();
    int is_fastq = 0;
    const int st = apc_ingest_fastx(ctx, static_cast<const uint8_t *>(file_bytes), n_bytes, n_seqs, &is_fastq);
    if (st == APC_ERR_FORMAT) return false;
    check(ctx, st, "apc_ingest_fastx");
    return true;
}

// sampleSequences (:415-476) over the records readRecordsResident left on the GPU: the ids are shuffled here exactly
// as in :423-429, the walk (:447-461) and the cuts (:463 / :466) happen on the device, and the sample stays there —
// no upload() afterwards.  Returns the number of sampled reads (`Sampled n sequences`, :870).
template <typename Rng>
inline uint64_t sampleSequencesResident(uint64_t n_seqs, unsigned nb_sample, unsigned cut_size, bool botom, Rng &g) {
    apc_ctx *ctx = context();
    std::vector<uint32_t> vec(n_seqs);
    std::iota(vec.begin(), vec.end(), 0u); // :424-426
    std::shuffle(vec.begin(), vec.end(), g); // :429
    uint64_t n_sampled = 0;
    check(ctx, apc_sample_resident(ctx, vec.data(), vec.size(), nb_sample, cut_size, botom ? 1 : 0, &n_sampled),
          "apc_sample_resident");
    return n_sampled;
}

// count_kmers (:487) followed by get_most_frequent (:396, as called at :898): the `limit` most
// frequent k-mers of the uploaded sample, in CompareCount order.  *n_distinct = count.size() (:883).
inline pair_vector count_kmers_topn(uint8_t k, float threshold, const kmer_set_t &kmer_set, uint64_t limit,
                                    uint64_t *n_distinct = nullptr, uint64_t *had_n = nullptr) {
    apc_ctx *ctx = context();
    std::vector<uint64_t> forb(kmer_set.begin(), kmer_set.end()), km(limit ? limit : 1), ct(limit ? limit : 1);
    uint64_t n = 0, nd = 0, hn = 0;
    check(ctx, apc_exact_topn(ctx, k, threshold, limit, forb.data(), forb.size(), km.data(), ct.data(), &n, &nd, &hn),
          "apc_exact_topn");
    if (n_distinct) *n_distinct = nd;
    if (had_n) *had_n = hn;
    pair_vector out(n);
    for (uint64_t i = 0; i < n; i++) out[i] = {km[i], ct[i]};
    return out;
}

// errorCount (:531): same arguments and return type; `sequences` must be the sample passed to
// upload() (kept in the signature for source compatibility), nb_thread and v are ignored.
inline counter errorCount(const sequence_set_type &sequences, pair_vector &exact_count, uint8_t nb_thread,
                          uint8_t k, uint8_t v) {
    (void)sequences; (void)nb_thread; (void)v;
    apc_ctx *ctx = context();
    std::vector<uint64_t> km(exact_count.size()), approx(exact_count.size());
    for (size_t i = 0; i < km.size(); i++) km[i] = exact_count[i].first; // only .first is read (:584)
    check(ctx, apc_approx_count(ctx, k, km.data(), (uint32_t)km.size(), approx.data()), "apc_approx_count");
    counter results;
    for (size_t i = 0; i < km.size(); i++) results[km[i]] = approx[i]; // :596
    return results;
}

} // namespace apc_shim
