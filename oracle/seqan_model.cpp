/*
 * seqan_model.cpp — literal restatement of SeqAn 2.4's
 *   find<0, 2>(delegate, index, needle, EditDistance())
 * (include/seqan/index/find2_index_approx.h: OptimalSearchSchemes<0,2>,
 * _optimalSearchScheme / …Children / …Exact / …Deletion) as called by the
 * reference at /root/reference/approx_counter.cpp:586, with the reference's
 * delegate (:556-565) and reduction (:589-596) on top.
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: SeqAn is an un-vendored
 * dependency of the reference (README.md:14, ">= 2.4.0") and is not in this
 * image, so this file is written from the published algorithm (Kianfar et
 * al. optimal search schemes as implemented by SeqAn), not checked against a
 * SeqAn build.  Its only job is to show that the recursion's observable
 * result — which reads get flagged at which error level — equals the closed
 * form the CUDA path and apc_oracle.c implement.
 *
 * The bidirectional FM index is replaced by an explicit occurrence list: an
 * "iterator" is the set of (read, begin, end) text intervals spelling the
 * string matched so far; extending right/left filters and widens the
 * intervals.  Reads are never joined (StringSet sentinels are not edges).
 */
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct Occ { uint32_t read, b, e; };
using Iter = std::vector<Occ>;

struct Search {
    std::array<int, 4> pi, l, u, blocklength;
    int startPos;
};

struct Model {
    const uint8_t *codes;
    const uint64_t *offs;
    uint64_t n_reads;
    int k;
    uint8_t needle[33];               // 0-based needle chars
    std::vector<uint8_t> *tcount;     // [3][n_reads]
    int variant;                      // 0 = first block goes right (SeqAn), 1 = pick by 2nd block

    uint8_t text(const Occ &o, uint32_t pos) const { return codes[offs[o.read] + pos]; }
    uint32_t len(const Occ &o) const { return (uint32_t)(offs[o.read + 1] - offs[o.read]); }

    // children of `it` in direction right (Rev) or left (Fwd): 5 Dna5 edges
    void children(const Iter &it, bool right, std::array<Iter, 5> &out) const {
        for (const Occ &o : it) {
            if (right) {
                if (o.e < len(o)) { uint8_t c = text(o, o.e); out[c > 4 ? 4 : c].push_back({o.read, o.b, o.e + 1}); }
            } else {
                if (o.b > 0) { uint8_t c = text(o, o.b - 1); out[c > 4 ? 4 : c].push_back({o.read, o.b - 1, o.e}); }
            }
        }
    }
    bool goDownChar(Iter &it, uint8_t c, bool right) const {
        Iter nx;
        for (const Occ &o : it) {
            if (right) { if (o.e < len(o) && text(o, o.e) == c) nx.push_back({o.read, o.b, o.e + 1}); }
            else       { if (o.b > 0 && text(o, o.b - 1) == c) nx.push_back({o.read, o.b - 1, o.e}); }
        }
        it.swap(nx);
        return !it.empty();
    }

    void delegate(const Iter &it, int errors) {  // approx_counter.cpp:556-565
        for (const Occ &o : it) tcount[errors][o.read] = 1;
    }

    // needleLeftPos / needleRightPos are SeqAn's 1-based exclusive sentinels.
    void search(const Iter &it, int L, int R, int errors, const Search &s, int b, bool right) {
        const int maxE = s.u[b] - errors;
        const int minE = s.l[b] > errors ? s.l[b] - errors : 0;
        if (minE == 0 && L == 0 && R == k + 1) {
            delegate(it, errors);
        } else if (maxE == 0 && R - L - 1 != s.blocklength[b]) {
            exact(it, L, R, errors, s, b, right);
        } else {
            // insertion: a needle char with no text char
            const int L2 = L - (right ? 0 : 1), R2 = R + (right ? 1 : 0);
            if (R - L == s.blocklength[b]) deletion(it, L2, R2, errors + 1, s, b, right);
            else                           search(it, L2, R2, errors + 1, s, b, right);
            childrenStep(it, L, R, errors, s, b, right);
        }
    }

    void childrenStep(const Iter &it, int L, int R, int errors, const Search &s, int b, bool right) {
        std::array<Iter, 5> ch;
        children(it, right, ch);
        const uint8_t want = needle[right ? R - 1 : L - 1];
        const int L2 = L - (right ? 0 : 1), R2 = R + (right ? 1 : 0);
        for (int c = 0; c < 5; c++) {
            if (ch[c].empty()) continue;
            const int delta = (c != want) ? 1 : 0;   // N never equals a needle char
            // (the Hamming-only lower-bound prune is compiled out for EditDistance)
            if (errors + delta <= s.u[b]) {          // uint8 maxErrorsLeftInBlock can't go negative in SeqAn: delta only taken when it fits
                if (R - L == s.blocklength[b]) deletion(ch[c], L2, R2, errors + delta, s, b, right);
                else                           search(ch[c], L2, R2, errors + delta, s, b, right);
            }
            // deletion: a text char with no needle char
            if (errors + 1 <= s.u[b]) search(ch[c], L, R, errors + 1, s, b, right);
        }
    }

    void exact(Iter it, int L, int R, int errors, const Search &s, int b, bool right) {
        const bool goToRight2 = (b < 3) ? s.pi[b + 1] > s.pi[b] : s.pi[b] > s.pi[b - 1];
        const int b2 = std::min(b + 1, 3);
        if (right) {
            const int infixPosLeft = R - 1;                       // 0-based first char
            const int infixPosRight = L + s.blocklength[b] - 1;   // 0-based last char
            for (int p = infixPosLeft; p <= infixPosRight; p++)
                if (!goDownChar(it, needle[p], true)) return;
            search(it, L, infixPosRight + 2, errors, s, b2, goToRight2);
        } else {
            const int infixPosLeft = R - s.blocklength[b] - 1;
            int infixPosRight = L - 1;
            while (infixPosRight >= infixPosLeft) {
                if (!goDownChar(it, needle[infixPosRight], false)) return;
                --infixPosRight;
            }
            search(it, infixPosLeft, R, errors, s, b2, goToRight2);
        }
    }

    // "_optimalSearchSchemeDeletion": at a block end, optionally swallow more text chars
    void deletion(const Iter &it, int L, int R, int errors, const Search &s, int b, bool right) {
        const int maxE = s.u[b] - errors;
        const int minE = s.l[b] > errors ? s.l[b] - errors : 0;
        if (minE == 0) {
            const int b2 = std::min(b + 1, 3);
            const bool goToRight2 = s.pi[b2] > s.pi[b2 - 1];
            search(it, L, R, errors, s, b2, goToRight2);
        }
        if (maxE > 0) {
            std::array<Iter, 5> ch;
            children(it, right, ch);
            for (int c = 0; c < 5; c++)
                if (!ch[c].empty()) deletion(ch[c], L, R, errors + 1, s, b, right);
        }
    }

    void find(uint64_t kmer) {
        for (int i = 0; i < k; i++) needle[i] = (kmer >> (2 * (k - 1 - i))) & 3;  // int2dna :70-78
        static const int PI[3][4] = {{1, 2, 3, 4}, {3, 2, 1, 4}, {4, 3, 2, 1}};
        static const int LO[3][4] = {{0, 0, 1, 1}, {0, 0, 0, 0}, {0, 0, 0, 2}};
        static const int UP[3][4] = {{0, 0, 2, 2}, {0, 1, 1, 2}, {0, 1, 2, 2}};
        int blocklengths[4];
        const int bl = k / 4, rest = k - 4 * bl;
        for (int i = 0; i < 4; i++) blocklengths[i] = bl + (i < rest ? 1 : 0);
        Iter root;
        for (uint64_t r = 0; r < n_reads; r++) {
            uint32_t n = (uint32_t)(offs[r + 1] - offs[r]);
            for (uint32_t p = 0; p <= n; p++) root.push_back({(uint32_t)r, p, p});
        }
        for (int si = 0; si < 3; si++) {
            Search s;
            for (int i = 0; i < 4; i++) { s.pi[i] = PI[si][i]; s.l[i] = LO[si][i]; s.u[i] = UP[si][i]; }
            for (int i = 0; i < 4; i++)
                s.blocklength[i] = blocklengths[s.pi[i] - 1] + (i > 0 ? s.blocklength[i - 1] : 0);
            s.startPos = 0;
            for (int i = 0; i < 4; i++)
                if (s.pi[i] < s.pi[0]) s.startPos += s.blocklength[i] - (i > 0 ? s.blocklength[i - 1] : 0);
            if (variant == 0 || s.pi[1] > s.pi[0]) {
                search(root, s.startPos, s.startPos + 1, 0, s, 0, true);
            } else {  // perturbation: start at the first block's right edge going left
                const int rightEdge = s.startPos + blocklengths[s.pi[0] - 1];
                search(root, rightEdge, rightEdge + 1, 0, s, 0, false);
            }
        }
    }
};

}  // namespace

extern "C" {

/* counts_out[q] = sum_e popcount(tcount[e]) as in approx_counter.cpp:589-596.
 * flags_out (optional, n_kmers*3*n_reads bytes) receives the raw bitsets. */
void seqan_model_error_count(const uint8_t *codes, const uint64_t *offs, uint64_t n_reads,
                             const uint64_t *kmers, uint64_t n_kmers, uint8_t k, int variant,
                             uint64_t *counts_out, uint8_t *flags_out) {
    for (uint64_t q = 0; q < n_kmers; q++) {
        std::vector<uint8_t> tcount[3];
        for (auto &t : tcount) t.assign(n_reads, 0);   // :580-582
        Model m;
        m.codes = codes; m.offs = offs; m.n_reads = n_reads; m.k = k;
        m.tcount = tcount; m.variant = variant;
        m.find(kmers[q]);
        uint64_t total = 0;
        for (int e = 0; e < 3; e++)
            for (uint64_t r = 0; r < n_reads; r++) total += tcount[e][r];
        counts_out[q] = total;
        if (flags_out)
            for (int e = 0; e < 3; e++)
                std::memcpy(flags_out + (q * 3 + e) * n_reads, tcount[e].data(), n_reads);
    }
}

}  // extern "C"
