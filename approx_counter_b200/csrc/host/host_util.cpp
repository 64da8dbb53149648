// host_util.cpp — see host_util.h.  Citations: /root/reference/approx_counter.cpp:line.
#include "host_util.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <random>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace apch {

static const char DNA[4] = {'A', 'C', 'G', 'T'}; // :22

bool dna2int(const char *seq, uint32_t k, uint64_t &out) { // :55-62
    if (k > 32) return false;
    uint64_t value = 0;
    for (uint32_t i = 0; i < k; i++) {
        uint64_t c;
        switch (seq[i]) {
        case 'A': case 'a': c = 0; break;
        case 'C': case 'c': c = 1; break;
        case 'G': case 'g': c = 2; break;
        case 'T': case 't': c = 3; break;
        default: return false;
        }
        value = value << 2 | c;
    }
    out = value;
    return true;
}

std::string int2dna(uint64_t value, uint32_t k) { // :70-78
    std::string seq(k, 'A');
    for (uint32_t i = 0; i < k; i++) {
        seq[k - 1 - i] = DNA[value & 3];
        value >>= 2;
    }
    return seq;
}

float adjust_threshold(float c_old, uint8_t k_old, uint8_t k_new) { // :183-186
    float c_new = c_old * float(std::pow(k_new - 2 + 1, 2) / std::pow(k_old - 2 + 1, 2));
    return c_new;
}

uint32_t dimer_sum(uint64_t kmer, uint8_t k) { // :216-231
    uint32_t counts[16] = {0};
    for (int i = 0; i < k - 1; i++) {
        counts[kmer & 15]++;
        kmer >>= 2;
    }
    uint32_t sum = 0;
    for (uint32_t v : counts) sum += v * (v - 1);
    return sum;
}

float get_complexity(uint64_t kmer, uint8_t k) { // :247-267
    float s = dimer_sum(kmer, k) / float(2 * (k - 2));
    return s;
}

bool have_low_complexity(uint64_t kmer, uint8_t k, float threshold) { // :214-234
    return get_complexity(kmer, k) >= threshold;
}

uint32_t lc_min_filtered_sum(uint8_t k, float threshold) {
    // dimer sums are <= 31*30 = 930; the quotient is monotone in the numerator
    for (uint32_t s = 0; s <= 1024; s++) {
        float q = s / float(2 * (k - 2));
        if (q >= threshold) return s;
    }
    return 0xFFFFFFFFu;
}

bool CompareCount::operator()(const std::pair<uint64_t, uint64_t> &a,
                              const std::pair<uint64_t, uint64_t> &b) const { // :283-302
    if (a.second == b.second) {
        float a_comp = get_complexity(a.first, (uint8_t)k);
        float b_comp = get_complexity(b.first, (uint8_t)k);
        if (a_comp == b_comp) return a.first > b.first;
        return a_comp < b_comp;
    }
    return a.second > b.second;
}

void get_most_frequent(pair_vector &v, uint64_t limit, int k) { // :396-405
    if (k <= 2) {
        // k == 2 makes the score 0/0 = NaN and the reference comparator stops being a
        // strict weak order; break those ties by k-mer value so the result is defined.
        std::sort(v.begin(), v.end(), [](const auto &a, const auto &b) {
            return a.second != b.second ? a.second > b.second : a.first > b.first;
        });
    } else {
        // same order as std::sort(CompareCount(k)) (:400), with getComplexity evaluated once
        // per entry instead of twice per comparison
        struct Item {
            uint64_t kmer, count;
            float comp;
        };
        std::vector<Item> items(v.size());
        for (size_t i = 0; i < v.size(); i++) items[i] = {v[i].first, v[i].second, get_complexity(v[i].first, (uint8_t)k)};
        std::sort(items.begin(), items.end(), [](const Item &a, const Item &b) {
            if (a.count == b.count) {
                if (a.comp == b.comp) return a.kmer > b.kmer;
                return a.comp < b.comp;
            }
            return a.count > b.count;
        });
        for (size_t i = 0; i < v.size(); i++) v[i] = {items[i].kmer, items[i].count};
    }
    if (v.size() > limit) v.resize(limit);
}

bool export_counter(const pair_vector &v, uint8_t k, const std::string &path) { // :157-174
    std::ofstream f(path);
    if (!f.is_open()) {
        fprintf(stderr, "/!\\ ERROR: COULD NOT OPEN FILE %s\n", path.c_str());
        return false;
    }
    std::string buf;
    buf.reserve(v.size() * (k + 12));
    for (const auto &p : v) {
        buf += int2dna(p.first, k);
        buf += '\t';
        buf += std::to_string(p.second);
        buf += '\n';
    }
    f.write(buf.data(), (std::streamsize)buf.size());
    f.close();
    return true;
}

bool parse_kmer_list(const std::string &path, std::vector<uint64_t> &out) { // :340-364
    std::ifstream f(path);
    if (!f.is_open()) return false;
    for (std::string line; std::getline(f, line);) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        if (line.empty() || line.size() > 32) continue;
        uint64_t v;
        if (dna2int(line.c_str(), (uint32_t)line.size(), v)) out.push_back(v); // only true DNA k-mers :353
    }
    return true;
}

bool parse_config(const std::string &path, std::vector<std::pair<std::string, std::string>> &out) { // :103-135
    std::ifstream f(path);
    if (!f.is_open()) {
        fprintf(stderr, "/!\\ WARNING: Could not open config file\n");
        return false;
    }
    for (std::string line; std::getline(f, line);) {
        std::string arg, val;
        bool sep = false;
        if (!line.empty() && line[0] == '#') continue;
        for (char c : line) {
            if (c == '=') sep = true;
            else if (c != ' ' && c != '\r') (sep ? val : arg) += c;
        }
        out.emplace_back(arg, val); // the reference also records empty lines as ""="" (:127)
    }
    return true;
}

// ---- FASTA / FASTQ ---------------------------------------------------------------
// Replaces SeqFileIn + readRecords (:819-825).  The whole file is read into one buffer,
// cut into one piece per host thread at record boundaries, and every piece compacts its
// sequence letters IN PLACE towards the start of the piece with memchr/memmove (no
// per-character work on the fast path, no second copy of the data); the gaps between the
// pieces are then closed front to back.  Multi-line records, CRLF and blank lines are
// accepted; ids and qualities are dropped.
namespace {

struct Piece {
    char *begin = nullptr;  // piece region; compacted letters end up at [begin, begin + n_bases)
    const char *end = nullptr;
    uint64_t n_bases = 0;
    std::vector<uint64_t> lens;
    bool ok = true;
    std::string err;
};

inline const char *line_end(const char *p, const char *end) {
    const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
    return nl ? nl : end;
}

// move the letters of the line [p, e) (without its '\n') to w; returns how many were taken.
// w <= p always holds (compaction only moves bytes towards the front).
inline uint64_t take_line(char *w, const char *p, const char *e) {
    while (e > p && (e[-1] == '\r' || e[-1] == ' ' || e[-1] == '\t')) e--;
    const size_t n = (size_t)(e - p);
    if (n == 0) return 0;
    if (!memchr(p, ' ', n) && !memchr(p, '\t', n) && !memchr(p, '\r', n)) {
        if (w != p) memmove(w, p, n);
        return n;
    }
    uint64_t taken = 0;
    for (; p < e; p++)
        if (*p != ' ' && *p != '\t' && *p != '\r') w[taken++] = *p;
    return taken;
}

inline const char *skip_blank(const char *p, const char *end) {
    while (p < end && (*p == '\n' || *p == '\r' || *p == ' ' || *p == '\t')) p++;
    return p;
}

// parse whole records from the piece; `fastq` selects the grammar
void parse_piece(Piece &pc, bool fastq) {
    const char *p = pc.begin, *end = pc.end;
    char *w = pc.begin;
    while (true) {
        p = skip_blank(p, end);
        if (p >= end) break;
        if (*p != (fastq ? '@' : '>')) {
            pc.ok = false;
            pc.err = fastq ? "malformed FASTQ record" : "malformed FASTA record";
            return;
        }
        p = line_end(p, end); // header line
        if (p < end) p++;
        uint64_t n = 0;
        const char stop = fastq ? '+' : '>';
        while (p < end && *p != stop) {
            const char *e = line_end(p, end);
            const uint64_t got = take_line(w, p, e);
            w += got;
            n += got;
            p = e < end ? e + 1 : end;
        }
        if (fastq) {
            if (p >= end) { pc.ok = false; pc.err = "truncated FASTQ record"; return; }
            p = line_end(p, end); // '+' line
            if (p < end) p++;
            uint64_t q = 0; // the quality string may contain '@' and '>': count characters
            while (p < end && q < n) {
                const char *e = line_end(p, end);
                const char *t = e;
                while (t > p && t[-1] == '\r') t--;
                q += (uint64_t)(t - p);
                p = e < end ? e + 1 : end;
            }
        }
        pc.lens.push_back(n);
    }
    pc.n_bases = (uint64_t)(w - pc.begin);
}

// Read-only check that [begin, end) is a whole number of FASTQ records: every record starts with '@', has a
// '+' line after its sequence lines and a quality string of exactly the sequence's length.  The boundary
// heuristic of next_record can be defeated by wrapped (multi-line) records whose quality lines start with '@'
// and '+'; a piece cut inside a record fails this check on either side of the cut, and the file is then parsed
// in one piece (the in-place compaction of parse_piece cannot be undone, so the check runs first).
bool verify_fastq_piece(const char *p, const char *end) {
    while (true) {
        p = skip_blank(p, end);
        if (p >= end) return true;
        if (*p != '@') return false;
        p = line_end(p, end);
        if (p < end) p++;
        uint64_t n = 0;
        while (p < end && *p != '+') {
            const char *e = line_end(p, end), *t = e;
            while (t > p && (t[-1] == '\r' || t[-1] == ' ' || t[-1] == '\t')) t--;
            for (const char *c = p; c < t; c++) n += (*c != ' ' && *c != '\t' && *c != '\r');
            p = e < end ? e + 1 : end;
        }
        if (p >= end) return false;
        p = line_end(p, end);
        if (p < end) p++;
        uint64_t q = 0;
        while (p < end && q < n) {
            const char *e = line_end(p, end), *t = e;
            while (t > p && t[-1] == '\r') t--;
            q += (uint64_t)(t - p);
            p = e < end ? e + 1 : end;
        }
        if (q != n) return false;
    }
}

// first record start at or after `p` (a position inside the file)
const char *next_record(const char *begin, const char *p, const char *end, bool fastq) {
    if (p <= begin) return begin;
    p = line_end(p - 1, end); // move to a line start
    if (p < end) p++;
    while (p < end) {
        if (!fastq) {
            if (*p == '>') return p;
        } else if (*p == '@') {
            // a header is a line starting with '@' whose second-next line starts with '+'
            // (4-line FASTQ; a quality line starting with '@' is followed by a header, whose
            // second-next line is a sequence and cannot start with '+')
            const char *l1 = line_end(p, end);
            if (l1 < end) {
                const char *l2 = line_end(l1 + 1, end);
                if (l2 < end && l2 + 1 < end && l2[1] == '+') return p;
            }
        }
        p = line_end(p, end);
        if (p < end) p++;
    }
    return end;
}

// ---- zero-copy path: records whose sequence sits on ONE line, parsed straight from a read-only mapping ----
struct MapPiece {
    const char *begin = nullptr, *end = nullptr;
    std::vector<uint64_t> starts, lens; // file offsets and lengths of the sequences
    bool ok = true;      // false: malformed (the copying parser will report it)
    bool simple = true;  // false: a multi-line record or blanks inside a sequence line -> use the copying parser
};

void parse_piece_mapped(MapPiece &pc, const char *file_begin, bool fastq) {
    const char *p = pc.begin, *end = pc.end;
    const char stop = fastq ? '+' : '>';
    while (true) {
        p = skip_blank(p, end);
        if (p >= end) break;
        if (*p != (fastq ? '@' : '>')) { pc.ok = false; return; }
        p = line_end(p, end); // header line
        if (p < end) p++;
        const char *s = p, *e = p;
        if (p < end && *p != stop) { // the sequence line
            e = line_end(p, end);
            p = e < end ? e + 1 : end;
            while (e > s && (e[-1] == '\r' || e[-1] == ' ' || e[-1] == '\t')) e--;
            if (memchr(s, ' ', (size_t)(e - s)) || memchr(s, '\t', (size_t)(e - s)) || memchr(s, '\r', (size_t)(e - s))) {
                pc.simple = false;
                return;
            }
            // anything but the stop character (or, FASTA, the end / blank lines) next means a second sequence line
            const char *q = fastq ? p : skip_blank(p, end);
            if (q < end && *q != stop) { pc.simple = false; return; }
            if (fastq && q >= end) { pc.ok = false; return; }
        } else if (fastq && p >= end) {
            pc.ok = false;
            return;
        }
        const uint64_t n = (uint64_t)(e - s);
        if (fastq) {
            p = line_end(p, end); // '+' line
            if (p < end) p++;
            uint64_t q = 0; // the quality string may contain '@' and '>': count characters
            while (p < end && q < n) {
                const char *le = line_end(p, end);
                const char *t = le;
                while (t > p && t[-1] == '\r') t--;
                q += (uint64_t)(t - p);
                p = le < end ? le + 1 : end;
            }
            // a quality string that is not exactly as long as the sequence means that this piece does not hold
            // whole records (a cut inside a wrapped record): the copying parser re-checks and parses in one piece
            if (q != n) { pc.ok = false; return; }
        }
        pc.starts.push_back((uint64_t)(s - file_begin));
        pc.lens.push_back(n);
    }
}

// Returns true when `out` was filled from a mapping of the file; false = use the copying parser (small file, pipe,
// multi-line records, anything unusual or malformed).
bool read_fastx_mapped(const std::string &path, Reads &out) {
    size_t min_bytes = (size_t)1 << 25;
    if (const char *env = getenv("APCH_MMAP_BYTES")) min_bytes = strtoull(env, nullptr, 10); // tests; 0 disables
    if (min_bytes == 0) return false;
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size < min_bytes) {
        close(fd);
        return false;
    }
    const size_t size = (size_t)st.st_size;
    void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return false;
    madvise(m, size, MADV_SEQUENTIAL);
    std::shared_ptr<const char> mapping((const char *)m, [size](const char *q) { munmap((void *)q, size); });
    const char *begin = mapping.get(), *end = begin + size;
    const char *p0 = skip_blank(begin, end);
    if (p0 >= end || (*p0 != '>' && *p0 != '@')) return false;
    const bool fastq = *p0 == '@';
    int n_pieces = 1;
#ifdef _OPENMP
    n_pieces = omp_get_max_threads();
#endif
    size_t piece_bytes = (size_t)1 << 25;
    if (const char *env = getenv("APCH_PIECE_BYTES")) piece_bytes = std::max<size_t>(1, strtoull(env, nullptr, 10)); // tests
    n_pieces = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_pieces * 4, size / piece_bytes));
    std::vector<MapPiece> pieces((size_t)n_pieces);
    const char *prev = p0;
    for (int i = 0; i < n_pieces; i++) {
        const char *next = end;
        if (i + 1 < n_pieces) next = std::max(prev, next_record(begin, begin + size / (size_t)n_pieces * (size_t)(i + 1), end, fastq));
        pieces[(size_t)i].begin = prev;
        pieces[(size_t)i].end = next;
        prev = next;
    }
    for (const MapPiece &pc : pieces)
        if (pc.begin < pc.end && *pc.begin != (fastq ? '@' : '>')) return false; // boundary heuristic defeated
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n_pieces; i++) parse_piece_mapped(pieces[(size_t)i], begin, fastq);
    size_t total = 0;
    for (const MapPiece &pc : pieces) {
        if (!pc.ok || !pc.simple) return false;
        total += pc.lens.size();
    }
    out.storage.reset();
    out.starts.resize(total);
    out.offsets.resize(total + 1);
    uint64_t off = 0;
    size_t r = 0;
    for (const MapPiece &pc : pieces)
        for (size_t j = 0; j < pc.lens.size(); j++) {
            out.starts[r] = pc.starts[j];
            out.offsets[r++] = off;
            off += pc.lens[j];
        }
    out.offsets[total] = off;
    out.n_bases = off;
    out.mapping = std::move(mapping);
    return true;
}

} // namespace

bool read_fastx(const std::string &path, Reads &out, std::string &err) {
    out.mapping.reset();
    out.starts.clear();
    if (read_fastx_mapped(path, out)) return true;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        err = "could not open " + path;
        return false;
    }
    // one read of the whole file into an uninitialised buffer (grown if the input is a pipe)
    size_t cap = 1 << 16, size = 0;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long fsize = ftell(f);
        if (fsize > 0) cap = (size_t)fsize + 1;
        fseek(f, 0, SEEK_SET);
    }
    std::unique_ptr<char[]> buf(new char[cap]);
    for (;;) {
        if (size == cap) {
            std::unique_ptr<char[]> bigger(new char[cap * 2]);
            memcpy(bigger.get(), buf.get(), size);
            buf.swap(bigger);
            cap *= 2;
        }
        const size_t n = fread(buf.get() + size, 1, cap - size, f);
        if (n == 0) break;
        size += n;
    }
    fclose(f);
    out.storage.reset();
    out.n_bases = 0;
    out.offsets.assign(1, 0);
    char *begin = buf.get(), *end = begin + size;
    char *p0 = const_cast<char *>(skip_blank(begin, end));
    if (p0 >= end) return true; // empty file: no records
    if (*p0 != '>' && *p0 != '@') {
        err = "unrecognised sequence file format (expected FASTA '>' or FASTQ '@')";
        return false;
    }
    const bool fastq = *p0 == '@';

    int n_pieces = 1;
#ifdef _OPENMP
    n_pieces = omp_get_max_threads();
#endif
    // parsing is memory-bound: one piece per 32 MB is enough, more threads only add wake-up cost
    size_t piece_bytes = (size_t)1 << 25;
    if (const char *env = getenv("APCH_PIECE_BYTES")) piece_bytes = std::max<size_t>(1, strtoull(env, nullptr, 10)); // tests
    n_pieces = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_pieces, size / piece_bytes));
    auto cut_pieces = [&](int n, std::vector<Piece> &pieces) {
        pieces.assign((size_t)n, Piece());
        const char *prev = p0;
        for (int i = 0; i < n; i++) {
            const char *next = end;
            if (i + 1 < n) next = std::max(prev, next_record(begin, begin + size / (size_t)n * (size_t)(i + 1), end, fastq));
            pieces[(size_t)i].begin = const_cast<char *>(prev);
            pieces[(size_t)i].end = next;
            prev = next;
        }
    };
    std::vector<Piece> pieces;
    if (n_pieces > 1) {
        // dry run on record boundaries only: a piece must start with a record marker and the
        // in-place compaction below cannot be undone, so unusual files (multi-line FASTQ whose
        // quality lines defeat the boundary heuristic) are detected first by a cheap check
        cut_pieces(n_pieces, pieces);
        bool whole = true;
        for (const Piece &pc : pieces)
            if (pc.begin < pc.end && *pc.begin != (fastq ? '@' : '>')) whole = false;
        if (whole && fastq) {
            const int n_chk = n_pieces;
#pragma omp parallel for schedule(static, 1) reduction(&& : whole)
            for (int i = 0; i < n_chk; i++) whole = whole && verify_fastq_piece(pieces[(size_t)i].begin, pieces[(size_t)i].end);
        }
        if (!whole) n_pieces = 1;
    }
    cut_pieces(n_pieces, pieces);
#pragma omp parallel for schedule(static, 1)
    for (int i = 0; i < n_pieces; i++) parse_piece(pieces[(size_t)i], fastq);
    for (const Piece &pc : pieces)
        if (!pc.ok) {
            err = pc.err;
            return false;
        }
    // close the gaps between pieces, front to back (destinations never overtake sources)
    size_t total_reads = 0;
    char *w = begin;
    for (Piece &pc : pieces) {
        if (pc.n_bases && w != pc.begin) memmove(w, pc.begin, pc.n_bases);
        w += pc.n_bases;
        total_reads += pc.lens.size();
    }
    out.n_bases = (uint64_t)(w - begin);
    out.offsets.resize(total_reads + 1);
    uint64_t off = 0;
    size_t r = 0;
    for (const Piece &pc : pieces)
        for (uint64_t len : pc.lens) {
            out.offsets[r++] = off;
            off += len;
        }
    out.offsets[total_reads] = off;
    out.storage = std::move(buf);
    return true;
}

// memcpy split over the host threads: one thread moves about 10 GB/s out of a file mapping (page-cache pages are
// mapped on first touch), the copy engine takes five times that from page-locked memory (apc_ingest_fastx)
void parallel_copy(void *dst, const void *src, size_t n) {
    const size_t piece = (size_t)1 << 20;
    const int64_t pieces = (int64_t)((n + piece - 1) / piece);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < pieces; i++) {
        const size_t off = (size_t)i * piece;
        memcpy((char *)dst + off, (const char *)src + off, std::min(piece, n - off));
    }
}

// ---- sampling --------------------------------------------------------------------
std::vector<int> shuffle_order(uint64_t n, int64_t seed) { // :423-429
    std::vector<int> vec(n);
    std::iota(vec.begin(), vec.end(), 0);
    std::mt19937 g;
    if (seed < 0) {
        std::random_device rd; // :427-428
        g.seed(rd());
    } else {
        g.seed((uint32_t)seed);
    }
    std::shuffle(vec.begin(), vec.end(), g); // :429
    return vec;
}

// :447-461 — the ids the walk over the shuffled order takes: the first nb_sample reads of at least 2 * cut bases
std::vector<uint64_t> sample_ids(const Reads &reads, uint64_t nb_sample, uint64_t cut, int64_t seed) {
    const uint64_t n = reads.size();
    const std::vector<int> vec = shuffle_order(n, seed);
    std::vector<uint64_t> chosen;
    chosen.reserve((size_t)std::min<uint64_t>(nb_sample, n));
    for (uint64_t i = 0; chosen.size() < nb_sample && i < n; i++) { // :447
        const uint64_t id = (uint64_t)vec[i];
        if (cut > 0 && reads.length(id) >= cut * 2) chosen.push_back(id); // :461 (current_cut_size == cut_size here)
    }
    return chosen;
}

// :463 / :466 — row r of `out` = suffix(seq, len-1-cut) (cut+1 bases) or prefix(seq, cut) of read chosen[r]; the copies
// only need the walk's result, so they run in parallel
void sample_gather(const Reads &reads, const std::vector<uint64_t> &chosen, uint64_t cut, bool bot, uint8_t *out) {
    const uint32_t row_len = (uint32_t)(cut + (bot ? 1 : 0));
    const int64_t n_sampled = (int64_t)chosen.size();
    auto source = [&](int64_t r) {
        const uint64_t id = chosen[(size_t)r];
        return reads.seq(id) + (bot ? reads.length(id) - 1 - cut : 0);
    };
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_sampled; r++) {
        // the reads sit at random places of a file-sized buffer: every row is a TLB and cache miss, so the rows a
        // few iterations ahead are requested now
        if (r + 8 < n_sampled) {
            const char *ahead = source(r + 8);
            for (uint32_t b = 0; b < row_len; b += 64) __builtin_prefetch(ahead + b, 0, 0);
        }
        memcpy(out + (size_t)r * row_len, source(r), row_len);
    }
}

SampleBytes sample_sequences(const Reads &reads, uint64_t nb_sample, uint64_t cut, bool bot, int64_t seed,
                             uint64_t &n_sampled, uint32_t &row_len) { // :415-476
    row_len = (uint32_t)(cut + (bot ? 1 : 0));
    const std::vector<uint64_t> chosen = sample_ids(reads, nb_sample, cut, seed);
    n_sampled = chosen.size();
    SampleBytes sample;
    sample.n = (size_t)n_sampled * row_len;
    sample.bytes.reset(new uint8_t[sample.n ? sample.n : 1]);
    sample_gather(reads, chosen, cut, bot, sample.data());
    return sample;
}

// ---- synthetic reads -------------------------------------------------
namespace {
struct Rng { // splitmix64-seeded xoshiro256**, raw outputs only
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t index) {
        uint64_t x = seed * 0xd1342543de82ef95ULL + index * 0x9e3779b97f4a7c15ULL + 0x2545f4914f6cdd1dULL;
        for (auto &v : s) v = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
};
const char ADAPTER_START[] = "AATGTACTTCGTTCAGTTACGTATTGCT";
const char ADAPTER_END[] = "GCAATACGTAACTGAACGAAGT";

std::string error_channel(Rng &g, const char *adapter) { // sub 3 %, ins 2 %, del 3 % per base
    std::string m;
    for (const char *p = adapter; *p; p++) {
        const uint64_t r = g.next() % 100;
        if (r < 3) {
            int b = (int)(strchr("ACGT", *p) - "ACGT");
            m += "ACGT"[(b + 1 + g.next() % 3) & 3];
        } else if (r < 5) {
            m += "ACGT"[g.next() & 3];
            m += *p;
        } else if (r < 8) {
            // deleted
        } else {
            m += *p;
        }
    }
    return m;
}
} // namespace

std::string synth_read(uint64_t seed, uint64_t index, uint32_t sl) {
    Rng g(seed, index);
    const uint64_t len = 2ull * sl + 50 + g.next() % 400;
    std::string s(len, 'A');
    for (uint64_t i = 0; i < len; i++) {
        const uint64_t r = g.next();
        s[i] = (r % 10000 == 0) ? 'N' : "ACGT"[(r >> 20) & 3];
    }
    // bit 40 of the seed selects the "wide" variant of a stream: adapter offsets uniform in 0..sl/2 instead of
    // 0..7 (bench.py reports both: how much of the throughput comes from the adapters sitting in the same columns)
    const uint64_t off_range = (seed >> 40) & 1 ? (uint64_t)sl / 2 + 1 : 8;
    if (g.next() % 10 != 0) { // 90 % of reads carry both adapters
        const uint64_t off_s = g.next() % off_range;
        const std::string a = error_channel(g, ADAPTER_START);
        if (off_s + a.size() <= len) s.replace(off_s, a.size(), a);
        const uint64_t off_e = g.next() % off_range;
        const std::string b = error_channel(g, ADAPTER_END);
        if (off_e + b.size() <= len) s.replace(len - off_e - b.size(), b.size(), b);
    }
    return s;
}

void synth_ends(uint64_t seed, uint64_t first, uint64_t n, uint32_t sl, bool bot, uint8_t *out) {
    const uint32_t row = sl + (bot ? 1 : 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const std::string s = synth_read(seed, first + (uint64_t)i, sl);
        const char *src = bot ? s.data() + (s.size() - 1 - sl) : s.data();
        memcpy(out + (size_t)i * row, src, row);
    }
}

bool synth_write(const std::string &path, uint64_t seed, uint64_t n, uint32_t sl, bool fastq) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    std::string buf;
    for (uint64_t i = 0; i < n; i++) {
        const std::string s = synth_read(seed, i, sl);
        buf += fastq ? '@' : '>';
        buf += 'r';
        buf += std::to_string(i);
        buf += '\n';
        buf += s;
        buf += '\n';
        if (fastq) {
            buf += "+\n";
            buf.append(s.size(), 'I');
            buf += '\n';
        }
        if (buf.size() > (1u << 22)) {
            fwrite(buf.data(), 1, buf.size(), f);
            buf.clear();
        }
    }
    fwrite(buf.data(), 1, buf.size(), f);
    fclose(f);
    return true;
}

} // namespace apch
