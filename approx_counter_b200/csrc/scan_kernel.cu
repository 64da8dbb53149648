// scan_kernel.cu — K1 approx_scan: the approximate-count hot path for sm_100a.
//
// Replaces errorCount (/root/reference/approx_counter.cpp:531-601): the SeqAn
// FM-index search find<0,2>(…, EditDistance()) (:586), the per-level read
// bitsets (:553-565, :580-582) and their popcount sum (:589-596).  Per
// (k-mer, read) the reference's result is [d<=0]+[d<=1]+[d<=2] with d the
// minimum edit distance between the k-mer and any substring of the read
// (oracle/seqan_model.cpp checks this against the literal
// search-scheme recursion).
//
// Algorithm: bit-parallel Wu–Manber automaton with three error levels.  Bit i
// of level e says "k-mer prefix of length i+1 matches a suffix of the text
// read so far with <= e edits".  Per text column:
//     R0' = ((R0 << 1) | 1) & Eq
//     Re' = ((Re << 1) & Eq) | R(e-1) | (R(e-1) << 1) | (R(e-1)' << 1) | 1
// and a read is flagged at level e when bit k-1 of Re was ever set.  The
// shifted levels are carried between columns (scan_core.cuh), so a column costs
// three shifts, not five.
//
// B200 mapping (integer-pipe bound, no tensor cores, text stays in L2/HBM):
//  * lane = read.  A warp walks one 32-read tile; its 128-bit tile loads are
//    fully coalesced and need no shared-memory staging.
//  * F k-mers are INTERLEAVED in one unit (bit i*F+f = row i of k-mer f), so a
//    shift by F moves every k-mer one row and the F vacated low bits take the
//    "| 1" of all F automata with a single multiply-add: R*2^F + (2^F-1).
//    No carries or cross-field leaks exist, so packing is free:
//      k<=10: 3 per 32-bit word, k<=16: 2 per word, k<=21: 3 per 64-bit pair.
//  * the shifts are issued as IMAD (FMA pipe) by passing 2^F as a runtime
//    operand; the boolean algebra is LOP3 (ALU pipe).  Per unit and column
//    that is 3 IMAD + 5 LOP3 + 1.5 LOP3 of hit accumulation.
//  * Eq comes from a 5-row x 16-byte match table in shared memory, one
//    LDS.128 per column per thread (4 state words); the text byte already is
//    the row offset, so no address arithmetic is spent on it.
//  * each thread keeps per-k-mer hit counters in registers across all tiles
//    of its job; one REDUX + one atomicAdd per (k-mer, warp) at the end.
#include <algorithm>

#include "apc_internal.h"
#include "scan_core.cuh"

namespace apc {


ScanVariant pick_variant(int k, int forced) {
    ScanVariant v{1, 1};
    if (forced == 0) return ScanVariant{0, 1}; // bit-sliced kernel with k-mer pairing: the default
    if (forced == 7) return ScanVariant{0, 0}; // bit-sliced kernel, every k-mer on its own
    switch (forced) {
    case 1: return ScanVariant{1, 1};
    case 2: if (k <= 16) return ScanVariant{1, 2}; break;
    case 3: if (k <= 10) return ScanVariant{1, 3}; break;
    case 6: if (k >= 12 && k <= 21) return ScanVariant{2, 3}; break;
    default: break; // 8 (or an unavailable packing): the best row-packed kernel for this k
    }
    if (k <= 10) v = ScanVariant{1, 3};
    else if (k <= 16) v = ScanVariant{1, 2};
    else if (k <= 21) v = ScanVariant{2, 3};
    return v;
}

// Match tables: table[g][c][w] for group g, base c (0..3; row 4 = N stays 0),
// thread word w.  k-mer q sits in group q / QG, unit (q % QG) / F, field
// (q % QG) % F; row i (i-th base of the k-mer, :70-78 order) is bit i*F+f of
// the unit.
void build_peq_tables(const uint64_t *kmers, uint32_t n_kmers, int k, ScanVariant v, uint32_t *table,
                      uint32_t &n_groups) {
    const uint32_t qg = v.queries_per_group();
    n_groups = (n_kmers + qg - 1) / qg;
    std::fill(table, table + (size_t)n_groups * kPeqRows * kWordsPerThread, 0u);
    // spare rows k .. rows_per_unit-1 compare equal to EVERY text code (N and padding
    // included): they are the delay line of stepT
    const int rows = v.rows_per_unit();
    if (rows > k) {
        uint32_t unit_mask[2] = {0u, 0u};
        for (int i = k; i < rows; i++)
            for (int f = 0; f < v.f; f++) {
                const int bit = i * v.f + f;
                unit_mask[bit / 32] |= 1u << (bit % 32);
            }
        for (uint32_t g = 0; g < n_groups; g++)
            for (int c = 0; c < kPeqRows; c++)
                for (int w = 0; w < kWordsPerThread; w++)
                    table[((size_t)g * kPeqRows + c) * kWordsPerThread + w] = unit_mask[w % v.nw];
    }
    for (uint32_t q = 0; q < n_kmers; q++) {
        const uint32_t g = q / qg, r = q % qg, u = r / v.f, f = r % v.f;
        uint32_t *t = table + (size_t)g * kPeqRows * kWordsPerThread;
        for (int i = 0; i < k; i++) {
            const uint32_t c = (kmers[q] >> (2 * (k - 1 - i))) & 3;
            const uint32_t bit = i * v.f + f;
            t[c * kWordsPerThread + u * v.nw + bit / 32] |= 1u << (bit % 32);
        }
    }
}

__device__ __forceinline__ uint4 ldg_tile(const uint4 *p) { return __ldg(p); }

// Persistent warps: the grid is one wave (SMs x resident CTAs); every WARP pulls jobs
// (tiles_per_job consecutive tiles x one k-mer group) from a global counter, so there is
// no tail wave, no CTA-wide barrier and no per-CTA prologue.  Jobs are numbered group-fastest:
// warps running at the same time read the same text tiles (L2 locality when the text
// outgrows L2).  Each warp owns a 256-byte slot of shared memory for the match table of
// its current group; the slot index is merged into the table offset by the same PRMT that
// extracts the text byte, so a column still costs one PRMT + one LDS.128 per thread.
constexpr int kSlotBytes = 256;
constexpr int kScanBlocksPerSM = 3; // 76-80 registers, 6 warps per SMSP (4 CTAs of 64 registers measured slower)
constexpr uint32_t kMaxTilesPerJob = 64; // 3 hits x 64 tiles < 256: the packed byte counters cannot overflow

template <int F>
__device__ __forceinline__ constexpr uint32_t kFieldMask() {
    return F == 1 ? 0x000001u : F == 2 ? 0x000101u : 0x010101u;
}

template <int NW, int F, int T>
__global__ void __launch_bounds__(kScanWarps * 32, kScanBlocksPerSM)
approx_scan_kernel(const uint4 *__restrict__ tiles, const uint32_t n_tiles, const uint64_t n_reads,
                   const uint32_t chunks, const uint32_t read_len, const uint32_t *__restrict__ peq,
                   const uint32_t n_groups, const uint32_t tiles_per_job, const uint32_t n_jobs,
                   const uint32_t mul, const uint32_t top_shift, const uint32_t hit_rows, const uint32_t n_kmers,
                   unsigned long long *__restrict__ counts, unsigned int *__restrict__ job_counter) {
    constexpr int UNITS = kWordsPerThread / NW;
    constexpr int ACC0 = NW - 1; // word of a unit that holds row k-1
    __shared__ __align__(kSlotBytes) uint32_t s_peq[kScanWarps * kSlotBytes / 4];

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t total_warps = gridDim.x * kScanWarps;
    const uint32_t m = mul - 1;
    const uint32_t full = read_len / kChunkBases, rem = read_len % kChunkBases;
    uint32_t *const slot = s_peq + warp * (kSlotBytes / 4);

    for (;;) {
        uint32_t job = 0;
        if (lane == 0) {
            job = atomicAdd(job_counter, 1u);
            // every warp fails exactly once, so the fetch numbered n_jobs + total_warps - 1 is
            // the last of this launch: it re-arms the counter for the next one
            if (job == n_jobs + total_warps - 1u) atomicExch(job_counter, 0u);
        }
        job = __shfl_sync(0xFFFFFFFFu, job, 0);
        if (job >= n_jobs) break;
        const uint32_t tb = job / n_groups, g = job - tb * n_groups;
        __syncwarp();
        if (lane < kPeqRows * kWordsPerThread) slot[lane] = __ldg(peq + (size_t)g * kPeqRows * kWordsPerThread + lane);
        __syncwarp();
        const uint32_t tile_begin = tb * tiles_per_job;
        const uint32_t tile_end = min(n_tiles, tile_begin + tiles_per_job);

        // per unit: F byte-wide hit counters packed in one register (<= 3 hits per read,
        // <= kMaxTilesPerJob tiles per job, so a byte cannot overflow)
        uint32_t cntw[UNITS];
#pragma unroll
        for (int u = 0; u < UNITS; u++) cntw[u] = 0;

        // Chunks of consecutive tiles are contiguous in memory, so "the next 512 bytes" is
        // always the right prefetch — also across a tile boundary (the buffer is padded by
        // one chunk so the very last prefetch stays in bounds).
        const uint4 *p = tiles + (size_t)tile_begin * chunks * kTileReads + lane;
        uint4 v = ldg_tile(p);
        for (uint32_t tile = tile_begin; tile < tile_end; tile++) {
            ScanState st;
            Column<NW>::init(st, mul, m);
            for (uint32_t ch = 0; ch < full; ch++) {
                p += kTileReads;
                const uint4 nxt = ldg_tile(p);
                const uint32_t tw[4] = {v.x, v.y, v.z, v.w};
                // table offset = slot * 256 + code byte: byte j of tw, byte 0 of `warp`, then zeros (bytes 1 of `warp`)
                if (T == 1) {
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        const uint32_t o0 = __byte_perm(tw[wi], warp, 0x5540u), o1 = __byte_perm(tw[wi], warp, 0x5541u);
                        const uint32_t o2 = __byte_perm(tw[wi], warp, 0x5542u), o3 = __byte_perm(tw[wi], warp, 0x5543u);
                        step2<NW>(st, lds_row(s_peq, o0), lds_row(s_peq, o1), mul, m);
                        step2<NW>(st, lds_row(s_peq, o2), lds_row(s_peq, o3), mul, m);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j += 2 * T) {
                        uint32_t og[2 * T];
#pragma unroll
                        for (int c = 0; c < 2 * T; c++) og[c] = __byte_perm(tw[(j + c) >> 2], warp, 0x5540u + ((j + c) & 3));
                        stepT<NW, T>(st, s_peq, og, mul, m);
                    }
                }
                v = nxt;
            }
            if (rem) {
                // The last, partial chunk also goes through the two-column step: an odd
                // length is rounded up with one padding column.  Padding is coded N (matches
                // nothing), and a text character that matches nothing can never lower the
                // minimum edit distance (aligning to it costs 1, like skipping the k-mer
                // character instead), so the extra column cannot create a hit.
                p += kTileReads;
                const uint4 nxt = ldg_tile(p);
                const uint32_t tw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int wi = 0; wi < 4; wi++) {
                    if ((uint32_t)(wi * 4) < rem) { // warp-uniform
                        const uint32_t o0 = __byte_perm(tw[wi], warp, 0x5540u), o1 = __byte_perm(tw[wi], warp, 0x5541u);
                        step2<NW>(st, lds_row(s_peq, o0), lds_row(s_peq, o1), mul, m);
                    }
                    if ((uint32_t)(wi * 4 + 2) < rem) {
                        const uint32_t o2 = __byte_perm(tw[wi], warp, 0x5542u), o3 = __byte_perm(tw[wi], warp, 0x5543u);
                        step2<NW>(st, lds_row(s_peq, o2), lds_row(s_peq, o3), mul, m);
                    }
                }
                v = nxt;
            }
            // hits of this read: [d<=0] + [d<=1] + [d<=2] per k-mer (:589-593).  The levels are
            // nested (a0 ⊆ a1 ⊆ a2), so the sum is the 2-bit number (a1, a0^a1^a2) per k-mer.
            // Lanes past the last read of a partial tile hold padding and must not count
            // (k <= 2 matches the empty string).
            const bool valid = (uint64_t)tile * kTileReads + lane < n_reads;
#pragma unroll
            for (int u = 0; u < UNITS; u++) {
                const int w = u * NW + ACC0;
                // any bit in rows k-1 .. k-2+hit_rows of an accumulator is a hit (stepT): fold
                // those rows onto row k-1, then keep its F bits
                uint32_t y[3] = {st.a0[w] >> top_shift, st.a1[w] >> top_shift, st.a2[w] >> top_shift};
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    if (T > 1 && hit_rows > 1) y[e] |= y[e] >> F; // warp-uniform branches (T == 1 <=> hit_rows == 1)
                    if (T > 1 && hit_rows > 2) y[e] |= y[e] >> (2 * F);
                    if (T > 1 && hit_rows > 4) y[e] |= y[e] >> (4 * F);
                    if (T > 1 && hit_rows > 8) y[e] |= y[e] >> (8 * F);
                    if (T > 1 && F == 1 && hit_rows > 16) y[e] |= y[e] >> 16;
                    y[e] &= (1u << F) - 1u;
                }
                const uint32_t lo = y[0] ^ y[1] ^ y[2], hi = y[1];
                // spread field f (bit f) to byte f: x * (1 + 2^7 + 2^14) & 0x010101
                const uint32_t slo = (lo * 0x4081u) & kFieldMask<F>(), shi = (hi * 0x4081u) & kFieldMask<F>();
                if (valid) cntw[u] += slo + 2u * shi;
            }
        }

#pragma unroll
        for (int u = 0; u < UNITS; u++) {
#pragma unroll
            for (int f = 0; f < F; f++) {
                const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, (cntw[u] >> (8 * f)) & 0xFFu);
                const uint32_t kslot = g * (UNITS * F) + u * F + f; // == index of the k-mer; the last group may be padded
                if (lane == 0 && total && kslot < n_kmers) atomicAdd(&counts[kslot], (unsigned long long)total);
            }
        }
    }
}

struct ScanRange {
    const uint4 *tiles; // first tile of the range
    uint32_t n_tiles;
    uint64_t n_reads;   // reads in the range (relative to its first tile)
};

template <int NW, int F, int T>
static cudaError_t launch_variant(const Ctx &c, const ScanRange &r, unsigned long long *d_counts,
                                  uint32_t tiles_per_job) {
    const uint64_t jobs = (uint64_t)((r.n_tiles + tiles_per_job - 1) / tiles_per_job) * c.n_groups;
    if (jobs == 0) return cudaSuccess;
    if (jobs > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    const uint32_t mul = 1u << F;
    uint32_t top = (uint32_t)(c.k - 1) * F;
    if (NW == 2) top -= 32; // relative to the high word
    const uint64_t wave = (uint64_t)c.sm_count * kScanBlocksPerSM;
    const unsigned grid = (unsigned)std::min<uint64_t>(wave, (jobs + kScanWarps - 1) / kScanWarps);
    // rows k-1 .. rows_per_unit-1 of an accumulator can hold hits (spare rows are the delay line)
    const uint32_t hit_rows = (uint32_t)(c.variant.rows_per_unit() - c.k + 1);
    approx_scan_kernel<NW, F, T><<<grid, kScanWarps * 32, 0, c.stream>>>(
        r.tiles, r.n_tiles, r.n_reads, c.chunks, c.max_len, c.d_peq, c.n_groups, tiles_per_job, (uint32_t)jobs, mul,
        top, hit_rows, c.n_kmers, d_counts, c.d_job_counter);
    return cudaGetLastError();
}

cudaError_t launch_scan(const Ctx &c, unsigned long long *d_counts, uint64_t *launches) {
    *launches = 0;
    if (c.n_kmers == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_counts, 0, (size_t)c.n_kmers * sizeof(unsigned long long), c.stream);
    if (e != cudaSuccess) return e;
    if (c.n_tiles == 0 || c.max_len == 0) return cudaSuccess;
    // optional sub-range of the resident sample (multi-GPU hosts give every GPU one shard)
    const uint64_t first = c.opt_first_read < c.n_reads ? c.opt_first_read : c.n_reads;
    uint64_t count = c.n_reads - first;
    if (c.opt_n_reads >= 0 && (uint64_t)c.opt_n_reads < count) count = (uint64_t)c.opt_n_reads;
    if (count == 0) return cudaSuccess;
    ScanRange r;
    r.tiles = c.d_tiles + (size_t)(first / kTileReads) * c.chunks * kTileReads;
    r.n_tiles = (uint32_t)((count + kTileReads - 1) / kTileReads);
    r.n_reads = count;
    if (c.variant.bitslice()) {
        // a job is one unit (one or several k-mers) x sg_per_job super-groups (1024 reads each) for one warp;
        // 0 = chosen per launch (bitslice_core.cuh)
        const uint32_t spj = (uint32_t)c.opt_tiles_per_job;
        return launch_bs_scan(c, first, first + count, d_counts, spj, launches);
    }

    uint32_t tpj = (uint32_t)c.opt_tiles_per_job;
    if (tpj == 0) {
        // a job is tiles_per_job tiles x one k-mer group for ONE warp.  Aim for >= 64 jobs per
        // resident warp (tail <= 1/64 of the run) but keep jobs long enough (>= 2 tiles when
        // possible) that the table load and the count flush stay below 1 % of a job.
        const uint64_t warps = (uint64_t)c.sm_count * kScanBlocksPerSM * kScanWarps;
        const uint64_t work = (uint64_t)r.n_tiles * c.n_groups;
        uint64_t t = work / (warps * 64);
        if (t < 1) t = 1;
        if (t > 16) t = 16;
        tpj = (uint32_t)t;
    }
    if (tpj > kMaxTilesPerJob) tpj = kMaxTilesPerJob;
    *launches = 1;
    // accumulator stride: every T-th column is enough when the unit has T-1 spare rows
    const int spare = c.variant.rows_per_unit() - c.k;
    const int t = spare >= 3 ? 4 : spare >= 1 ? 2 : 1;
#define APC_LAUNCH(NW_, F_)                                                            \
    do {                                                                               \
        if (t == 4) return launch_variant<NW_, F_, 4>(c, r, d_counts, tpj);            \
        if (t == 2) return launch_variant<NW_, F_, 2>(c, r, d_counts, tpj);            \
        return launch_variant<NW_, F_, 1>(c, r, d_counts, tpj);                        \
    } while (0)
    if (c.variant.nw == 1 && c.variant.f == 1) APC_LAUNCH(1, 1);
    if (c.variant.nw == 1 && c.variant.f == 2) APC_LAUNCH(1, 2);
    if (c.variant.nw == 1 && c.variant.f == 3) APC_LAUNCH(1, 3);
    if (c.variant.nw == 2 && c.variant.f == 3) APC_LAUNCH(2, 3);
#undef APC_LAUNCH
    return cudaErrorInvalidValue;
}

} // namespace apc
