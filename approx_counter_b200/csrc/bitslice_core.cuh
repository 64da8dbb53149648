// bitslice_core.cuh — the bit-sliced scan kernels (K1) and their per-k launcher.  Included by
// bitslice_part.cu, which is compiled once per slice of k so that the ~250 instantiations
// build in parallel.  See bitslice_kernel.cu for the design notes.
#pragma once

#include <algorithm>

#include "apc_internal.h"
#include "scan_core.cuh"

namespace apc {

constexpr int kGroupsPerSuper = 32; // lanes
constexpr int kPlaneRow = 128;      // bytes between the A, C, G, T rows of the per-warp mask slot

#ifndef APC_BS_12W_ROWS
#define APC_BS_12W_ROWS 24
#endif
constexpr int bs_warps_per_sm_c(int rows) {
#ifdef APC_BS_MB_OVERRIDE
    (void)rows;
    return APC_BS_MB_OVERRIDE; // A/B builds (tools/build_ab.sh)
#endif
    // registers are handed out per SM sub-partition (16384 each), so only multiples of 4 warps matter:
    // 24 -> 80 registers, 20 -> 96, 16 -> 128, 12 -> 168, 8 -> 255
    return rows <= 8 ? 24 : rows <= 12 ? 20 : rows <= 16 ? 16 : rows <= APC_BS_12W_ROWS ? 12 : 8;
}

// Values handed from row i-1 to row i inside one column.
struct BsCarry {
    uint32_t p0, p1, p2; // previous column's row i-1, levels 0..2 (row -1 = empty prefix: always matches)
    uint32_t n0p, n1p;   // this column's row i-1, levels 0 and 1
};

__device__ __forceinline__ BsCarry bs_carry_init() {
    const uint32_t ALL = 0xFFFFFFFFu;
    return BsCarry{ALL, ALL, ALL, ALL, ALL};
}

// N consecutive rows of one text column, the first of them being row FIRST of the k-mer: e_i
// from the warp's mask slot (LDS with a uniform-register offset), then the five LOP3 of the row.
// HIT: the last of the N rows is row k-1, the row whose bits are the hits.  It is made sticky —
// each level keeps every bit it ever had — so that at the end of a read it holds "level e was
// reached in SOME column" and no separate accumulators are needed.  Because the levels are nested
// (R0 <= R1 <= R2, also with the sticky bits), OR-ing in the old value of the same level subsumes
// the insertion term (the old value of the level below): the sticky row costs the same five LOP3.
template <int N, int FIRST, bool HIT, int J0 = 0, int J1 = N>
__device__ __forceinline__ void bs_rows(uint32_t (&r0)[N], uint32_t (&r1)[N], uint32_t (&r2)[N], BsCarry &c,
                                        const char *slot_lane, const uint32_t (&off)[N]) {
    const uint32_t ALL = 0xFFFFFFFFu;
#pragma unroll
    for (int j = J0; j < J1; j++) { // rows J0..J1-1 of the N held in these arrays
        const int i = FIRST + j;
        const uint32_t e = *reinterpret_cast<const uint32_t *>(slot_lane + off[j]);
        const uint32_t o0 = r0[j], o1 = r1[j], o2 = r2[j];
        uint32_t n0, n1, n2;
        if (HIT && j == N - 1) {
            n0 = i < 1 ? (e | o0) : and_or(c.p0, e, o0);
            n1 = i < 1 ? ALL : or3(and_or(c.p1, e, o1), c.p0, c.n0p);
            n2 = i < 2 ? ALL : or3(and_or(c.p2, e, o2), c.p1, c.n1p);
        } else {
            n0 = i < 1 ? e : and2(c.p0, e);
            // rows 0 (level 1) and 0..1 (level 2) always match: that many k-mer bases can be skipped
            n1 = i < 1 ? ALL : or3(and_or(c.p1, e, o0), c.p0, c.n0p);
            n2 = i < 2 ? ALL : or3(and_or(c.p2, e, o1), c.p1, c.n1p);
        }
        r0[j] = n0; r1[j] = n1; r2[j] = n2;
        c.p0 = o0; c.p1 = o1; c.p2 = o2;
        c.n0p = n0; c.n1p = n1;
    }
}

// ---- dead-row skipping ------------------------------------------------------------------------------
// Row i is non-zero in a column only where the k-mer's first i+1 bases match the text with <= 2 edits.  For
// i >= 12 that is rare in unrelated sequence (about 4 % of the columns for the 1024 reads of a warp) — the
// query k-mers are an adapter's windows, and the adapter sits in a few dozen columns of the sampled read
// end.  So the rows are split at row M: rows < M (the TOP) are computed in every column; rows >= M (the
// DEEP part: the end of the trunk and the tails, or the ends of the tails) only while something can reach
// them.  Deep rows cannot change in a column when (a) level 2 of row M-1 is zero before and after the
// column for every read of the warp (nested levels: then levels 0 and 1 are zero too) and (b) every deep
// row except the sticky hit rows is zero: each term of their update is then zero, and a hit row only
// keeps what it has.  (a) is one OR + one warp vote per column; (b) is re-established with an OR over the
// deep rows + a vote only on the first quiet column after a busy stretch.
constexpr int bs_check_row(int k) { return bs_check_row_host(k); } // apc_internal.h (the planner needs it too)

// Rows of a unit (trunk of P rows + G tails of K - P rows; P = K, G = 1: a single k-mer) that are computed in every
// column (top) and only where something can reach them (deep), and the columns one dead-row test covers.
struct BsRowSplit {
    int top, deep, cols_per_test;
};
__host__ __device__ constexpr BsRowSplit bs_rows_split(int k, int p, int g, int m) {
    const int t = k - p, rows = p + g * t;
    if (m >= k) return BsRowSplit{rows, 0, 2};
    if (m <= p) return BsRowSplit{m, rows - m, 2};          // one test per column pair
    return BsRowSplit{p + g * (m - p), g * (k - m), 1};     // split inside the tails: one test per column
}

template <int N, int FIRST>
__device__ __forceinline__ void bs_rows_init(uint32_t (&r0)[N], uint32_t (&r1)[N], uint32_t (&r2)[N]) {
    const uint32_t ALL = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < N; j++) {
        r0[j] = 0;
        r1[j] = FIRST + j < 1 ? ALL : 0; // prefix 1 by one deletion
        r2[j] = FIRST + j < 2 ? ALL : 0; // prefixes 1..2 by deletions
    }
}

// mask of the reads of group (sg, lane) that lie inside the scanned range
__device__ __forceinline__ uint32_t bs_valid_mask(uint64_t first, uint64_t range_lo, uint64_t range_hi) {
    const uint32_t ALL = 0xFFFFFFFFu;
    uint32_t vm = 0;
    if (first < range_hi && first + 32 > range_lo) {
        vm = ALL;
        if (range_lo > first) vm &= ALL << (uint32_t)(range_lo - first);
        if (range_hi < first + 32) vm &= ALL >> (uint32_t)(first + 32 - range_hi);
    }
    return vm;
}

// job fetch of the persistent warps: warp-uniform result (ptxas keeps it in uniform registers).  Claiming one job
// ahead (the atomic's round trip hidden behind the job) was measured and is not worth its lines: C2 372.9 vs 375.3,
// C3 674.7 vs 671.0, C4 651.5 vs 654.6 kGCUPS with / without (tools/ab_bench.sh).
__device__ __forceinline__ uint32_t bs_next_job(unsigned int *job_counter, uint32_t n_jobs, uint32_t lane) {
    uint32_t job = 0;
    if (lane == 0) {
        job = atomicAdd(job_counter, 1u);
        if (job == n_jobs + gridDim.x - 1u) atomicExch(job_counter, 0u); // last fetch of the launch re-arms the queue
    }
    return __shfl_sync(0xFFFFFFFFu, job, 0);
}

// A/B builds with -DAPC_BS_STATS count, per column pair (or column), how often the deep rows were computed:
// apc_microbench("bs_stats") returns computed / total and resets the counters (tools/build_ab.sh).
#ifdef APC_BS_STATS
static __device__ unsigned long long g_bs_stats[2];
#define APC_BS_STAT(run_)                                                             \
    do {                                                                              \
        if (lane == 0) {                                                              \
            atomicAdd(&g_bs_stats[1], 1ull);                                          \
            if (run_) atomicAdd(&g_bs_stats[0], 1ull);                                \
        }                                                                             \
    } while (0)
#else
#define APC_BS_STAT(run_) do { } while (0)
#endif

// Plane prefetch: the masks of a column pair are dead once they are parked in the warp's shared-memory slot, so the
// next pair is loaded into the same registers right after the staging — one pair of prefetch distance (two pairs in
// flight were measured in round 1 and again in round 2: slower; so was a prefetch.global.L1 one or two pairs ahead:
// C2 385 -> 377, C3 673 -> 665 kGCUPS), no second set of registers, no moves.
#define APC_BS_STAGE_MASKS()                                                                                          \
    s_mask[0][lane] = ma.x; s_mask[0][32 + lane] = ma.y; s_mask[0][64 + lane] = ma.z; s_mask[0][96 + lane] = ma.w;   \
    s_mask[1][lane] = mb.x; s_mask[1][32 + lane] = mb.y; s_mask[1][64 + lane] = mb.z; s_mask[1][96 + lane] = mb.w;

// The column-pair loops are unrolled twice for k >= 18: the second copy writes the plane double buffer back into the
// registers the first read it from, which removes ~19 register moves per pair (A/B on B200, tools/ab_bench.sh:
// C3, k = 20: 646 -> 675 kGCUPS; C2, k = 16: 373 -> 364, the doubled loop body costs more than the moves there).
// -DAPC_BS_PAIR_UNROLL=n forces n for every k (A/B builds).
#ifdef APC_BS_PAIR_UNROLL
__host__ __device__ constexpr int bs_pair_unroll(int) { return APC_BS_PAIR_UNROLL; }
#else
__host__ __device__ constexpr int bs_pair_unroll(int k) { return k >= 18 ? 2 : 1; }
#endif

// One k-mer per warp.  kmers[u] is the k-mer of unit u, perm[u] its index in the caller's order.
template <int K, int MB, int M = bs_check_row(K)>
__global__ void __launch_bounds__(32, MB)
bs_scan_kernel(const uint4 *__restrict__ planes, const uint32_t sg_first, const uint32_t n_sg, const uint32_t cols,
               const uint32_t read_len, const uint64_t range_lo, const uint64_t range_hi,
               const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ perm, const uint32_t n_units,
               const uint32_t sg_per_job, const uint32_t n_jobs, unsigned long long *__restrict__ counts,
               unsigned int *__restrict__ job_counter, unsigned long long *__restrict__ deep_lop3) {
    __shared__ __align__(16) uint32_t s_mask[2][4 * kGroupsPerSuper];
    const uint32_t lane = threadIdx.x;
    const uint32_t pairs = (read_len + 1) / 2; // an odd length is rounded up with one padding column (N: matches nothing)

    for (;;) {
        const uint32_t job = bs_next_job(job_counter, n_jobs, lane);
        if (job >= n_jobs) break;
        const uint32_t u = job % n_units, jb = job / n_units;
        // the first column pair of the job is requested before the row offsets are computed
        const uint4 *p = planes + ((size_t)(sg_first + jb * sg_per_job) * cols) * kGroupsPerSuper + lane;
        uint4 ma = __ldg(p), mb = __ldg(p + kGroupsPerSuper);
        const uint64_t kmer = __ldg(kmers + u);
        uint32_t off[K]; // byte offset of the mask row (A, C, G, T) that k-mer base i selects
#pragma unroll
        for (int i = 0; i < K; i++) off[i] = (uint32_t)((kmer >> (2 * (K - 1 - i))) & 3u) * kPlaneRow;

        uint32_t cnt = 0;
        uint32_t n_run = 0; // dead-row tests of this job after which the deep rows were computed (apc_scan_stats_read)
        const uint32_t sg_end = min(n_sg, (jb + 1) * sg_per_job);
        for (uint32_t sg = jb * sg_per_job; sg < sg_end; sg++) {
            uint32_t r0[K], r1[K], r2[K];
            bs_rows_init<K, 0>(r0, r1, r2);
            bool deep_zero = true; // rows M..K-2 are all zero (they start that way)
#pragma unroll(bs_pair_unroll(K))
            for (uint32_t pr = 0; pr < pairs; pr++) {
                APC_BS_STAGE_MASKS()
                p += 2 * kGroupsPerSuper; // the buffer is padded by kBsPadCols columns
                ma = __ldg(p);
                mb = __ldg(p + kGroupsPerSuper);
                const char *slot_a = reinterpret_cast<const char *>(s_mask[0]) + lane * 4;
                const char *slot_b = reinterpret_cast<const char *>(s_mask[1]) + lane * 4;
                if constexpr (M >= K) {
                    BsCarry c = bs_carry_init();
                    bs_rows<K, 0, true>(r0, r1, r2, c, slot_a, off);
                    c = bs_carry_init();
                    bs_rows<K, 0, true>(r0, r1, r2, c, slot_b, off);
                } else {
                    // top rows of both columns first (they do not depend on the deep rows), one test per column pair
                    uint32_t ck = r2[M - 1];
                    BsCarry ca = bs_carry_init();
                    bs_rows<K, 0, true, 0, M>(r0, r1, r2, ca, slot_a, off);
                    ck |= r2[M - 1];
                    BsCarry cb = bs_carry_init();
                    bs_rows<K, 0, true, 0, M>(r0, r1, r2, cb, slot_b, off);
                    ck |= r2[M - 1];
                    bool run = __any_sync(0xFFFFFFFFu, ck != 0);
                    if (!run && !deep_zero) { // first quiet pair: have the deep rows drained?
                        uint32_t z = 0;
#pragma unroll
                        for (int j = M; j < K - 1; j++) z |= r2[j];
                        run = __any_sync(0xFFFFFFFFu, z != 0);
                    }
                    deep_zero = !run;
                    n_run += run;
                    APC_BS_STAT(run);
                    if (run) {
                        bs_rows<K, 0, true, M, K>(r0, r1, r2, ca, slot_a, off);
                        bs_rows<K, 0, true, M, K>(r0, r1, r2, cb, slot_b, off);
                    }
                }
            }
            if (sg + 1 < sg_end) { // the next 1024 reads' first column pair, before the epilogue of these
                p = planes + ((size_t)(sg_first + sg + 1) * cols) * kGroupsPerSuper + lane;
                ma = __ldg(p);
                mb = __ldg(p + kGroupsPerSuper);
            }
            // hits of these 32 reads: [d<=0] + [d<=1] + [d<=2] (:589-593) = the sticky row k-1; reads outside
            // the scanned range (padding of the last group, or a sub-range scan) masked out
            const uint32_t vm = bs_valid_mask(((uint64_t)(sg_first + sg) * kGroupsPerSuper + lane) * 32, range_lo, range_hi);
            cnt += __popc(r0[K - 1] & vm) + __popc(r1[K - 1] & vm) + __popc(r2[K - 1] & vm);
        }
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0 && total) atomicAdd(&counts[__ldg(perm + u)], (unsigned long long)total);
        // LOP3 (warp instructions) spent on deep rows by this job: one test per column pair, 5 per row and column
        constexpr int kDeepLop3PerTest = 2 * 5 * bs_rows_split(K, K, 1, M).deep;
        if (lane == 0 && n_run) atomicAdd(deep_lop3, (unsigned long long)n_run * kDeepLop3PerTest);
    }
}

// G k-mers with a common prefix of P = K - T bases per warp.  Rows 0..P-1 of their tables are
// identical in every column, so they are computed once: 5(P + G T) instead of 5 G K LOP3 per
// column.  The query k-mers of the reference's pipeline are the most frequent k-mers of the
// sample, i.e. mostly an adapter's windows and their one-error variants, which share long
// prefixes — or long SUFFIXES.  The minimum edit distance of a k-mer to the substrings of a read
// is that of the reversed k-mer to the substrings of the reversed read, so a unit whose members
// share a suffix is stored reversed (bit 31 of its first perm entry) and walks the columns from
// the last to the first: same code, negative column stride.
// kmers[G*u .. G*u+G-1] are the k-mers of unit u, perm[] their indices in the caller's order.
template <int K, int P, int G, int MB, int M = bs_check_row(K)>
__global__ void __launch_bounds__(32, MB)
bs_group_kernel(const uint4 *__restrict__ planes, const uint32_t sg_first, const uint32_t n_sg, const uint32_t cols,
                const uint32_t read_len, const uint64_t range_lo, const uint64_t range_hi,
                const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ perm, const uint32_t n_units,
                const uint32_t sg_per_job, const uint32_t n_jobs, unsigned long long *__restrict__ counts,
                unsigned int *__restrict__ job_counter, unsigned long long *__restrict__ deep_lop3) {
    constexpr int T = K - P; // rows of the private tails
    static_assert(P >= 2 && T >= 1, "the always-matching rows 0..1 must lie in the shared part");
    __shared__ __align__(16) uint32_t s_mask[2][4 * kGroupsPerSuper];
    const uint32_t lane = threadIdx.x;
    const uint32_t pairs = (read_len + 1) / 2;

    for (;;) {
        const uint32_t job = bs_next_job(job_counter, n_jobs, lane);
        if (job >= n_jobs) break;
        const uint32_t u = job % n_units, jb = job / n_units;
        // direction of the walk: first column and the (signed) distance between consecutive columns
        const bool reverse = (__ldg(perm + (size_t)G * u) >> 31) != 0;
        const int64_t cstep = reverse ? -(int64_t)kGroupsPerSuper : (int64_t)kGroupsPerSuper;
        const size_t col0 = reverse ? 2 * (size_t)pairs - 1 : 0;
        // the first column pair of the job is requested before anything else: the row offsets below are computed
        // while it travels (the prologue of a job used to stall on these loads)
        const uint4 *p = planes + ((size_t)(sg_first + jb * sg_per_job) * cols + col0) * kGroupsPerSuper + lane;
        uint4 ma = __ldg(p), mb = __ldg(p + cstep);
        uint32_t off_s[P], off_t[G][T];
        {
            const uint64_t k0 = __ldg(kmers + (size_t)G * u);
#pragma unroll
            for (int i = 0; i < P; i++) off_s[i] = (uint32_t)((k0 >> (2 * (K - 1 - i))) & 3u) * kPlaneRow;
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            const uint64_t kg = __ldg(kmers + (size_t)G * u + g);
#pragma unroll
            for (int i = 0; i < T; i++) off_t[g][i] = (uint32_t)((kg >> (2 * (T - 1 - i))) & 3u) * kPlaneRow;
        }
        uint32_t cnt[G];
#pragma unroll
        for (int g = 0; g < G; g++) cnt[g] = 0;
        uint32_t n_run = 0; // dead-row tests of this job after which the deep rows were computed (apc_scan_stats_read)
        const uint32_t sg_end = min(n_sg, (jb + 1) * sg_per_job);
        for (uint32_t sg = jb * sg_per_job; sg < sg_end; sg++) {
            uint32_t s0[P], s1[P], s2[P], x0[G][T], x1[G][T], x2[G][T];
            bs_rows_init<P, 0>(s0, s1, s2);
#pragma unroll
            for (int g = 0; g < G; g++) bs_rows_init<T, P>(x0[g], x1[g], x2[g]);
            bool deep_zero = true; // the deep rows other than the hit rows are all zero (they start that way)
#pragma unroll(bs_pair_unroll(K))
            for (uint32_t pr = 0; pr < pairs; pr++) {
                APC_BS_STAGE_MASKS()
                p += 2 * cstep; // the buffer is padded by kBsPadCols columns at both ends
                ma = __ldg(p);
                mb = __ldg(p + cstep);
                if constexpr (M < K && M <= P) {
                    // top = trunk rows 0..M-1; deep = the rest of the trunk and the tails.  The top rows of both
                    // columns first (they do not depend on the deep rows), then one test per column pair.
                    const char *slot_a = reinterpret_cast<const char *>(s_mask[0]) + lane * 4;
                    const char *slot_b = reinterpret_cast<const char *>(s_mask[1]) + lane * 4;
                    uint32_t ck = s2[M - 1];
                    BsCarry ca = bs_carry_init();
                    bs_rows<P, 0, false, 0, M>(s0, s1, s2, ca, slot_a, off_s);
                    ck |= s2[M - 1];
                    BsCarry cb = bs_carry_init();
                    bs_rows<P, 0, false, 0, M>(s0, s1, s2, cb, slot_b, off_s);
                    ck |= s2[M - 1];
                    bool run = __any_sync(0xFFFFFFFFu, ck != 0);
                    if (!run && !deep_zero) { // first quiet pair: have the deep rows drained?
                        uint32_t z = 0;
#pragma unroll
                        for (int j = M; j < P; j++) z |= s2[j];
#pragma unroll
                        for (int g = 0; g < G; g++)
#pragma unroll
                            for (int j = 0; j < T - 1; j++) z |= x2[g][j];
                        run = __any_sync(0xFFFFFFFFu, z != 0);
                    }
                    deep_zero = !run;
                    n_run += run;
                    APC_BS_STAT(run);
                    if (run) {
                        bs_rows<P, 0, false, M, P>(s0, s1, s2, ca, slot_a, off_s);
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            BsCarry cg = ca;
                            bs_rows<T, P, true>(x0[g], x1[g], x2[g], cg, slot_a, off_t[g]);
                        }
                        bs_rows<P, 0, false, M, P>(s0, s1, s2, cb, slot_b, off_s);
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            BsCarry cg = cb;
                            bs_rows<T, P, true>(x0[g], x1[g], x2[g], cg, slot_b, off_t[g]);
                        }
                    }
                } else {
#pragma unroll
                for (int col = 0; col < 2; col++) {
                    const char *slot = reinterpret_cast<const char *>(s_mask[col]) + lane * 4;
                    BsCarry c = bs_carry_init();
                    if constexpr (M >= K) { // no deep part
                        bs_rows<P, 0, false>(s0, s1, s2, c, slot, off_s);
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            BsCarry cg = c;
                            bs_rows<T, P, true>(x0[g], x1[g], x2[g], cg, slot, off_t[g]);
                        }
                    } else { // M > P: top = trunk + the first M-P rows of every tail; deep = the ends of the tails (per column)
                        constexpr int S = M - P;
                        bs_rows<P, 0, false>(s0, s1, s2, c, slot, off_s);
                        BsCarry cg[G];
                        uint32_t ck = 0;
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            cg[g] = c;
                            ck |= x2[g][S - 1];
                            bs_rows<T, P, true, 0, S>(x0[g], x1[g], x2[g], cg[g], slot, off_t[g]);
                            ck |= x2[g][S - 1];
                        }
                        const bool alive = __any_sync(0xFFFFFFFFu, ck != 0);
                        bool run = alive;
                        if (!alive && !deep_zero) {
                            uint32_t z = 0;
#pragma unroll
                            for (int g = 0; g < G; g++)
#pragma unroll
                                for (int j = S; j < T - 1; j++) z |= x2[g][j];
                            run = __any_sync(0xFFFFFFFFu, z != 0);
                        }
                        deep_zero = !run;
                        n_run += run;
                    APC_BS_STAT(run);
                        if (run) {
#pragma unroll
                            for (int g = 0; g < G; g++) bs_rows<T, P, true, S, T>(x0[g], x1[g], x2[g], cg[g], slot, off_t[g]);
                        }
                    }
                }
                }
            }
            if (sg + 1 < sg_end) { // the next 1024 reads' first column pair, before the epilogue of these
                p = planes + ((size_t)(sg_first + sg + 1) * cols + col0) * kGroupsPerSuper + lane;
                ma = __ldg(p);
                mb = __ldg(p + cstep);
            }
            const uint32_t vm = bs_valid_mask(((uint64_t)(sg_first + sg) * kGroupsPerSuper + lane) * 32, range_lo, range_hi);
#pragma unroll
            for (int g = 0; g < G; g++) // the sticky last row of each tail holds the hits
                cnt[g] += __popc(x0[g][T - 1] & vm) + __popc(x1[g][T - 1] & vm) + __popc(x2[g][T - 1] & vm);
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, cnt[g]);
            if (lane == 0 && t)
                atomicAdd(&counts[__ldg(perm + (size_t)G * u + g) & 0x7FFFFFFFu], (unsigned long long)t);
        }
        if (lane == 0 && n_run) {
            constexpr int kDeepLop3PerTest = bs_rows_split(K, P, G, M).cols_per_test * 5 * bs_rows_split(K, P, G, M).deep;
            atomicAdd(deep_lop3, (unsigned long long)n_run * kDeepLop3PerTest);
        }
    }
}

// ---- launch --------------------------------------------------------------------------------------
// One launch per shape that has units, each on its own stream with its own job queue, so that the
// persistent warps of the next shape fill the SMs as those of the previous one run out of jobs.
struct BsLaunchCtx {
    const Ctx *c;
    BsRange r;
    unsigned long long *d_counts;
    uint32_t sg_per_job_opt; // 0 = choose per launch
    uint64_t *launches;
    int slot; // next stream / job counter
    // LOP3 warp instructions of this scan: in the rows computed in every column, and in all rows if nothing were skipped
    double lop3_top = 0., lop3_all = 0.;
    bool warm = false; // load every kernel of this k instead of launching (apc_reserve)
};

inline uint32_t bs_sg_per_job(const BsLaunchCtx &l, uint32_t n_units, int mb) {
    if (l.sg_per_job_opt) return l.sg_per_job_opt;
    // >= 32 jobs per resident warp where the work allows it (tail <= 3 %), jobs of at most 16 super-groups
    const uint64_t warps = (uint64_t)l.c->sm_count * mb;
    const uint64_t jobs1 = (uint64_t)l.r.n_sg * n_units;
    return (uint32_t)std::min<uint64_t>(16, std::max<uint64_t>(1, jobs1 / (warps * 32)));
}

template <typename Kernel>
static cudaError_t bs_launch_one(BsLaunchCtx &l, Kernel kernel, int mb, uint32_t first_kmer, uint32_t n_units,
                                 const BsRowSplit split) {
    const Ctx &c = *l.c;
    if (l.warm) { // CUDA loads a kernel at its first launch (lazy loading): asking for its attributes loads it now
        cudaFuncAttributes attr;
        return cudaFuncGetAttributes(&attr, kernel);
    }
    {   // rows 0 and 1 hold constants at levels 1 and 2: 7 LOP3 fewer than 5 per row (bs_rows)
        const double unit_cols = (double)n_units * l.r.n_sg * (2 * ((c.max_len + 1) / 2));
        l.lop3_top += unit_cols * (5 * split.top - 7);
        l.lop3_all += unit_cols * (5 * (split.top + split.deep) - 7);
    }
    const uint32_t spj = bs_sg_per_job(l, n_units, mb);
    const uint64_t jobs = (uint64_t)((l.r.n_sg + spj - 1) / spj) * n_units;
    if (jobs > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    const uint32_t *perm = reinterpret_cast<const uint32_t *>(c.d_kmers + c.n_kmers);
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c.sm_count * mb, jobs);
    const int slot = l.slot++;
    cudaStream_t s = c.stream;
    cudaError_t e;
    if (slot > 0) { // fork: the side stream starts after everything queued on the caller's stream so far
        s = c.bs_streams[slot - 1];
        if ((e = cudaStreamWaitEvent(s, c.bs_fork, 0)) != cudaSuccess) return e;
    }
    kernel<<<grid, 32, 0, s>>>(c.planes(), l.r.sg_first, l.r.n_sg, c.chunks * kChunkBases, c.max_len, l.r.lo, l.r.hi,
                              c.d_kmers + first_kmer, perm + first_kmer, n_units, spj, (uint32_t)jobs, l.d_counts,
                              c.d_job_counter + slot, c.d_deep_lop3);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    (*l.launches)++;
    if (slot > 0) { // join
        if ((e = cudaEventRecord(c.bs_join[slot - 1], s)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(c.stream, c.bs_join[slot - 1], 0)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int K, int S>
static cudaError_t bs_launch_shapes(BsLaunchCtx &l, uint32_t &first_kmer) {
    if constexpr (S < kBsShapes) {
        constexpr BsShape sh = bs_shape(K, S);
        if constexpr (sh.g > 0) {
            const uint32_t n_units = l.c->bs_units[S];
            if (n_units || l.warm) {
                constexpr int MB = bs_warps_per_sm_c(K - sh.t + sh.g * sh.t);
                cudaError_t e = bs_launch_one(l, bs_group_kernel<K, K - sh.t, sh.g, MB>, MB, first_kmer, n_units,
                                              bs_rows_split(K, K - sh.t, sh.g, bs_check_row(K)));
                if (e != cudaSuccess) return e;
                first_kmer += n_units * sh.g;
            }
        }
        return bs_launch_shapes<K, S + 1>(l, first_kmer);
    } else {
        return cudaSuccess;
    }
}

template <int K>
static cudaError_t launch_bs_k(BsLaunchCtx &l) {
    const Ctx &c = *l.c;
    uint32_t first = 0;
    // shapes in table order (the heaviest units first), then the k-mers that found no partner
    cudaError_t e = bs_launch_shapes<K, 0>(l, first);
    if (e != cudaSuccess) return e;
    if (first < c.n_kmers || l.warm) {
        // registers: 3K of state + the row masks of two columns in flight (ptxas wants about 6K + 26);
        // CTAs (= warps) per SM chosen so that nothing spills
        constexpr int MB = bs_warps_per_sm_c(K);
        e = bs_launch_one(l, bs_scan_kernel<K, MB>, MB, first, c.n_kmers - first, bs_rows_split(K, K, 1, bs_check_row(K)));
    }
    return e;
}

} // namespace apc
