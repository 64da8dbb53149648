// bitslice_kernel.cu — K1 in its bit-sliced form: the approximate-count hot path with
// the READS, not the k-mer rows, packed into machine words.
//
// Same automaton as scan_kernel.cu (Wu–Manber, three error levels; replaces errorCount,
// /root/reference/approx_counter.cpp:531-601), transposed: a 32-bit word holds one cell
// (row i, level e) of 32 different reads.  Per text column and row
//     R0'[i] =  R0[i-1] & Eq_i
//     R1'[i] = (R1[i-1] & Eq_i) | R0[i] | R0[i-1] | R0'[i-1]
//     R2'[i] = (R2[i-1] & Eq_i) | R1[i] | R1[i-1] | R1'[i-1]          (R_e[-1] = all ones)
// with Eq_i = "the read's base in this column equals base i of the k-mer", one bit per read.
// That is 5 LOP3 per row for 32 reads — no shifts at all (the row packing of
// scan_kernel.cu needs 3 IMAD + 6.5 LOP3 + 0.25 PRMT per 32 cells) — and hits are
// accumulated from row k-1 only (1.5 LOP3 per COLUMN instead of per word).  Per (read,
// k-mer, column) the ALU pipe sees (5k - 6 + 1.5) / 32 instructions: 2.4 at k=16 against 3.4.
//
// B200 mapping:
//  * text = bit planes in HBM: for every group of 32 reads and every column one uint4
//    {A, C, G, T} of 32-bit masks (N and padding set no bit); 32 groups form a super-group
//    and are interleaved so that lane = group makes a column one coalesced 512-byte load.
//    0.5 byte per base; built once per sample from the scan tiles (build_planes_kernel).
//  * one warp per CTA; lane = read group, so a warp advances 1024 reads x ONE k-mer.  The
//    k-mer is warp-uniform: which of the four masks row i needs is a uniform byte offset.
//    Each lane parks its four masks in shared memory ([base][lane], conflict-free) and
//    reads row i's mask back with LDS [lane*4 + UR_i]: the selection costs no ALU
//    instruction, the offsets live in uniform registers (ptxas keeps the k-mer uniform
//    through the SHFL broadcast of the job number and the load from a uniform address).
//  * all k rows of a column are independent given the previous column, so a single warp
//    has 3k-way instruction-level parallelism; state = 3k registers per thread
//    (launch bounds per k), no shared state between warps, no barriers.
//  * persistent warps pull (k-mer, super-group range) jobs from the self re-arming queue
//    used by scan_kernel.cu, k-mer fastest: warps running together share text in L2.
//  * hits: per super-group popc of the three accumulators under the valid-read mask,
//    one REDUX + atomicAdd per job.
#include <algorithm>

#include "apc_internal.h"
#include "scan_core.cuh"

namespace apc {

constexpr int kGroupsPerSuper = 32; // lanes
constexpr int bs_warps_per_sm_c(int k) {
#ifdef APC_BS_MB_OVERRIDE
    (void)k;
    return APC_BS_MB_OVERRIDE; // A/B builds (tools/build_ab.sh)
#endif
    // registers are handed out per SM sub-partition (16384 each), so only multiples of 4 warps matter:
    // 24 -> 80 registers, 20 -> 96, 16 -> 128, 12 -> 168, 8 -> 255
    return k <= 8 ? 24 : k <= 12 ? 20 : k <= 16 ? 16 : k <= 24 ? 12 : 8;
}
// the quad kernel carries 3(3k/4 + 4(k - 3k/4)) state registers: built where that fits without spills
constexpr bool bs_quads_ok(int k) { return k >= 8 && k <= 20; }
constexpr int kPlaneRow = 128;      // bytes between the A, C, G, T rows of the per-warp mask slot

// ---- K2b: scan tiles -> bit planes ------------------------------------------------------
// One warp per tile (= one 32-read group): lane = read, four ballots per column.
__global__ void __launch_bounds__(256)
build_planes_kernel(const uint4 *__restrict__ tiles, const uint32_t n_tiles, const uint32_t chunks,
                    uint4 *__restrict__ planes) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const uint32_t sg = tile / kGroupsPerSuper, grp = tile % kGroupsPerSuper;
    const uint32_t cols = chunks * kChunkBases;
    for (uint32_t ch = 0; ch < chunks; ch++) {
        const uint4 v = __ldg(tiles + ((size_t)tile * chunks + ch) * kTileReads + lane);
        const uint32_t tw[4] = {v.x, v.y, v.z, v.w};
        uint4 mine = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const uint32_t code = (tw[j >> 2] >> (8 * (j & 3))) & 0xFFu; // 0x00,0x10,0x20,0x30 | 0x40 = N/pad
            const uint32_t a = __ballot_sync(0xFFFFFFFFu, code == 0x00u), c = __ballot_sync(0xFFFFFFFFu, code == 0x10u);
            const uint32_t g = __ballot_sync(0xFFFFFFFFu, code == 0x20u), t = __ballot_sync(0xFFFFFFFFu, code == 0x30u);
            if (lane == (uint32_t)j) mine = make_uint4(a, c, g, t);
        }
        if (lane < 16) planes[((size_t)sg * cols + ch * kChunkBases + lane) * kGroupsPerSuper + grp] = mine;
    }
}

cudaError_t launch_build_planes(const Ctx &c) {
    if (c.n_tiles == 0 || c.chunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(c.d_planes, 0, c.planes_bytes, c.stream); // groups past the last tile match nothing
    if (e != cudaSuccess) return e;
    const unsigned blocks = (c.n_tiles + 7) / 8;
    build_planes_kernel<<<blocks, 256, 0, c.stream>>>(c.d_tiles, c.n_tiles, c.chunks, c.d_planes);
    return cudaGetLastError();
}

// ---- K1 (bit-sliced) --------------------------------------------------------------------------
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
template <int N, int FIRST>
__device__ __forceinline__ void bs_rows(uint32_t (&r0)[N], uint32_t (&r1)[N], uint32_t (&r2)[N], BsCarry &c,
                                        const char *slot_lane, const uint32_t (&off)[N]) {
    const uint32_t ALL = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < N; j++) {
        constexpr int dummy = 0;
        (void)dummy;
        const int i = FIRST + j;
        const uint32_t e = *reinterpret_cast<const uint32_t *>(slot_lane + off[j]);
        const uint32_t o0 = r0[j], o1 = r1[j], o2 = r2[j];
        const uint32_t n0 = i < 1 ? e : and2(c.p0, e);
        // rows 0 (level 1) and 0..1 (level 2) always match: that many k-mer bases can be skipped
        const uint32_t n1 = i < 1 ? ALL : or3(and_or(c.p1, e, o0), c.p0, c.n0p);
        const uint32_t n2 = i < 2 ? ALL : or3(and_or(c.p2, e, o1), c.p1, c.n1p);
        r0[j] = n0; r1[j] = n1; r2[j] = n2;
        c.p0 = o0; c.p1 = o1; c.p2 = o2;
        c.n0p = n0; c.n1p = n1;
    }
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

// job fetch of the persistent warps: warp-uniform result (ptxas keeps it in uniform registers)
__device__ __forceinline__ uint32_t bs_next_job(unsigned int *job_counter, uint32_t n_jobs, uint32_t lane) {
    uint32_t job = 0;
    if (lane == 0) {
        job = atomicAdd(job_counter, 1u);
        if (job == n_jobs + gridDim.x - 1u) atomicExch(job_counter, 0u); // last fetch of the launch re-arms the queue
    }
    return __shfl_sync(0xFFFFFFFFu, job, 0);
}

#define APC_BS_STAGE_MASKS()                                                                                          \
    s_mask[0][lane] = ma.x; s_mask[0][32 + lane] = ma.y; s_mask[0][64 + lane] = ma.z; s_mask[0][96 + lane] = ma.w;   \
    s_mask[1][lane] = mb.x; s_mask[1][32 + lane] = mb.y; s_mask[1][64 + lane] = mb.z; s_mask[1][96 + lane] = mb.w;

// One k-mer per warp.  kmers[q0 + u] is the k-mer of unit u, perm[q0 + u] its index in the caller's order.
template <int K, int MB>
__global__ void __launch_bounds__(32, MB)
bs_scan_kernel(const uint4 *__restrict__ planes, const uint32_t sg_first, const uint32_t n_sg, const uint32_t cols,
               const uint32_t read_len, const uint64_t range_lo, const uint64_t range_hi,
               const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ perm, const uint32_t n_units,
               const uint32_t sg_per_job, const uint32_t n_jobs, unsigned long long *__restrict__ counts,
               unsigned int *__restrict__ job_counter) {
    __shared__ __align__(16) uint32_t s_mask[2][4 * kGroupsPerSuper];
    const uint32_t lane = threadIdx.x;
    const uint32_t ALL = 0xFFFFFFFFu;
    const uint32_t pairs = (read_len + 1) / 2; // an odd length is rounded up with one padding column (N: matches nothing)

    for (;;) {
        const uint32_t job = bs_next_job(job_counter, n_jobs, lane);
        if (job >= n_jobs) break;
        const uint32_t u = job % n_units, jb = job / n_units;
        const uint64_t kmer = __ldg(kmers + u);
        uint32_t off[K]; // byte offset of the mask row (A, C, G, T) that k-mer base i selects
#pragma unroll
        for (int i = 0; i < K; i++) off[i] = (uint32_t)((kmer >> (2 * (K - 1 - i))) & 3u) * kPlaneRow;

        uint32_t cnt = 0;
        const uint32_t sg_end = min(n_sg, (jb + 1) * sg_per_job);
        for (uint32_t sg = jb * sg_per_job; sg < sg_end; sg++) {
            uint32_t r0[K], r1[K], r2[K];
            bs_rows_init<K, 0>(r0, r1, r2);
            uint32_t a0 = 0, a1 = K <= 1 ? ALL : 0, a2 = K <= 2 ? ALL : 0;
            const uint4 *p = planes + ((size_t)(sg_first + sg) * cols) * kGroupsPerSuper + lane;
            uint4 ma = __ldg(p), mb = __ldg(p + kGroupsPerSuper);
            for (uint32_t pr = 0; pr < pairs; pr++) {
                p += 2 * kGroupsPerSuper;
                const uint4 na = __ldg(p), nb = __ldg(p + kGroupsPerSuper); // buffer is padded by two columns
                APC_BS_STAGE_MASKS()
                BsCarry c = bs_carry_init();
                bs_rows<K, 0>(r0, r1, r2, c, reinterpret_cast<const char *>(s_mask[0]) + lane * 4, off);
                const uint32_t h0 = r0[K - 1], h1 = r1[K - 1], h2 = r2[K - 1];
                c = bs_carry_init();
                bs_rows<K, 0>(r0, r1, r2, c, reinterpret_cast<const char *>(s_mask[1]) + lane * 4, off);
                a0 = or3(a0, h0, r0[K - 1]);
                a1 = or3(a1, h1, r1[K - 1]);
                a2 = or3(a2, h2, r2[K - 1]);
                ma = na; mb = nb;
            }
            // hits of these 32 reads: [d<=0] + [d<=1] + [d<=2] (:589-593), reads outside the
            // scanned range (padding of the last group, or a sub-range scan) masked out
            const uint32_t vm = bs_valid_mask(((uint64_t)(sg_first + sg) * kGroupsPerSuper + lane) * 32, range_lo, range_hi);
            cnt += __popc(a0 & vm) + __popc(a1 & vm) + __popc(a2 & vm);
        }
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0 && total) atomicAdd(&counts[__ldg(perm + u)], (unsigned long long)total);
    }
}

// G k-mers with a common prefix of at least P bases per warp.  Rows 0..P-1 of their tables
// are identical in every column, so they are computed once: 5(P + G(K-P)) instead of 5GK LOP3
// per column.  The query k-mers of the reference's pipeline are the most frequent k-mers of the
// sample, i.e. mostly an adapter's windows and their one-error variants, which share long
// prefixes once sorted (C2: quads with P = 3k/4 and pairs with P = k/2 cut the row work to 0.71).
// kmers[G*u .. G*u+G-1] are the k-mers of unit u, perm[] their indices in the caller's order.
template <int K, int P, int G, int MB>
__global__ void __launch_bounds__(32, MB)
bs_group_kernel(const uint4 *__restrict__ planes, const uint32_t sg_first, const uint32_t n_sg, const uint32_t cols,
                const uint32_t read_len, const uint64_t range_lo, const uint64_t range_hi,
                const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ perm, const uint32_t n_units,
                const uint32_t sg_per_job, const uint32_t n_jobs, unsigned long long *__restrict__ counts,
                unsigned int *__restrict__ job_counter) {
    constexpr int T = K - P; // rows of the private tails
    static_assert(P >= 2 && T >= 1, "the always-matching rows 0..1 must lie in the shared part");
    __shared__ __align__(16) uint32_t s_mask[2][4 * kGroupsPerSuper];
    const uint32_t lane = threadIdx.x;
    const uint32_t pairs = (read_len + 1) / 2;

    for (;;) {
        const uint32_t job = bs_next_job(job_counter, n_jobs, lane);
        if (job >= n_jobs) break;
        const uint32_t u = job % n_units, jb = job / n_units;
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
        const uint32_t sg_end = min(n_sg, (jb + 1) * sg_per_job);
        for (uint32_t sg = jb * sg_per_job; sg < sg_end; sg++) {
            uint32_t s0[P], s1[P], s2[P], x0[G][T], x1[G][T], x2[G][T], a0[G], a1[G], a2[G];
            bs_rows_init<P, 0>(s0, s1, s2);
#pragma unroll
            for (int g = 0; g < G; g++) {
                bs_rows_init<T, P>(x0[g], x1[g], x2[g]);
                a0[g] = a1[g] = a2[g] = 0;
            }
            const uint4 *p = planes + ((size_t)(sg_first + sg) * cols) * kGroupsPerSuper + lane;
            uint4 ma = __ldg(p), mb = __ldg(p + kGroupsPerSuper);
            for (uint32_t pr = 0; pr < pairs; pr++) {
                p += 2 * kGroupsPerSuper;
                const uint4 na = __ldg(p), nb = __ldg(p + kGroupsPerSuper);
                APC_BS_STAGE_MASKS()
                uint32_t h0[G], h1[G], h2[G];
#pragma unroll
                for (int col = 0; col < 2; col++) {
                    const char *slot = reinterpret_cast<const char *>(s_mask[col]) + lane * 4;
                    BsCarry c = bs_carry_init();
                    bs_rows<P, 0>(s0, s1, s2, c, slot, off_s);
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        BsCarry cg = c;
                        bs_rows<T, P>(x0[g], x1[g], x2[g], cg, slot, off_t[g]);
                        if (col == 0) {
                            h0[g] = x0[g][T - 1]; h1[g] = x1[g][T - 1]; h2[g] = x2[g][T - 1];
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                    a0[g] = or3(a0[g], h0[g], x0[g][T - 1]);
                    a1[g] = or3(a1[g], h1[g], x1[g][T - 1]);
                    a2[g] = or3(a2[g], h2[g], x2[g][T - 1]);
                }
                ma = na; mb = nb;
            }
            const uint32_t vm = bs_valid_mask(((uint64_t)(sg_first + sg) * kGroupsPerSuper + lane) * 32, range_lo, range_hi);
#pragma unroll
            for (int g = 0; g < G; g++) cnt[g] += __popc(a0[g] & vm) + __popc(a1[g] & vm) + __popc(a2[g] & vm);
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, cnt[g]);
            if (lane == 0 && t) atomicAdd(&counts[__ldg(perm + (size_t)G * u + g)], (unsigned long long)t);
        }
    }
}

struct BsRange {
    uint32_t sg_first, n_sg;
    uint64_t lo, hi;
};

template <int K, int P, int G>
static cudaError_t launch_bs_group(const Ctx &c, const BsRange &r, unsigned long long *d_counts, uint32_t sg_blocks,
                                   uint32_t sg_per_job, uint32_t first_kmer, uint32_t n_units, uint64_t *launches) {
    constexpr int MB = bs_warps_per_sm_c(P + G * (K - P)); // same register need as a single k-mer of that many rows
    const uint64_t jobs = (uint64_t)sg_blocks * n_units;
    if (jobs > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    const uint32_t *perm = reinterpret_cast<const uint32_t *>(c.d_kmers + c.n_kmers);
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c.sm_count * MB, jobs);
    bs_group_kernel<K, P, G, MB><<<grid, 32, 0, c.stream>>>(
        c.d_planes, r.sg_first, r.n_sg, c.chunks * kChunkBases, c.max_len, r.lo, r.hi, c.d_kmers + first_kmer,
        perm + first_kmer, n_units, sg_per_job, (uint32_t)jobs, d_counts, c.d_job_counter);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) (*launches)++;
    return e;
}

template <int K>
static cudaError_t launch_bs_k(const Ctx &c, const BsRange &r, unsigned long long *d_counts, uint32_t sg_per_job,
                               uint64_t *launches) {
    const uint32_t sg_blocks = (r.n_sg + sg_per_job - 1) / sg_per_job;
    // quads first (the longest jobs), then pairs, then the k-mers that found no partner
    if constexpr (bs_quads_ok(K)) {
        if (c.n_quads) {
            cudaError_t e = launch_bs_group<K, 3 * K / 4, 4>(c, r, d_counts, sg_blocks, sg_per_job, 0, c.n_quads, launches);
            if (e != cudaSuccess) return e;
        }
    }
    if constexpr (K >= 4) {
        if (c.n_pairs) {
            cudaError_t e = launch_bs_group<K, K / 2, 2>(c, r, d_counts, sg_blocks, sg_per_job, 4 * c.n_quads, c.n_pairs, launches);
            if (e != cudaSuccess) return e;
        }
    }
    const uint32_t first_single = 4 * c.n_quads + 2 * c.n_pairs;
    const uint32_t n_single = c.n_kmers - first_single;
    if (n_single) {
        // registers: 3K of state + the row masks of two columns in flight (ptxas wants about 6K + 26);
        // CTAs (= warps) per SM chosen so that nothing spills
        constexpr int MB = bs_warps_per_sm_c(K);
        const uint64_t jobs = (uint64_t)sg_blocks * n_single;
        if (jobs > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
        const uint32_t *perm = reinterpret_cast<const uint32_t *>(c.d_kmers + c.n_kmers);
        const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c.sm_count * MB, jobs);
        bs_scan_kernel<K, MB><<<grid, 32, 0, c.stream>>>(c.d_planes, r.sg_first, r.n_sg, c.chunks * kChunkBases, c.max_len,
                                                         r.lo, r.hi, c.d_kmers + first_single, perm + first_single,
                                                         n_single, sg_per_job, (uint32_t)jobs, d_counts, c.d_job_counter);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        (*launches)++;
    }
    return cudaSuccess;
}

int bs_warps_per_sm(int k) { return bs_warps_per_sm_c(k); }

// Host side of the grouping: sort the k-mers, then greedily take runs of four neighbours whose
// common prefix is at least 3k/4 bases (where the quad kernel exists), then pairs of neighbours
// with a common prefix of at least k/2.  order[] receives the k-mer indices: quad members
// first, then pair members, then the rest.
bool bs_quads_available(int k) { return bs_quads_ok(k); }

void bs_group_queries(const uint64_t *kmers, uint32_t n, int k, bool enable, std::vector<uint32_t> &order,
                      uint32_t &n_quads, uint32_t &n_pairs) {
    order.resize(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    n_quads = n_pairs = 0;
    if (!enable || k < 4 || n < 2) return;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return kmers[a] != kmers[b] ? kmers[a] < kmers[b] : a < b; });
    auto share = [&](uint32_t i, uint32_t j, int p) { // sorted positions i < j: common prefix of at least p bases
        return ((kmers[order[i]] ^ kmers[order[j]]) >> (2 * (k - p))) == 0;
    };
    std::vector<uint8_t> used(n, 0);
    std::vector<uint32_t> quads, pairs, singles;
    if (bs_quads_ok(k)) {
        const int p4 = 3 * k / 4;
        for (uint32_t i = 0; i + 3 < n;) {
            if (share(i, i + 3, p4)) { // sorted, so the two in between share it too
                for (int t = 0; t < 4; t++) { quads.push_back(order[i + t]); used[i + t] = 1; }
                i += 4;
            } else {
                i += 1;
            }
        }
    }
    const int p2 = k / 2;
    for (uint32_t i = 0; i < n;) {
        if (used[i]) { i++; continue; }
        if (i + 1 < n && !used[i + 1] && share(i, i + 1, p2)) {
            pairs.push_back(order[i]);
            pairs.push_back(order[i + 1]);
            i += 2;
        } else {
            singles.push_back(order[i]);
            i += 1;
        }
    }
    n_quads = (uint32_t)(quads.size() / 4);
    n_pairs = (uint32_t)(pairs.size() / 2);
    std::copy(quads.begin(), quads.end(), order.begin());
    std::copy(pairs.begin(), pairs.end(), order.begin() + quads.size());
    std::copy(singles.begin(), singles.end(), order.begin() + quads.size() + pairs.size());
}

cudaError_t launch_bs_scan(const Ctx &c, uint64_t lo, uint64_t hi, unsigned long long *d_counts, uint32_t sg_per_job,
                           uint64_t *launches) {
    BsRange r;
    r.lo = lo;
    r.hi = hi;
    r.sg_first = (uint32_t)(lo / (32 * kGroupsPerSuper));
    r.n_sg = (uint32_t)((hi + 32 * kGroupsPerSuper - 1) / (32 * kGroupsPerSuper)) - r.sg_first;
    *launches = 0;
    if (r.n_sg == 0 || c.n_kmers == 0) return cudaSuccess;
    switch (c.k) {
#define APC_BS_CASE(K_) case K_: return launch_bs_k<K_>(c, r, d_counts, sg_per_job, launches);
        APC_BS_CASE(2) APC_BS_CASE(3) APC_BS_CASE(4) APC_BS_CASE(5) APC_BS_CASE(6) APC_BS_CASE(7) APC_BS_CASE(8)
        APC_BS_CASE(9) APC_BS_CASE(10) APC_BS_CASE(11) APC_BS_CASE(12) APC_BS_CASE(13) APC_BS_CASE(14)
        APC_BS_CASE(15) APC_BS_CASE(16) APC_BS_CASE(17) APC_BS_CASE(18) APC_BS_CASE(19) APC_BS_CASE(20)
        APC_BS_CASE(21) APC_BS_CASE(22) APC_BS_CASE(23) APC_BS_CASE(24) APC_BS_CASE(25) APC_BS_CASE(26)
        APC_BS_CASE(27) APC_BS_CASE(28) APC_BS_CASE(29) APC_BS_CASE(30) APC_BS_CASE(31) APC_BS_CASE(32)
#undef APC_BS_CASE
    default: return cudaErrorInvalidValue;
    }
}

} // namespace apc
