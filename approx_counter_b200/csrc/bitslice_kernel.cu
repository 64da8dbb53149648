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
template <int K>
struct BsState {
    uint32_t r0[K], r1[K], r2[K];
};

// one text column: e_i from the warp's mask slot, then the K rows
template <int K>
__device__ __forceinline__ void bs_column(BsState<K> &st, const char *slot_lane, const uint32_t (&off)[K]) {
    const uint32_t ALL = 0xFFFFFFFFu;
    uint32_t p0 = ALL, p1 = ALL, p2 = ALL; // previous column's row i-1 (row -1 = empty prefix: always matches)
    uint32_t n0p = ALL, n1p = ALL;         // this column's row i-1, levels 0 and 1
#pragma unroll
    for (int i = 0; i < K; i++) {
        const uint32_t e = *reinterpret_cast<const uint32_t *>(slot_lane + off[i]);
        const uint32_t o0 = st.r0[i], o1 = st.r1[i], o2 = st.r2[i];
        const uint32_t n0 = i < 1 ? e : and2(p0, e);
        // rows 0 (level 1) and 0..1 (level 2) always match: that many k-mer bases can be skipped
        const uint32_t n1 = i < 1 ? ALL : or3(and_or(p1, e, o0), p0, n0p);
        const uint32_t n2 = i < 2 ? ALL : or3(and_or(p2, e, o1), p1, n1p);
        st.r0[i] = n0; st.r1[i] = n1; st.r2[i] = n2;
        p0 = o0; p1 = o1; p2 = o2;
        n0p = n0; n1p = n1;
    }
}

template <int K, int MB>
__global__ void __launch_bounds__(32, MB)
bs_scan_kernel(const uint4 *__restrict__ planes, const uint32_t sg_first, const uint32_t n_sg, const uint32_t cols,
               const uint32_t read_len, const uint64_t range_lo, const uint64_t range_hi,
               const uint64_t *__restrict__ kmers, const uint32_t n_kmers, const uint32_t sg_per_job,
               const uint32_t n_jobs, unsigned long long *__restrict__ counts,
               unsigned int *__restrict__ job_counter) {
    __shared__ __align__(16) uint32_t s_mask[2][4 * kGroupsPerSuper];
    const uint32_t lane = threadIdx.x;
    const uint32_t ALL = 0xFFFFFFFFu;
    const uint32_t pairs = (read_len + 1) / 2; // an odd length is rounded up with one padding column (N: matches nothing)

    for (;;) {
        uint32_t job = 0;
        if (lane == 0) {
            job = atomicAdd(job_counter, 1u);
            if (job == n_jobs + gridDim.x - 1u) atomicExch(job_counter, 0u); // last fetch of the launch re-arms the queue
        }
        job = __shfl_sync(0xFFFFFFFFu, job, 0); // warp-uniform from here on (ptxas keeps it in uniform registers)
        if (job >= n_jobs) break;
        const uint32_t q = job % n_kmers, jb = job / n_kmers;
        const uint64_t kmer = __ldg(kmers + q);
        uint32_t off[K]; // byte offset of the mask row (A, C, G, T) that k-mer base i selects
#pragma unroll
        for (int i = 0; i < K; i++) off[i] = (uint32_t)((kmer >> (2 * (K - 1 - i))) & 3u) * kPlaneRow;

        uint32_t cnt = 0;
        const uint32_t sg_end = min(n_sg, (jb + 1) * sg_per_job);
        for (uint32_t sg = jb * sg_per_job; sg < sg_end; sg++) {
            BsState<K> st;
#pragma unroll
            for (int i = 0; i < K; i++) {
                st.r0[i] = 0;
                st.r1[i] = i < 1 ? ALL : 0; // prefix 1 by one deletion
                st.r2[i] = i < 2 ? ALL : 0; // prefixes 1..2 by deletions
            }
            uint32_t a0 = 0, a1 = K <= 1 ? ALL : 0, a2 = K <= 2 ? ALL : 0;
            const uint4 *p = planes + ((size_t)(sg_first + sg) * cols) * kGroupsPerSuper + lane;
            uint4 ma = __ldg(p), mb = __ldg(p + kGroupsPerSuper);
            for (uint32_t pr = 0; pr < pairs; pr++) {
                p += 2 * kGroupsPerSuper;
                const uint4 na = __ldg(p), nb = __ldg(p + kGroupsPerSuper); // buffer is padded by two columns
                s_mask[0][lane] = ma.x; s_mask[0][32 + lane] = ma.y; s_mask[0][64 + lane] = ma.z; s_mask[0][96 + lane] = ma.w;
                s_mask[1][lane] = mb.x; s_mask[1][32 + lane] = mb.y; s_mask[1][64 + lane] = mb.z; s_mask[1][96 + lane] = mb.w;
                bs_column<K>(st, reinterpret_cast<const char *>(s_mask[0]) + lane * 4, off);
                const uint32_t h0 = st.r0[K - 1], h1 = st.r1[K - 1], h2 = st.r2[K - 1];
                bs_column<K>(st, reinterpret_cast<const char *>(s_mask[1]) + lane * 4, off);
                a0 = or3(a0, h0, st.r0[K - 1]);
                a1 = or3(a1, h1, st.r1[K - 1]);
                a2 = or3(a2, h2, st.r2[K - 1]);
                ma = na; mb = nb;
            }
            // hits of these 32 reads: [d<=0] + [d<=1] + [d<=2] (:589-593), reads outside the
            // scanned range (padding of the last group, or a sub-range scan) masked out
            const uint64_t first = ((uint64_t)(sg_first + sg) * kGroupsPerSuper + lane) * 32;
            uint32_t vm = 0;
            if (first < range_hi && first + 32 > range_lo) {
                vm = ALL;
                if (range_lo > first) vm &= ALL << (uint32_t)(range_lo - first);
                if (range_hi < first + 32) vm &= ALL >> (uint32_t)(first + 32 - range_hi);
            }
            cnt += __popc(a0 & vm) + __popc(a1 & vm) + __popc(a2 & vm);
        }
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0 && total) atomicAdd(&counts[q], (unsigned long long)total);
    }
}

template <int K>
static cudaError_t launch_bs_k(const Ctx &c, uint64_t lo, uint64_t hi, unsigned long long *d_counts, uint32_t sg_per_job) {
    // registers: 3K of state + the row masks of two columns in flight (ptxas wants about 6K + 26);
    // CTAs (= warps) per SM chosen so that nothing spills
    constexpr int MB = bs_warps_per_sm_c(K);
    const uint32_t sg_first = (uint32_t)(lo / (32 * kGroupsPerSuper));
    const uint32_t sg_last = (uint32_t)((hi + 32 * kGroupsPerSuper - 1) / (32 * kGroupsPerSuper));
    const uint32_t n_sg = sg_last - sg_first;
    const uint64_t jobs = (uint64_t)((n_sg + sg_per_job - 1) / sg_per_job) * c.n_kmers;
    if (jobs == 0) return cudaSuccess;
    if (jobs > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c.sm_count * MB, jobs);
    bs_scan_kernel<K, MB><<<grid, 32, 0, c.stream>>>(c.d_planes, sg_first, n_sg, c.chunks * kChunkBases, c.max_len, lo, hi,
                                                     c.d_kmers, c.n_kmers, sg_per_job, (uint32_t)jobs, d_counts,
                                                     c.d_job_counter);
    return cudaGetLastError();
}

int bs_warps_per_sm(int k) { return bs_warps_per_sm_c(k); }

cudaError_t launch_bs_scan(const Ctx &c, uint64_t lo, uint64_t hi, unsigned long long *d_counts, uint32_t sg_per_job) {
    switch (c.k) {
#define APC_BS_CASE(K_) case K_: return launch_bs_k<K_>(c, lo, hi, d_counts, sg_per_job);
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
