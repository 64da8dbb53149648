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
//  * k-mers that share a prefix (or, scanned backwards, a suffix) are grouped into units whose
//    common rows are computed once (bs_group_kernel, shapes in apc_internal.h); the grouping is a
//    dynamic programme over the sorted k-mers on the host (bs_group_queries below).
//
// The kernels themselves live in bitslice_core.cuh and are instantiated by bitslice_part.cu
// (one object per slice of k); this file holds the plane builder, the grouping and the dispatch.
#include <algorithm>
#include <numeric>

#include "bitslice_core.cuh"

namespace apc {

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
    // groups past the last tile match nothing; the padding columns in front and behind are loaded but never used
    cudaError_t e = cudaMemsetAsync(c.d_planes, 0, c.planes_bytes + 2 * kBsPadCols * kGroupsPerSuper * sizeof(uint4), c.stream);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (c.n_tiles + 7) / 8;
    build_planes_kernel<<<blocks, 256, 0, c.stream>>>(c.d_tiles, c.n_tiles, c.chunks, c.planes());
    return cudaGetLastError();
}

// ---- host side of the grouping ---------------------------------------------------------------------
uint64_t bs_reverse_kmer(uint64_t kmer, int k) { // base order reversed (NOT complemented)
    // swap the 2-bit groups of the word end for end, then drop the 32 - k unused groups
    uint64_t r = kmer;
    r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
    r = ((r >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((r & 0x0F0F0F0F0F0F0F0Full) << 4);
    r = __builtin_bswap64(r);
    return r >> (64 - 2 * k);
}

namespace {

// Estimated cost of a unit in row-equivalents (one row = 5 LOP3 per column and 32 reads).  Measured
// (tools/shape_bench.py, profiles/r01_shape_bench.jsonl): the time of a unit is proportional to the rows it
// computes whatever the shape, plus about half a row for a grouped unit (fewer resident warps than the
// one-k-mer kernel).  With dead-row skipping (bitslice_core.cuh) a unit computes all its rows only in the
// `alive` share of the columns and its top rows in the others.
inline float bs_unit_cost(int k, int t, int g, float alive) {
    const int p = k - t, rows = p + g * t, m = bs_check_row_host(k);
    const int top = m >= k ? rows : m <= p ? m : p + g * (m - p);
    return alive * (float)rows + (1.f - alive) * (float)top + (g > 1 ? 0.5f : 0.f);
}

struct BsGroup {
    uint32_t first; // position in the sorted array
    int shape;      // -1 = single
};

struct BsShapeSet { // the shapes available for this k, most members first
    int n = 0, max_t = 0;
    int id[kBsShapes], g[kBsShapes], shift[kBsShapes];
    float cost[kBsShapes];
    float single = 0.f;
    BsShapeSet(int k, uint32_t shape_mask, float alive) {
        for (int s = 0; s < kBsShapes; s++) {
            const BsShape sh = (shape_mask >> s) & 1u ? bs_shape(k, s) : BsShape{0, 0};
            if (!sh.g) continue;
            id[n] = s;
            g[n] = sh.g;
            shift[n] = 2 * sh.t;
            cost[n] = bs_unit_cost(k, sh.t, sh.g, alive);
            max_t = std::max(max_t, sh.t);
            n++;
        }
        single = bs_unit_cost(k, k, 1, alive); // one k-mer: no trunk, one tail of k rows
    }
};

// Cheapest cover of the sorted k-mers v[0..n) by units (runs of g neighbours whose common prefix has at
// least k - t bases) and singles.
void bs_cover(const uint64_t *v, size_t n, const BsShapeSet &ss, std::vector<float> &f, std::vector<int8_t> &choice,
              std::vector<BsGroup> &groups) {
    f.resize(n + 1);
    choice.resize(n + 1);
    f[0] = 0.f;
    const int max_shift = std::min(2 * ss.max_t, 63);
    for (size_t i = 1; i <= n; i++) {
        float best = f[i - 1] + ss.single;
        int8_t pick = -1;
        // no shape fits unless the last two of the run share at least k - max_t bases
        if (i >= 2 && ((v[i - 2] ^ v[i - 1]) >> max_shift) == 0) {
            for (int s = 0; s < ss.n; s++) {
                const size_t g = (size_t)ss.g[s];
                if (g > i) continue;
                // sorted: the first and the last of the run share the prefix, so all in between do
                if (((v[i - g] ^ v[i - 1]) >> ss.shift[s]) != 0) continue;
                const float c = f[i - g] + ss.cost[s];
                if (c < best) {
                    best = c;
                    pick = (int8_t)s;
                }
            }
        }
        f[i] = best;
        choice[i] = pick;
    }
    groups.clear();
    for (size_t i = n; i > 0;) {
        const int s = choice[i];
        const size_t g = s < 0 ? 1 : (size_t)ss.g[s];
        groups.push_back(BsGroup{(uint32_t)(i - g), s < 0 ? -1 : ss.id[s]});
        i -= g;
    }
    std::reverse(groups.begin(), groups.end());
}

struct BsKey {
    uint64_t v;
    uint32_t i;
};

// Stable LSD radix sort by v over its 2k significant bits (the keys arrive in index order, so equal
// k-mers stay in index order): the comparison sort was most of the planning time at lim = 10 000.
void bs_sort_keys(std::vector<BsKey> &a, std::vector<BsKey> &tmp, int k) {
    const size_t n = a.size();
    tmp.resize(n);
    BsKey *src = a.data(), *dst = tmp.data();
    for (int shift = 0; shift < 2 * k; shift += 8) {
        size_t count[257] = {0};
        for (size_t i = 0; i < n; i++) count[((src[i].v >> shift) & 0xFFu) + 1]++;
        for (int b = 0; b < 256; b++) count[b + 1] += count[b];
        for (size_t i = 0; i < n; i++) dst[count[(src[i].v >> shift) & 0xFFu]++] = src[i];
        std::swap(src, dst);
    }
    if (src != a.data()) a.swap(tmp);
}

} // namespace

// Groups the query k-mers into units.  Each k-mer is looked at forwards and reversed (suffix sharing);
// a first cover of ALL k-mers in either direction tells which direction serves a k-mer better, then
// each direction's k-mers are covered on their own (bs_cover).  order[] receives the k-mer indices in scan order
// (units of shape 0, 1, ..., then singles), reversed[] whether the k-mer at that position is to be
// stored reversed, units[s] the number of units of shape s.  shape_mask: the shapes that may be used;
// alive: the expected share of text columns in which a unit's deep rows are computed (bs_unit_cost).
void bs_group_queries(const uint64_t *kmers, uint32_t n, int k, uint32_t shape_mask, float alive,
                      std::vector<uint32_t> &order, std::vector<uint8_t> &reversed, uint32_t (&units)[kBsShapes]) {
    order.resize(n);
    std::iota(order.begin(), order.end(), 0u);
    reversed.assign(n, 0);
    for (auto &u : units) u = 0;
    if (!shape_mask || k < 3 || n < 2) return;
    const BsShapeSet ss(k, shape_mask, alive);
    if (ss.n == 0) return;

    std::vector<BsKey> sorted[2], tmp;
    std::vector<float> share[2]; // by caller's index: the k-mer's share of its unit's cost when ALL k-mers go in direction d
    std::vector<BsGroup> groups;
    std::vector<uint64_t> vals(n);
    std::vector<float> f;
    std::vector<int8_t> choice;
    for (int d = 0; d < 2; d++) {
        sorted[d].resize(n);
        for (uint32_t i = 0; i < n; i++) sorted[d][i] = BsKey{d ? bs_reverse_kmer(kmers[i], k) : kmers[i], i};
        bs_sort_keys(sorted[d], tmp, k);
        for (uint32_t i = 0; i < n; i++) vals[i] = sorted[d][i].v;
        bs_cover(vals.data(), n, ss, f, choice, groups);
        share[d].resize(n);
        for (const BsGroup &grp : groups) {
            const int g = grp.shape < 0 ? 1 : bs_shape(k, grp.shape).g;
            const float c = (f[grp.first + g] - f[grp.first]) / (float)g;
            for (int j = 0; j < g; j++) share[d][sorted[d][grp.first + j].i] = c;
        }
    }

    // each k-mer goes to the direction in which it was cheaper, then each direction is covered on its own
    std::vector<uint64_t> sv[2];
    std::vector<uint32_t> si[2];
    std::vector<BsGroup> sg[2];
    for (int d = 0; d < 2; d++) {
        sv[d].reserve(n);
        si[d].reserve(n);
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t q = sorted[d][i].i;
            if ((share[1][q] < share[0][q] ? 1 : 0) == d) {
                sv[d].push_back(sorted[d][i].v);
                si[d].push_back(q);
            }
        }
        bs_cover(sv[d].data(), sv[d].size(), ss, f, choice, sg[d]);
    }

    // scan order: shape by shape, forward units then backward units, then the singles (stored forwards)
    uint32_t at = 0;
    std::vector<uint32_t> start(kBsShapes + 1, 0); // first position of each shape's members, singles last
    for (int d = 0; d < 2; d++)
        for (const BsGroup &grp : sg[d]) {
            if (grp.shape >= 0) {
                units[grp.shape]++;
                start[grp.shape] += (uint32_t)bs_shape(k, grp.shape).g;
            }
        }
    for (int s = 0; s <= kBsShapes; s++) {
        const uint32_t members = s < kBsShapes ? start[s] : 0;
        start[s] = at;
        at += members;
    }
    for (int d = 0; d < 2; d++)
        for (const BsGroup &grp : sg[d]) {
            const int s = grp.shape < 0 ? kBsShapes : grp.shape;
            const int g = grp.shape < 0 ? 1 : bs_shape(k, grp.shape).g;
            for (int j = 0; j < g; j++) {
                order[start[s]] = si[d][grp.first + j];
                reversed[start[s]] = (uint8_t)(grp.shape < 0 ? 0 : d);
                start[s]++;
            }
        }
}

// ---- dispatch ------------------------------------------------------------------------------------------
cudaError_t launch_bs_part0(BsLaunchCtx &l);
cudaError_t launch_bs_part1(BsLaunchCtx &l);
cudaError_t launch_bs_part2(BsLaunchCtx &l);
cudaError_t launch_bs_part3(BsLaunchCtx &l);

static cudaError_t bs_dispatch(BsLaunchCtx &l) {
    switch (l.c->k & 3) { // the instantiations are spread over four objects by k mod 4 (bitslice_part.cu)
    case 0: return launch_bs_part0(l);
    case 1: return launch_bs_part1(l);
    case 2: return launch_bs_part2(l);
    default: return launch_bs_part3(l);
    }
}

// Loads every scan kernel of this k (the unit shapes and the single-k-mer kernel) without launching anything.
cudaError_t warm_bs_kernels(const Ctx &c, int k) {
    if (k < 2 || k > 32) return cudaErrorInvalidValue;
    Ctx tmp = c; // only k is looked at
    tmp.k = (uint8_t)k;
    uint64_t launches = 0;
    BsLaunchCtx l;
    l.c = &tmp;
    l.r = BsRange{0, 0, 0, 0};
    l.d_counts = nullptr;
    l.sg_per_job_opt = 0;
    l.launches = &launches;
    l.slot = 0;
    l.warm = true;
    return bs_dispatch(l);
}

cudaError_t launch_bs_scan(const Ctx &c, uint64_t lo, uint64_t hi, unsigned long long *d_counts, uint32_t sg_per_job,
                           uint64_t *launches) {
    BsLaunchCtx l;
    l.c = &c;
    l.r.lo = lo;
    l.r.hi = hi;
    l.r.sg_first = (uint32_t)(lo / (32 * kGroupsPerSuper));
    l.r.n_sg = (uint32_t)((hi + 32 * kGroupsPerSuper - 1) / (32 * kGroupsPerSuper)) - l.r.sg_first;
    l.d_counts = d_counts;
    l.sg_per_job_opt = sg_per_job;
    l.launches = launches;
    l.slot = 0;
    *launches = 0;
    if (l.r.n_sg == 0 || c.n_kmers == 0) return cudaSuccess;
    if (c.k < 2 || c.k > 32) return cudaErrorInvalidValue;
    // the launches of the shapes run side by side: fork point for the side streams
    cudaError_t e = cudaEventRecord(c.bs_fork, c.stream);
    if (e != cudaSuccess) return e;
    e = bs_dispatch(l);
    // statistics (apc_scan_stats_read): what this scan costs on the ALU pipe according to its plan
    c.stat_scans++;
    c.stat_lop3_top += l.lop3_top;
    c.stat_lop3_all += l.lop3_all;
    c.stat_lop3_single += (double)c.n_kmers * l.r.n_sg * (2 * ((c.max_len + 1) / 2)) * (5 * c.k - 7);
    return e;
}

} // namespace apc
