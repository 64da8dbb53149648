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

// ---- families (family_core.cuh) ---------------------------------------------------------------------------
namespace {

inline int bs_lcp(uint64_t a, uint64_t b, int k) { // common prefix of two k-mers, in bases
    const uint64_t x = a ^ b;
    if (!x) return k;
    return k - (64 - __builtin_clzll(x) + 1) / 2;
}

struct FamUnitPlan {
    uint32_t first, count; // members [first, first + count) of the family's sorted member list
    int shape;
    float cost;            // row-equivalents per column
};

// rows of a family unit of shape s that are computed in every column / only while alive
inline void fam_unit_rows(int k, int s, int &top, int &deep) {
    const FamShape sh = fam_shape(s);
    const int a = k - sh.t - sh.st, m = bs_check_row_host(k);
    const int tops = std::max(0, std::min(sh.st, m - a)), topt = std::max(0, std::min(sh.t, m - a - sh.st));
    top = tops + sh.g * topt;
    deep = sh.st + sh.g * sh.t - top;
}

// Plans the families of one direction.  v[0..n) are the (oriented) k-mers of the direction in ascending order,
// idx[] their indices in the caller's order.  Members that end up in a family are flagged in taken[].
void fam_plan_direction(const uint64_t *v, const uint32_t *idx, size_t n, int k, bool reverse, float alive, FamPlan &out,
                        std::vector<uint8_t> &taken) {
    const int root = k / 2 - 2;               // a family shares this many leading bases
    const int a_min = k - fam_pub(k);          // shallowest trunk row a unit can attach below
    const int m = bs_check_row_host(k);
    const size_t kMinMembers = 12;
    const float single = alive * (float)k + (1.f - alive) * (float)std::min(m, k); // one k-mer scanned on its own
    const float leftover = 0.5f * single; // a member left to the unit kernels still shares rows there
    const int cap_rows = fam_state_rows(k);
    int shape_top[kFamShapes], shape_deep[kFamShapes];
    float shape_cost[kFamShapes];
    bool shape_ok[kFamShapes];
    for (int s = 0; s < kFamShapes; s++) {
        shape_ok[s] = fam_shape_valid(k, s);
        fam_unit_rows(k, s, shape_top[s], shape_deep[s]);
        // + what parking the state costs (6 LDS/STS per row and column block) and the dispatch of a unit
        shape_cost[s] = (float)shape_top[s] + alive * (float)shape_deep[s] + 0.15f * (float)(shape_top[s] + shape_deep[s]) + 1.0f;
    }
    std::vector<uint64_t> mem;
    std::vector<uint32_t> mem_idx;
    std::vector<int> lcp_trunk;
    std::vector<float> f;
    std::vector<FamUnitPlan> choice, units;
    for (size_t lo = 0; lo < n;) {
        size_t hi = lo + 1;
        while (hi < n && bs_lcp(v[lo], v[hi], k) >= root) hi++;
        const size_t fam_lo = lo, fam_hi = hi;
        lo = hi;
        if (fam_hi - fam_lo < kMinMembers) continue;
        // trunk = the member on the heavy path: at every depth follow the base most members continue with
        size_t tl = fam_lo, th = fam_hi;
        for (int d = root; d < k && th - tl > 1; d++) {
            const int sh = 2 * (k - 1 - d);
            size_t best_lo = tl, best_hi = tl, i = tl;
            while (i < th) {
                const uint64_t b = (v[i] >> sh) & 3u;
                size_t j = i + 1;
                while (j < th && ((v[j] >> sh) & 3u) == b) j++;
                if (j - i > best_hi - best_lo) { best_lo = i; best_hi = j; }
                i = j;
            }
            tl = best_lo;
            th = best_hi;
        }
        const size_t trunk_at = tl;
        const uint64_t trunk = v[trunk_at];
        mem.clear();
        mem_idx.clear();
        lcp_trunk.clear();
        for (size_t i = fam_lo; i < fam_hi; i++) {
            if (i == trunk_at) continue;
            mem.push_back(v[i]);
            mem_idx.push_back(idx[i]);
            lcp_trunk.push_back(std::min(bs_lcp(v[i], trunk, k), k - 1)); // a copy of the trunk attaches to its last row
        }
        // cheapest cover of the members (ascending order) by units and leftovers
        const size_t nm = mem.size();
        f.assign(nm + 1, 0.f);
        choice.assign(nm + 1, FamUnitPlan{0, 0, -1, 0.f});
        for (size_t i = 1; i <= nm; i++) {
            f[i] = f[i - 1] + leftover;
            choice[i] = FamUnitPlan{(uint32_t)(i - 1), 1, -1, leftover};
            for (int s = 0; s < kFamShapes; s++) {
                if (!shape_ok[s]) continue;
                const FamShape sh = fam_shape(s);
                const int a = k - sh.t - sh.st, p = k - sh.t;
                if (a < a_min || lcp_trunk[i - 1] < a) continue;
                // the longest run ending at member i-1 whose members share p bases with each other and a with the trunk
                size_t g = 1;
                while (g < (size_t)sh.g && g < i && lcp_trunk[i - 1 - g] >= a && bs_lcp(mem[i - 1 - g], mem[i - 1], k) >= p) g++;
                const float c = f[i - g] + shape_cost[s];
                if (c < f[i]) {
                    f[i] = c;
                    choice[i] = FamUnitPlan{(uint32_t)(i - g), (uint32_t)g, s, shape_cost[s]};
                }
            }
        }
        units.clear();
        for (size_t i = nm; i > 0;) {
            const FamUnitPlan u = choice[i];
            if (u.shape >= 0) units.push_back(u);
            i -= u.count;
        }
        int total_rows = 0;
        for (const FamUnitPlan &u : units) total_rows += shape_top[u.shape] + shape_deep[u.shape];
        // a CTA needs enough rows to keep its seven consumer warps busy
        if (units.size() < 4 || total_rows < 100) continue;
        // passes: as few as the parking area allows, units dealt out heaviest first
        const int n_pass = (total_rows + cap_rows - 49) / (cap_rows - 48);
        std::sort(units.begin(), units.end(), [](const FamUnitPlan &x, const FamUnitPlan &y) { return x.cost > y.cost; });
        std::vector<std::vector<FamUnitPlan>> pass_units((size_t)n_pass);
        std::vector<int> pass_rows((size_t)n_pass, 0);
        for (const FamUnitPlan &u : units) { // to the pass with the fewest rows so far
            const size_t best = (size_t)(std::min_element(pass_rows.begin(), pass_rows.end()) - pass_rows.begin());
            pass_units[best].push_back(u);
            pass_rows[best] += shape_top[u.shape] + shape_deep[u.shape];
        }
        taken[idx[trunk_at]] = 1;
        out.n_members++;
        for (size_t pi = 0; pi < pass_units.size(); pi++) {
            if (pass_units[pi].empty()) continue;
            FamPass pass{};
            pass.trunk = trunk;
            // the trunk's hits are counted by the family's first pass only
            pass.trunk_perm = (pi == 0 ? idx[trunk_at] : 0x7FFFFFFFu) | (reverse ? 0x80000000u : 0u);
            pass.unit_first = (uint32_t)out.units.size();
            // units to consumers: heaviest first to the least loaded warp
            float load[7] = {0, 0, 0, 0, 0, 0, 0};
            std::vector<std::vector<FamUnitPlan>> warp_units(7);
            for (const FamUnitPlan &u : pass_units[pi]) {
                int w = 0;
                for (int j = 1; j < 7; j++)
                    if (load[j] < load[w]) w = j;
                load[w] += u.cost;
                warp_units[(size_t)w].push_back(u);
            }
            uint32_t state_row = 0, n_in_pass = 0;
            for (int w = 0; w < 7; w++) {
                for (const FamUnitPlan &u : warp_units[(size_t)w]) {
                    const FamShape sh = fam_shape(u.shape);
                    FamUnit fu{};
                    fu.first_kmer = (uint32_t)out.kmers.size();
                    fu.shape = (uint16_t)u.shape;
                    fu.state_row = (uint16_t)state_row;
                    state_row += (uint32_t)(sh.st + sh.g * sh.t);
                    for (int g = 0; g < sh.g; g++) { // tails past the unit's members repeat the last member and count nothing
                        const size_t j = u.first + std::min<uint32_t>((uint32_t)g, u.count - 1);
                        out.kmers.push_back(mem[j]);
                        out.perm.push_back((uint32_t)g < u.count ? mem_idx[j] : 0xFFFFFFFFu);
                        if ((uint32_t)g < u.count) {
                            taken[mem_idx[j]] = 1;
                            out.n_members++;
                        }
                    }
                    out.units.push_back(fu);
                    n_in_pass++;
                    out.lop3_top_per_col += 5.0 * shape_top[u.shape];
                    out.lop3_all_per_col += 5.0 * (shape_top[u.shape] + shape_deep[u.shape]);
                }
                pass.warp_end[w] = n_in_pass;
            }
            out.lop3_top_per_col += 5.0 * k - 7.0; // the producer computes every trunk row in every column
            out.lop3_all_per_col += 5.0 * k - 7.0;
            out.passes.push_back(pass);
        }
    }
}

} // namespace

// Groups the query k-mers into units.  Each k-mer is looked at forwards and reversed (suffix sharing);
// a first cover of ALL k-mers in either direction tells which direction serves a k-mer better, then
// each direction's k-mers are covered on their own (bs_cover).  order[] receives the k-mer indices in scan order
// (units of shape 0, 1, ..., then singles), reversed[] whether the k-mer at that position is to be
// stored reversed, units[s] the number of units of shape s.  shape_mask: the shapes that may be used;
// alive: the expected share of text columns in which a unit's deep rows are computed (bs_unit_cost).
void bs_group_queries(const uint64_t *kmers, uint32_t n, int k, uint32_t shape_mask, float alive,
                      std::vector<uint32_t> &order, std::vector<uint8_t> &reversed, uint32_t (&units)[kBsShapes],
                      FamPlan *families) {
    order.resize(n);
    std::iota(order.begin(), order.end(), 0u);
    reversed.assign(n, 0);
    for (auto &u : units) u = 0;
    if (families) *families = FamPlan{};
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

    // each k-mer goes to the direction in which it was cheaper, then each direction is covered on its own:
    // first by families (one CTA per family and 1024 reads, family_core.cuh), the rest by one-warp units
    std::vector<uint64_t> sv[2];
    std::vector<uint32_t> si[2];
    std::vector<BsGroup> sg[2];
    std::vector<uint8_t> taken(n, 0);
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
        if (families && k >= kFamMinK) {
            fam_plan_direction(sv[d].data(), si[d].data(), sv[d].size(), k, d != 0, alive, *families, taken);
            size_t w = 0;
            for (size_t i = 0; i < sv[d].size(); i++)
                if (!taken[si[d][i]]) {
                    sv[d][w] = sv[d][i];
                    si[d][w] = si[d][i];
                    w++;
                }
            sv[d].resize(w);
            si[d].resize(w);
        }
        bs_cover(sv[d].data(), sv[d].size(), ss, f, choice, sg[d]);
    }
    // the unit kernels see only the k-mers no family took
    const uint32_t n_rest = (uint32_t)(sv[0].size() + sv[1].size());
    order.resize(n_rest);
    reversed.assign(n_rest, 0);

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
    switch (c.k & 3) { // the instantiations are spread over four objects by k mod 4 (bitslice_part.cu)
    case 0: e = launch_bs_part0(l); break;
    case 1: e = launch_bs_part1(l); break;
    case 2: e = launch_bs_part2(l); break;
    default: e = launch_bs_part3(l); break;
    }
    // statistics (apc_scan_stats_read): what this scan costs on the ALU pipe according to its plan
    c.stat_scans++;
    c.stat_lop3_top += l.lop3_top;
    c.stat_lop3_all += l.lop3_all;
    c.stat_lop3_single += (double)c.n_kmers * l.r.n_sg * (2 * ((c.max_len + 1) / 2)) * (5 * c.k - 7);
    return e;
}

} // namespace apc
