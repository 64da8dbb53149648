// exact_kernels.cu — K3..K5: exact k-mer count, low-complexity filter and
// top-N selection on the device.
//
// Replaces count_kmers (/root/reference/approx_counter.cpp:487-519: slide a
// k-window over every sampled read, skip windows with N (:313-321), 2-bit pack
// (:55-62), drop low-complexity (:214-234) and forbidden (:330-332) k-mers,
// count[n] += 1 in an unordered_map) and get_most_frequent (:396-405: full
// std::sort by CompareCount :275-305, keep the first `limit`) /
// get_solid_kmers (:372-388).
//
// Device plan (all HBM-bound integer work, no tensor cores):
//   K3 extract_keys   one lane per read walks the scan tiles (the same resident
//                     layout K1 reads), keeps the rolling 2-bit window, the
//                     length of the current N-free run and the rolling dimer
//                     score; surviving windows are staged per warp in shared
//                     memory and flushed with one atomic per 16 columns, so
//                     the key stream is written coalesced and compact.
//   K4 sort + RLE     LSD radix sort of the keys over 2k bits (cub, 32-bit keys
//                     for k<=16 else 64-bit) then a hand-written run-length
//                     pass (heads per block -> scan -> unique keys + run
//                     starts).  count(i) = start[i+1]-start[i].
//   K5 select         the first `lim` entries of the CompareCount order are
//                     found WITHOUT sorting the D distinct k-mers: radix-select
//                     the lim-th largest count c*, then the dimer-sum threshold
//                     s* inside the tie class count==c*, then the k-mer value
//                     boundary inside (c*, s*) — the unique keys are already
//                     ascending, so "largest k-mers first" is "last by index".
//                     Only the <= lim survivors go to the host, where the
//                     reference comparator itself (float getComplexity) orders
//                     them.
//
// Float semantics of the filter (:227-233): s = sum / float(2*(k-2)) with an
// integer sum <= 930, one correctly rounded fp32 division, compared `>=` with
// the adjusted threshold.  The quotient is monotone in the integer numerator,
// so the device compares integers against the smallest filtered sum, which the
// host computes with exactly that fp32 expression (apch::lc_min_filtered_sum).
// The comparator's `a_comp < b_comp` (:283-301) is likewise the order of the
// integer sums (same denominator, distinct sums give distinct floats).
#include <algorithm>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "apc_internal.h"
#include "host/host_util.h"

namespace apc {

namespace {

constexpr int kExtractWarps = 8;
constexpr int kPassThreads = 256;
constexpr int kPassItems = 8; // consecutive items per thread in the streaming passes
constexpr int kPassTile = kPassThreads * kPassItems;

// counters[] slots
enum { CNT_KEYS = 0, CNT_HAD_N = 1, CNT_SELECTED = 2, CNT_SLOTS = 8 };

// ---- dimer score -------------------------------------------------------------------
// sum over the 16 dimer bins of v*(v-1) for the k-1 overlapping dimers of a
// k-mer (:216-231).  Direct form, used on the distinct k-mers in K5.
__device__ __forceinline__ uint32_t dimer_sum_direct(uint64_t kmer, int k) {
    // E[c]: bit 2i set iff base i (2-bit field at bit 2i) equals c.  The dimer read at
    // field i (:219 `kmer & 15` after i shifts) is (base i+1, base i) = hi*4 + lo, so its
    // count is popc(E[lo] & (E[hi] >> 2)) over the k-1 dimer positions.
    const uint64_t lo_bits = 0x5555555555555555ull;
    const uint64_t b0 = kmer & lo_bits, b1 = (kmer >> 1) & lo_bits;
    const uint64_t valid = k >= 33 ? lo_bits : (lo_bits & ((1ull << (2 * (k - 1))) - 1ull)); // positions 0..k-2
    const uint64_t e[4] = {~b1 & ~b0 & lo_bits, ~b1 & b0, b1 & ~b0 & lo_bits, b1 & b0};
    uint32_t sum = 0;
#pragma unroll
    for (int hi = 0; hi < 4; hi++) {
        const uint64_t eh = (e[hi] >> 2) & valid;
#pragma unroll
        for (int lo = 0; lo < 4; lo++) {
            const uint32_t v = (uint32_t)__popcll(e[lo] & eh);
            sum += v * (v - 1u);
        }
    }
    return sum;
}

// Rolling form for K3: bins of the current k-window, updated by one leaving and
// one entering dimer per base.
struct DimerRoll {
    unsigned long long lo, hi;
    uint32_t sum;
    __device__ __forceinline__ void init(int k) { // window of k 'A's: bin 0 holds k-1
        lo = (unsigned long long)(k - 1);
        hi = 0;
        sum = (uint32_t)((k - 1) * (k - 2));
    }
    __device__ __forceinline__ void remove(uint32_t idx) {
        const uint32_t sh = (idx & 7u) * 8u;
        const unsigned long long w = (idx & 8u) ? hi : lo;
        const uint32_t c = (uint32_t)((w >> sh) & 0xFFu);
        sum -= 2u * (c - 1u);
        const unsigned long long dec = 1ull << sh;
        if (idx & 8u) hi -= dec; else lo -= dec;
    }
    __device__ __forceinline__ void add(uint32_t idx) {
        const uint32_t sh = (idx & 7u) * 8u;
        const unsigned long long w = (idx & 8u) ? hi : lo;
        const uint32_t c = (uint32_t)((w >> sh) & 0xFFu);
        sum += 2u * c;
        const unsigned long long inc = 1ull << sh;
        if (idx & 8u) hi += inc; else lo += inc;
    }
};

__device__ __forceinline__ bool is_forbidden(const uint64_t *__restrict__ forb, uint32_t n, uint64_t key) {
    uint32_t lo = 0, hi = n; // sorted ascending (std::set order, :44)
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint64_t v = forb[mid];
        if (v == key) return true;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return false;
}

// ---- K3 ---------------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kExtractWarps * 32)
extract_keys_kernel(const uint4 *__restrict__ tiles, const uint32_t n_tiles, const uint32_t chunks,
                    const int k, const uint32_t lc_min_sum, const uint64_t *__restrict__ forb,
                    const uint32_t n_forb, K *__restrict__ keys, unsigned long long *__restrict__ counters) {
    __shared__ K s_buf[kExtractWarps][kTileReads * kChunkBases];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const uint32_t top_shift = 2u * (uint32_t)(k - 2);

    for (uint32_t tile = blockIdx.x * kExtractWarps + warp; tile < n_tiles; tile += gridDim.x * kExtractWarps) {
        const uint4 *p = tiles + (size_t)tile * chunks * kTileReads + lane;
        uint64_t kmer = 0;
        uint32_t run = 0; // consecutive non-N bases ending here (saturates at k)
        DimerRoll roll;
        roll.init(k);
        for (uint32_t ch = 0; ch < chunks; ch++) {
            const uint4 v = __ldg(p + (size_t)ch * kTileReads);
            const uint32_t tw[4] = {v.x, v.y, v.z, v.w};
            uint32_t fill = 0; // keys staged by this warp for this chunk (warp-uniform)
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t code = (tw[wi] >> (8 * j)) & 0xFFu; // 0x00,0x10,0x20,0x30 | 0x40 = N/pad
                    const bool is_n = code >= kCodeN;
                    const uint32_t b = is_n ? 0u : (code >> 4);
                    roll.remove((uint32_t)(kmer >> top_shift) & 15u);
                    kmer = ((kmer << 2) | b) & kmask;
                    roll.add((uint32_t)kmer & 15u);
                    run = is_n ? 0u : min(run + 1u, (uint32_t)k);
                    // a window ends here; the reference skips it when it holds an N
                    // (:313-321).  Padding is N-coded, so a window running past the
                    // read's end never has run == k.
                    bool emit = run == (uint32_t)k;
                    if (emit && roll.sum >= lc_min_sum) emit = false;
                    if (emit && n_forb) emit = !is_forbidden(forb, n_forb, kmer);
                    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, emit);
                    if (emit) s_buf[warp][fill + __popc(ballot & lt_mask)] = (K)kmer;
                    fill += __popc(ballot);
                }
            }
            if (fill) { // warp-uniform
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&counters[CNT_KEYS], (unsigned long long)fill);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                __syncwarp();
                for (uint32_t i = lane; i < fill; i += 32) keys[base + i] = s_buf[warp][i];
                __syncwarp();
            }
        }
    }
}

// windows skipped because they hold an N (:506): per read, the windows inside the
// read minus those that are N-free.  One lane per read, same walk as K3 but only the
// run length is tracked.
__global__ void __launch_bounds__(kExtractWarps * 32)
count_n_windows_kernel(const uint4 *__restrict__ tiles, const uint32_t n_tiles, const uint32_t chunks,
                       const uint32_t *__restrict__ lens, const int k,
                       unsigned long long *__restrict__ counters) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long had_n = 0;
    for (uint32_t tile = blockIdx.x * kExtractWarps + warp; tile < n_tiles; tile += gridDim.x * kExtractWarps) {
        const uint4 *p = tiles + (size_t)tile * chunks * kTileReads + lane;
        const uint32_t len = lens[(size_t)tile * kTileReads + lane];
        uint32_t run = 0, clean = 0;
        for (uint32_t ch = 0; ch < chunks; ch++) {
            const uint4 v = __ldg(p + (size_t)ch * kTileReads);
            const uint32_t tw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool is_n = ((tw[wi] >> (8 * j)) & 0xFFu) >= kCodeN;
                    run = is_n ? 0u : min(run + 1u, (uint32_t)k);
                    clean += run == (uint32_t)k;
                }
            }
        }
        const uint32_t windows = len >= (uint32_t)k ? len - (uint32_t)k + 1u : 0u;
        had_n += windows - clean;
    }
    for (int o = 16; o; o >>= 1) had_n += __shfl_down_sync(0xFFFFFFFFu, had_n, o);
    if (lane == 0 && had_n) atomicAdd(&counters[CNT_HAD_N], had_n);
}

// ---- K4: run-length pass over the sorted keys ------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kPassThreads)
count_heads_kernel(const K *__restrict__ sorted, const uint64_t n, uint32_t *__restrict__ block_heads) {
    const uint64_t base = (uint64_t)blockIdx.x * kPassTile + (uint64_t)threadIdx.x * kPassItems;
    uint32_t heads = 0;
    if (base < n) {
        K prev = base ? sorted[base - 1] : (K)0;
#pragma unroll
        for (int i = 0; i < kPassItems; i++) {
            const uint64_t idx = base + i;
            if (idx < n) {
                const K cur = sorted[idx];
                heads += (idx == 0) || (cur != prev);
                prev = cur;
            }
        }
    }
    typedef cub::BlockReduce<uint32_t, kPassThreads> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const uint32_t total = Reduce(tmp).Sum(heads);
    if (threadIdx.x == 0) block_heads[blockIdx.x] = total;
}

template <typename K>
__global__ void __launch_bounds__(kPassThreads)
write_runs_kernel(const K *__restrict__ sorted, const uint64_t n, const uint32_t *__restrict__ block_offs,
                  K *__restrict__ uniq, uint32_t *__restrict__ start) {
    const uint64_t base = (uint64_t)blockIdx.x * kPassTile + (uint64_t)threadIdx.x * kPassItems;
    K cur[kPassItems];
    bool head[kPassItems];
    uint32_t heads = 0;
    K prev = (base && base < n) ? sorted[base - 1] : (K)0;
#pragma unroll
    for (int i = 0; i < kPassItems; i++) {
        const uint64_t idx = base + i;
        head[i] = false;
        cur[i] = (K)0;
        if (idx < n) {
            cur[i] = sorted[idx];
            head[i] = (idx == 0) || (cur[i] != prev);
            prev = cur[i];
            heads += head[i];
        }
    }
    typedef cub::BlockScan<uint32_t, kPassThreads> Scan;
    __shared__ typename Scan::TempStorage tmp;
    uint32_t rank;
    Scan(tmp).ExclusiveSum(heads, rank);
    rank += block_offs[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kPassItems; i++) {
        if (head[i]) {
            uniq[rank] = cur[i];
            start[rank] = (uint32_t)(base + i);
            rank++;
        }
    }
}

// ---- K5 -------------------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kPassThreads)
dimer_sums_kernel(const K *__restrict__ uniq, const uint64_t d, const int k, uint16_t *__restrict__ dsum) {
    const uint64_t i = (uint64_t)blockIdx.x * kPassThreads + threadIdx.x;
    if (i < d) dsum[i] = (uint16_t)dimer_sum_direct((uint64_t)uniq[i], k);
}

// histogram of one 8-bit digit of count(i) over the entries whose higher digits match
// (shift == 0xFFFFFFFF: histogram of min(count, 255) over all entries)
__global__ void __launch_bounds__(kPassThreads)
count_digit_hist_kernel(const uint32_t *__restrict__ start, const uint64_t d, const uint32_t prefix_mask,
                        const uint32_t prefix_val, const uint32_t shift, unsigned long long *__restrict__ hist) {
    __shared__ uint32_t s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kPassTile + (uint64_t)threadIdx.x * kPassItems;
    uint32_t bin = 0xFFFFFFFFu, runlen = 0; // counts repeat a lot: flush once per run
    if (base < d) {
        uint32_t s0 = start[base];
#pragma unroll
        for (int i = 0; i < kPassItems; i++) {
            const uint64_t idx = base + i;
            if (idx < d) {
                const uint32_t s1 = start[idx + 1];
                const uint32_t c = s1 - s0;
                s0 = s1;
                if ((c & prefix_mask) == prefix_val) {
                    const uint32_t b = shift == 0xFFFFFFFFu ? min(c, 255u) : (c >> shift) & 0xFFu;
                    if (b != bin) {
                        if (runlen) atomicAdd(&s_hist[bin], runlen);
                        bin = b;
                        runlen = 0;
                    }
                    runlen++;
                }
            }
        }
    }
    if (runlen) atomicAdd(&s_hist[bin], runlen);
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

// histogram of dsum/2 (0..511) over the tie class count == c_star
__global__ void __launch_bounds__(kPassThreads)
dsum_hist_kernel(const uint32_t *__restrict__ start, const uint16_t *__restrict__ dsum, const uint64_t d,
                 const uint32_t c_star, unsigned long long *__restrict__ hist) {
    __shared__ uint32_t s_hist[512];
    s_hist[threadIdx.x] = 0;
    s_hist[threadIdx.x + 256] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kPassTile + (uint64_t)threadIdx.x * kPassItems;
    if (base < d) {
        uint32_t s0 = start[base];
#pragma unroll
        for (int i = 0; i < kPassItems; i++) {
            const uint64_t idx = base + i;
            if (idx < d) {
                const uint32_t s1 = start[idx + 1];
                if (s1 - s0 == c_star) atomicAdd(&s_hist[(dsum[idx] >> 1) & 511u], 1u);
                s0 = s1;
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 512; b += kPassThreads)
        if (s_hist[b]) atomicAdd(&hist[b], (unsigned long long)s_hist[b]);
}

// per-block population of the innermost tie class (count == c_star, dsum == s_star)
__global__ void __launch_bounds__(kPassThreads)
tie_block_counts_kernel(const uint32_t *__restrict__ start, const uint16_t *__restrict__ dsum, const uint64_t d,
                        const uint32_t c_star, const uint32_t s_star, uint32_t *__restrict__ block_cnt) {
    const uint64_t base = (uint64_t)blockIdx.x * kPassTile + (uint64_t)threadIdx.x * kPassItems;
    uint32_t n = 0;
    if (base < d) {
        uint32_t s0 = start[base];
#pragma unroll
        for (int i = 0; i < kPassItems; i++) {
            const uint64_t idx = base + i;
            if (idx < d) {
                const uint32_t s1 = start[idx + 1];
                n += (s1 - s0 == c_star) && (dsum[idx] == s_star);
                s0 = s1;
            }
        }
    }
    typedef cub::BlockReduce<uint32_t, kPassThreads> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const uint32_t total = Reduce(tmp).Sum(n);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = total;
}

struct SelectRule {
    uint32_t c_min;       // take every entry with count >= c_min ...
    uint32_t c_star;      // ... and inside count == c_star (c_star = c_min - 1):
    uint32_t s_star;      //     dsum < s_star: take; dsum == s_star: by position
    uint32_t use_tie;     // 0: only the c_min rule applies
    uint32_t tie_block;   // blocks > tie_block take all of the innermost class,
    uint32_t tie_take;    // block == tie_block takes its last tie_take members
    uint32_t capacity;    // entries the output arrays can hold
};

template <typename K>
__global__ void __launch_bounds__(kPassThreads)
select_kernel(const K *__restrict__ uniq, const uint32_t *__restrict__ start, const uint16_t *__restrict__ dsum,
              const uint64_t d, const SelectRule rule, uint64_t *__restrict__ out_kmers,
              uint64_t *__restrict__ out_counts, unsigned long long *__restrict__ counters) {
    const uint64_t base = (uint64_t)blockIdx.x * kPassTile + (uint64_t)threadIdx.x * kPassItems;
    uint32_t cnt[kPassItems];
    bool take[kPassItems], inner[kPassItems];
    uint32_t n_inner = 0;
    uint32_t s0 = base < d ? start[base] : 0;
#pragma unroll
    for (int i = 0; i < kPassItems; i++) {
        const uint64_t idx = base + i;
        take[i] = inner[i] = false;
        cnt[i] = 0;
        if (idx < d) {
            const uint32_t s1 = start[idx + 1];
            cnt[i] = s1 - s0;
            s0 = s1;
            if (cnt[i] >= rule.c_min) take[i] = true;
            else if (rule.use_tie && cnt[i] == rule.c_star) {
                const uint32_t s = dsum[idx];
                if (s < rule.s_star) take[i] = true;
                else if (s == rule.s_star) {
                    inner[i] = true;
                    n_inner++;
                }
            }
        }
    }
    if (rule.use_tie) {
        if (blockIdx.x > rule.tie_block) {
#pragma unroll
            for (int i = 0; i < kPassItems; i++) take[i] = take[i] || inner[i];
        } else if (blockIdx.x == rule.tie_block) { // block-uniform branch
            typedef cub::BlockScan<uint32_t, kPassThreads> Scan;
            __shared__ typename Scan::TempStorage tmp;
            uint32_t rank, total;
            Scan(tmp).ExclusiveSum(n_inner, rank, total);
            const uint32_t first_taken = total - rule.tie_take; // ranks >= this are the largest k-mers
#pragma unroll
            for (int i = 0; i < kPassItems; i++) {
                if (inner[i]) {
                    if (rank >= first_taken) take[i] = true;
                    rank++;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kPassItems; i++) {
        if (take[i]) {
            const unsigned long long slot = atomicAdd(&counters[CNT_SELECTED], 1ull);
            if (slot < rule.capacity) {
                out_kmers[slot] = (uint64_t)uniq[base + i];
                out_counts[slot] = cnt[i];
            }
        }
    }
}

// ---- host orchestration -----------------------------------------------------------------------
template <typename T>
int grow_dev(Ctx *c, T *&ptr, size_t &cap, size_t need_bytes) {
    if (need_bytes <= cap && ptr) return APC_OK;
    if (ptr) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) e = cudaFree(ptr);
        if (e != cudaSuccess) return fail(c, APC_ERR_CUDA, "cudaFree", e);
        ptr = nullptr;
        cap = 0;
    }
    if (need_bytes == 0) need_bytes = 16;
    // 1/8 of slack: the `end` sample of a run is one base per read longer than the `start`
    // sample (:463) and must not force a re-allocation of every buffer
    need_bytes += need_bytes / 8;
    cudaError_t e = cudaMalloc((void **)&ptr, need_bytes);
    if (e != cudaSuccess) {
        ptr = nullptr;
        cudaGetLastError();
        return fail(c, e == cudaErrorMemoryAllocation ? APC_ERR_NOMEM : APC_ERR_CUDA, "cudaMalloc (exact stage)", e);
    }
    cap = need_bytes;
    return APC_OK;
}

inline unsigned blocks_for(uint64_t n, uint64_t per_block) { return (unsigned)((n + per_block - 1) / per_block); }

template <typename K>
int run_exact(Ctx *c, int k, uint32_t lc_min_sum, uint64_t lim, uint64_t solid_km, const uint64_t *forbidden,
              uint64_t n_forbidden, uint64_t capacity, std::vector<uint64_t> &kmers, std::vector<uint64_t> &counts,
              uint64_t *n_needed, uint64_t *n_distinct, uint64_t *n_had_n) {
    ExactScratch &x = c->exact;
    cudaStream_t s = c->stream;
    int st;
    kmers.clear();
    counts.clear();
    x.launches = 0;

    // upper bound of the key stream: every window of every read
    const uint64_t max_windows = c->total_windows((uint32_t)k);
    if (max_windows >= 0xFFFFFFFFull) return fail(c, APC_ERR_INVALID, "exact stage: more than 2^32-2 windows");

    if ((st = grow_dev(c, x.d_counters, x.counters_cap, CNT_SLOTS * sizeof(unsigned long long)))) return st;
    if ((st = grow_dev(c, x.d_hist, x.hist_cap, 512 * sizeof(unsigned long long)))) return st;
    APC_CUDA(c, cudaMemsetAsync(x.d_counters, 0, CNT_SLOTS * sizeof(unsigned long long), s));

    // forbidden set, sorted like the std::set (:44)
    uint32_t n_forb = 0;
    if (n_forbidden) {
        std::vector<uint64_t> f(forbidden, forbidden + n_forbidden);
        std::sort(f.begin(), f.end());
        f.erase(std::unique(f.begin(), f.end()), f.end());
        if (f.size() > 0x7FFFFFFFull) return fail(c, APC_ERR_INVALID, "too many forbidden k-mers");
        n_forb = (uint32_t)f.size();
        if ((st = grow_dev(c, x.d_forb, x.forb_cap, f.size() * sizeof(uint64_t)))) return st;
        APC_CUDA(c, cudaMemcpyAsync(x.d_forb, f.data(), f.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        APC_CUDA(c, cudaStreamSynchronize(s)); // f is a local
    }

    unsigned long long h_counters[CNT_SLOTS] = {0};
    uint64_t n_keys = 0;
    if (max_windows && c->n_tiles) {
        if ((st = grow_dev(c, x.d_keys_a, x.keys_a_cap, max_windows * sizeof(K)))) return st;
        const unsigned grid = std::min<unsigned>(blocks_for(c->n_tiles, kExtractWarps), (unsigned)c->sm_count * 8u);
        extract_keys_kernel<K><<<grid, kExtractWarps * 32, 0, s>>>(
            c->d_tiles, c->n_tiles, c->chunks, k, lc_min_sum, x.d_forb, n_forb, (K *)x.d_keys_a, x.d_counters);
        APC_CUDA(c, cudaGetLastError());
        count_n_windows_kernel<<<grid, kExtractWarps * 32, 0, s>>>(c->d_tiles, c->n_tiles, c->chunks, c->d_lens, k,
                                                                    x.d_counters);
        APC_CUDA(c, cudaGetLastError());
        x.launches += 2;
        APC_CUDA(c, cudaMemcpyAsync(h_counters, x.d_counters, sizeof h_counters, cudaMemcpyDeviceToHost, s));
        APC_CUDA(c, cudaStreamSynchronize(s));
        n_keys = h_counters[CNT_KEYS];
    }
    if (n_had_n) *n_had_n = h_counters[CNT_HAD_N];
    if (n_distinct) *n_distinct = 0;
    if (n_needed) *n_needed = 0;
    if (n_keys == 0) return APC_OK;

    // ---- K4: sort, then unique keys + run starts
    if ((st = grow_dev(c, x.d_keys_b, x.keys_b_cap, n_keys * sizeof(K)))) return st;
    size_t temp_bytes = 0;
    APC_CUDA(c, cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const K *)x.d_keys_a, (K *)x.d_keys_b,
                                               (unsigned long long)n_keys, 0, 2 * k, s));
    const unsigned n_pass_blocks = blocks_for(n_keys, kPassTile);
    size_t scan_bytes = 0;
    APC_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                              (int)n_pass_blocks, s));
    if ((st = grow_dev(c, x.d_temp, x.temp_cap, std::max(temp_bytes, scan_bytes)))) return st;
    APC_CUDA(c, cub::DeviceRadixSort::SortKeys(x.d_temp, temp_bytes, (const K *)x.d_keys_a, (K *)x.d_keys_b,
                                               (unsigned long long)n_keys, 0, 2 * k, s));
    x.launches += 1; // counted as one library call
    const K *sorted = (const K *)x.d_keys_b;

    if ((st = grow_dev(c, x.d_block_a, x.block_a_cap, (size_t)n_pass_blocks * sizeof(uint32_t)))) return st;
    if ((st = grow_dev(c, x.d_block_b, x.block_b_cap, (size_t)n_pass_blocks * sizeof(uint32_t)))) return st;
    count_heads_kernel<K><<<n_pass_blocks, kPassThreads, 0, s>>>(sorted, n_keys, x.d_block_a);
    APC_CUDA(c, cudaGetLastError());
    APC_CUDA(c, cub::DeviceScan::ExclusiveSum(x.d_temp, scan_bytes, x.d_block_a, x.d_block_b, (int)n_pass_blocks, s));
    uint32_t last_off = 0, last_heads = 0;
    APC_CUDA(c, cudaMemcpyAsync(&last_off, x.d_block_b + (n_pass_blocks - 1), sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, s));
    APC_CUDA(c, cudaMemcpyAsync(&last_heads, x.d_block_a + (n_pass_blocks - 1), sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, s));
    APC_CUDA(c, cudaStreamSynchronize(s));
    const uint64_t d = (uint64_t)last_off + last_heads; // distinct k-mers = count.size() (:883)
    if (n_distinct) *n_distinct = d;

    K *uniq = (K *)x.d_keys_a; // the unsorted stream is dead: reuse its buffer
    if ((st = grow_dev(c, x.d_start, x.start_cap, (d + 1) * sizeof(uint32_t)))) return st;
    write_runs_kernel<K><<<n_pass_blocks, kPassThreads, 0, s>>>(sorted, n_keys, x.d_block_b, uniq, x.d_start);
    APC_CUDA(c, cudaGetLastError());
    const uint32_t n_keys32 = (uint32_t)n_keys;
    APC_CUDA(c, cudaMemcpyAsync(x.d_start + d, &n_keys32, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    x.launches += 3;

    // ---- K5: which entries survive
    const unsigned d_blocks = blocks_for(d, kPassTile);
    SelectRule rule{};
    uint64_t n_take = 0;
    unsigned long long h_hist[512];
    auto read_hist = [&](int bins) -> int {
        APC_CUDA(c, cudaMemcpyAsync(h_hist, x.d_hist, bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        APC_CUDA(c, cudaStreamSynchronize(s));
        return APC_OK;
    };

    if (solid_km) { // :372-388 — everything with count >= solid_km
        rule.c_min = solid_km > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)solid_km;
        rule.use_tie = 0;
        n_take = ~0ull; // unknown until counted
    } else if (d <= lim) {
        rule.c_min = 0;
        rule.use_tie = 0;
        n_take = d;
    } else if (lim == 0) {
        return APC_OK;
    } else {
        // lim-th largest count by MSB-first radix select over count(i)
        uint64_t need = lim; // rank still to be located inside the current prefix class
        uint32_t prefix_mask = 0, prefix_val = 0;
        uint64_t class_size = d;
        // Fast path: counts are small for almost every k-mer.  One histogram of min(count, 255)
        // settles the threshold whenever fewer than `lim` k-mers have a count >= 255.
        bool settled = false;
        {
            APC_CUDA(c, cudaMemsetAsync(x.d_hist, 0, 256 * sizeof(unsigned long long), s));
            count_digit_hist_kernel<<<d_blocks, kPassThreads, 0, s>>>(x.d_start, d, 0u, 0u, 0xFFFFFFFFu, x.d_hist);
            APC_CUDA(c, cudaGetLastError());
            x.launches++;
            if ((st = read_hist(256))) return st;
            if (h_hist[255] < need) {
                int digit = 254;
                uint64_t above = h_hist[255];
                for (; digit > 0; digit--) {
                    if (above + h_hist[digit] >= need) break;
                    above += h_hist[digit];
                }
                need -= above;
                class_size = h_hist[digit];
                prefix_val = (uint32_t)digit;
                settled = true;
            }
        }
        for (int shift = settled ? -8 : 24; shift >= 0; shift -= 8) {
            APC_CUDA(c, cudaMemsetAsync(x.d_hist, 0, 256 * sizeof(unsigned long long), s));
            count_digit_hist_kernel<<<d_blocks, kPassThreads, 0, s>>>(x.d_start, d, prefix_mask, prefix_val,
                                                                      (uint32_t)shift, x.d_hist);
            APC_CUDA(c, cudaGetLastError());
            x.launches++;
            if ((st = read_hist(256))) return st;
            int digit = 255;
            uint64_t above = 0;
            for (; digit > 0; digit--) {
                if (above + h_hist[digit] >= need) break;
                above += h_hist[digit];
            }
            need -= above;
            class_size = h_hist[digit];
            prefix_mask |= 0xFFu << shift;
            prefix_val |= (uint32_t)digit << shift;
        }
        const uint32_t c_star = prefix_val; // `need` of the class_size entries with count == c_star are kept
        rule.c_min = c_star + 1;
        rule.c_star = c_star;
        n_take = lim;
        if (need == class_size) {
            rule.c_min = c_star;
            rule.use_tie = 0;
        } else {
            rule.use_tie = 1;
            if ((st = grow_dev(c, x.d_dsum, x.dsum_cap, d * sizeof(uint16_t)))) return st;
            dimer_sums_kernel<K><<<blocks_for(d, kPassThreads), kPassThreads, 0, s>>>(uniq, d, k, x.d_dsum);
            APC_CUDA(c, cudaGetLastError());
            APC_CUDA(c, cudaMemsetAsync(x.d_hist, 0, 512 * sizeof(unsigned long long), s));
            dsum_hist_kernel<<<d_blocks, kPassThreads, 0, s>>>(x.d_start, x.d_dsum, d, c_star, x.d_hist);
            APC_CUDA(c, cudaGetLastError());
            x.launches += 2;
            if ((st = read_hist(512))) return st;
            int bin = 0;
            uint64_t below = 0;
            for (; bin < 511; bin++) { // complexity ascending (:301)
                if (below + h_hist[bin] >= need) break;
                below += h_hist[bin];
            }
            need -= below;
            rule.s_star = (uint32_t)bin * 2u;
            const uint64_t inner = h_hist[bin];
            if (need == inner) { // the whole innermost class fits
                rule.s_star += 2; // "dsum < s_star" now covers it
                rule.tie_block = 0xFFFFFFFFu;
                rule.tie_take = 0;
            } else {
                // largest k-mer values first (:297) = last by index in the ascending unique keys
                tie_block_counts_kernel<<<d_blocks, kPassThreads, 0, s>>>(x.d_start, x.d_dsum, d, c_star, rule.s_star,
                                                                          x.d_block_a);
                APC_CUDA(c, cudaGetLastError());
                x.launches++;
                std::vector<uint32_t> bc(d_blocks);
                APC_CUDA(c, cudaMemcpyAsync(bc.data(), x.d_block_a, (size_t)d_blocks * sizeof(uint32_t),
                                            cudaMemcpyDeviceToHost, s));
                APC_CUDA(c, cudaStreamSynchronize(s));
                uint64_t left = need;
                uint32_t b = d_blocks;
                while (b > 0 && bc[b - 1] <= left) { // whole trailing blocks
                    left -= bc[b - 1];
                    b--;
                }
                if (left == 0) { // boundary falls between blocks b-1 and b
                    rule.tie_block = b - 1; // (b >= 1 because need < inner)
                    rule.tie_take = 0;
                } else {
                    rule.tie_block = b - 1;
                    rule.tie_take = (uint32_t)left;
                }
            }
        }
    }

    // ---- gather the survivors
    uint64_t out_cap = solid_km ? capacity : n_take;
    if (out_cap > d) out_cap = d;
    if (out_cap > 0xFFFFFFFFull) out_cap = 0xFFFFFFFFull;
    rule.capacity = (uint32_t)out_cap;
    if ((st = grow_dev(c, x.d_sel_k, x.sel_k_cap, std::max<uint64_t>(out_cap, 1) * sizeof(uint64_t)))) return st;
    if ((st = grow_dev(c, x.d_sel_c, x.sel_c_cap, std::max<uint64_t>(out_cap, 1) * sizeof(uint64_t)))) return st;
    select_kernel<K><<<d_blocks, kPassThreads, 0, s>>>(uniq, x.d_start, x.d_dsum, d, rule, x.d_sel_k, x.d_sel_c,
                                                       x.d_counters);
    APC_CUDA(c, cudaGetLastError());
    x.launches++;
    APC_CUDA(c, cudaMemcpyAsync(h_counters, x.d_counters, sizeof h_counters, cudaMemcpyDeviceToHost, s));
    APC_CUDA(c, cudaStreamSynchronize(s));
    const uint64_t n_sel = h_counters[CNT_SELECTED];
    if (n_needed) *n_needed = n_sel;
    if (solid_km && n_sel > capacity) return APC_ERR_CAPACITY;
    if (!solid_km && n_sel != n_take) return fail(c, APC_ERR_CUDA, "exact stage: selection size mismatch (internal)");
    kmers.resize(n_sel);
    counts.resize(n_sel);
    if (n_sel) {
        APC_CUDA(c, cudaMemcpyAsync(kmers.data(), x.d_sel_k, n_sel * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        APC_CUDA(c, cudaMemcpyAsync(counts.data(), x.d_sel_c, n_sel * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        APC_CUDA(c, cudaStreamSynchronize(s));
    }
    return APC_OK;
}

} // namespace

// Allocates the scratch of the exact stage for up to max_windows k-windows (every window distinct: the upper bound of
// each buffer), so that the first apc_exact_topn of a one-shot run does not spend its time in cudaMalloc.
int exact_reserve(Ctx *c, uint8_t k, uint64_t max_windows) {
    ExactScratch &x = c->exact;
    if (max_windows == 0) return APC_OK;
    if (max_windows >= 0xFFFFFFFFull) return fail(c, APC_ERR_INVALID, "exact stage: more than 2^32-2 windows");
    const size_t kb = k <= 16 ? sizeof(uint32_t) : sizeof(uint64_t);
    int st;
    if ((st = grow_dev(c, x.d_counters, x.counters_cap, CNT_SLOTS * sizeof(unsigned long long)))) return st;
    if ((st = grow_dev(c, x.d_hist, x.hist_cap, 512 * sizeof(unsigned long long)))) return st;
    if ((st = grow_dev(c, x.d_keys_a, x.keys_a_cap, max_windows * kb))) return st;
    if ((st = grow_dev(c, x.d_keys_b, x.keys_b_cap, max_windows * kb))) return st;
    size_t temp_bytes = 0, scan_bytes = 0;
    const unsigned n_pass_blocks = blocks_for(max_windows, kPassTile);
    if (k <= 16)
        APC_CUDA(c, cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                   (unsigned long long)max_windows, 0, 2 * k, c->stream));
    else
        APC_CUDA(c, cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                                   (unsigned long long)max_windows, 0, 2 * k, c->stream));
    APC_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                              (int)n_pass_blocks, c->stream));
    if ((st = grow_dev(c, x.d_temp, x.temp_cap, std::max(temp_bytes, scan_bytes)))) return st;
    if ((st = grow_dev(c, x.d_block_a, x.block_a_cap, (size_t)n_pass_blocks * sizeof(uint32_t)))) return st;
    if ((st = grow_dev(c, x.d_block_b, x.block_b_cap, (size_t)n_pass_blocks * sizeof(uint32_t)))) return st;
    if ((st = grow_dev(c, x.d_start, x.start_cap, (max_windows + 1) * sizeof(uint32_t)))) return st;
    if ((st = grow_dev(c, x.d_dsum, x.dsum_cap, max_windows * sizeof(uint16_t)))) return st;
    return APC_OK;
}

void free_exact_scratch(Ctx *c) {
    ExactScratch &x = c->exact;
    cudaFree(x.d_keys_a);
    cudaFree(x.d_keys_b);
    cudaFree(x.d_temp);
    cudaFree(x.d_start);
    cudaFree(x.d_dsum);
    cudaFree(x.d_block_a);
    cudaFree(x.d_block_b);
    cudaFree(x.d_hist);
    cudaFree(x.d_counters);
    cudaFree(x.d_forb);
    cudaFree(x.d_sel_k);
    cudaFree(x.d_sel_c);
    x = ExactScratch{};
}

int exact_count_select(Ctx *c, uint8_t k, float lc_adjusted, uint64_t lim, uint64_t solid_km,
                       const uint64_t *forbidden, uint64_t n_forbidden, uint64_t capacity,
                       std::vector<uint64_t> &kmers, std::vector<uint64_t> &counts, uint64_t *n_needed,
                       uint64_t *n_distinct, uint64_t *n_had_n) {
    const uint32_t lc_min_sum = apch::lc_min_filtered_sum(k, lc_adjusted);
    cudaEvent_t e0 = c->ev[6], e1 = c->ev[7];
    APC_CUDA(c, cudaEventRecord(e0, c->stream));
    int st = k <= 16 ? run_exact<uint32_t>(c, k, lc_min_sum, lim, solid_km, forbidden, n_forbidden, capacity, kmers,
                                           counts, n_needed, n_distinct, n_had_n)
                     : run_exact<uint64_t>(c, k, lc_min_sum, lim, solid_km, forbidden, n_forbidden, capacity, kmers,
                                           counts, n_needed, n_distinct, n_had_n);
    if (st != APC_OK) return st;
    APC_CUDA(c, cudaEventRecord(e1, c->stream));
    APC_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    APC_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    c->timing.exact_ms = ms;
    c->timing.exact_launches = c->exact.launches;

    // order the survivors with the reference's comparator (:275-305, :400-403)
    apch::pair_vector v(kmers.size());
    for (size_t i = 0; i < v.size(); i++) v[i] = {kmers[i], counts[i]};
    apch::get_most_frequent(v, solid_km ? ~0ull : lim, k);
    kmers.resize(v.size());
    counts.resize(v.size());
    for (size_t i = 0; i < v.size(); i++) {
        kmers[i] = v[i].first;
        counts[i] = v[i].second;
    }
    return APC_OK;
}

} // namespace apc
